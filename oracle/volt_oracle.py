"""CPU oracle for the Volt GP hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch (CPU, float32 by default, float64 on request) restatement of the
reference's algorithm for the path BASELINE.json names: covariance build -> Cholesky ->
exact MLL + hyper-parameter gradients -> posterior -> Monte-Carlo rollout.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import this module, and only as the checker / the CPU arm.  The product package
(`volt_b200`) never imports it and has no CPU fallback.

Pinning status
  * Pinned against the reference's OWN files (VolKernel.py, BMKernel.py, EWMA.py, BMGP.py,
    VoltMagpie.py, train_utils.py, rollout_utils.py) executed unchanged in the build
    container under a GPyTorch stub -> `tests/golden/*.pt` (generator: tests/golden/make_golden.py),
    and against closed-form known-answer tests (tests/test_oracle.py, KAT-1..5).
  * The GPyTorch slice (GaussianLikelihood noise transform, MVN.log_prob, ExactMLL / T,
    psd_safe_cholesky policy, exact prediction, rsample) is restated FROM MEMORY of
    gpytorch 1.6-1.8 (GPyTorch is not in /root/reference and not installable offline):
    for that slice PARITY IS UNPINNED.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
import math
import warnings

import torch
import torch.nn.functional as F

LOG_2PI = math.log(2.0 * math.pi)


# =============================================================================== covariance
def cum_trapz(y, x):
    """voltron/kernels/VolKernel.py:4-10 -- trapezoid-weighted cumulative sum.

    w = dx*[1/2, 1, ..., 1, 1/2] with dx taken from the FIRST TWO grid points only."""
    dx = x[..., 1] - x[..., 0]
    dx = dx if x.ndim == 1 else dx.unsqueeze(-1)
    w = dx * torch.ones_like(x)
    w[..., 0] *= 0.5
    w[..., -1] *= 0.5
    return torch.cumsum(w * y, -1)


def vol_kernel(x, vol_path, diag=False):
    """voltron/kernels/VolKernel.py:18-41 (no last_dim_is_batch) -- K[..., i, j] = V[..., min(i, j)].

    x: (T,) or (B, T) time grid; vol_path: (T,) or (B, T) volatility sigma (not log sigma)."""
    if x.shape[-1] == 1 and x.ndim > 1:
        x = x.squeeze(-1)
    if vol_path.shape[-1] == 1 and vol_path.ndim > 1:
        vol_path = vol_path.squeeze(-1)
    V = cum_trapz(vol_path * vol_path, x)
    if diag:
        return V  # diagonal of V[min(i,i)] (VolKernel.py:39-40)
    T = x.shape[-1]
    idx = torch.arange(T)
    mn = torch.minimum(idx[:, None], idx[None, :])
    return V[..., mn]


def bm_kernel(x1, x2, vol):
    """voltron/kernels/BMKernel.py:38-51 (non-batch branch) -- vol * min(x1_i, x2_j)."""
    x1 = x1.reshape(-1)
    x2 = x2.reshape(-1)
    return vol * torch.minimum(x1[:, None], x2[None, :])


def bm_vol_from_raw(raw_vol):
    """voltron/kernels/BMKernel.py:8,30-32 + [GPyTorch] Interval(0,1).transform = sigmoid."""
    return torch.sigmoid(raw_vol)


def noise_from_raw(raw_noise):
    """[GPyTorch] GaussianLikelihood: noise = softplus(raw_noise) + 1e-4 (GreaterThan(1e-4))."""
    return F.softplus(raw_noise) + 1e-4


# =============================================================================== moving-average means
def ewma_weights(k, dtype=torch.float32):
    """voltron/means/EWMA.py:21-24 -- alpha(1-alpha)^(k-1..0), normalised."""
    alpha = 2.0 / (k + 1)
    w = alpha * (1 - alpha) ** (torch.arange(k - 1, -1, -1))
    w = w.to(torch.float32)
    return (w / w.sum()).to(dtype)


def ewma(y, k):
    """voltron/means/EWMA.py:20-37 -- causal k-tap weighted mean, left-padded with k copies of y[...,0].

    y: (T,) or (S, T) -> (T+1,) or (S, T+1); out[j] = weighted mean of the k values before y[j]."""
    w = ewma_weights(k, y.dtype)
    pad = y[..., :1].expand(*y.shape[:-1], k)
    padded = torch.cat((pad, y), dim=-1)
    batch = y.shape[-2] if y.ndim > 1 else 1
    out = F.conv1d(padded.reshape(batch, 1, -1), w.view(1, 1, -1))
    return out.reshape(*y.shape[:-1], -1)


def ma_mean(kind, train_y, k, theta=0.5):
    """Full-length (T+1) moving-average path for the four mean families.

    ewma: EWMA.py:46-47 | dewma: :81-84 | tewma: :102-106 | meanrevert: :126-128."""
    kind = kind.lower()
    e = ewma(train_y, k)
    if kind == "ewma":
        return e
    if kind == "dewma":
        ee = ewma(e, k)[..., :-1]
        return 2 * e - ee
    if kind == "tewma":
        ee = ewma(e, k)[..., :-1]
        eee = ewma(ee, k)[..., :-1]
        return 3 * e - 3 * ee + eee
    if kind == "meanrevert":
        latent = train_y.mean()
        e = e.clone()
        e[..., 1:] -= theta * (e[..., :-1] - latent)
        return e
    raise ValueError(kind)


def ma_mean_forward(kind, train_x, train_y, k, x, theta=0.5):
    """The `forward(x)` selection rule shared by all MA means (EWMA.py:48-54 and twins)."""
    m = ma_mean(kind, train_y, k, theta)
    if x.numel() == 1:
        return m[..., -1].unsqueeze(0)
    if torch.equal(x.squeeze(), train_x.squeeze()):
        return m[..., :-1]
    return m


def loglinear_mean(x, weights, bias):
    """voltron/means/loglinear_mean.py:18-21 -- log(clamp(x @ w + b, 1e-6))."""
    return (x.reshape(-1, 1).matmul(weights).squeeze(-1) + bias).clamp(min=1e-6).log()


# =============================================================================== Cholesky policy
class NotPSDError(RuntimeError):
    pass


def psd_safe_cholesky(A, jitter=None, max_tries=3, return_jitter=False):
    """[GPyTorch] gpytorch.utils.cholesky.psd_safe_cholesky (call sites rollout_utils.py:35,46;
    VoltMagpie.py:87,92).  Jitter is added ONLY on failure and only to failing batch members,
    escalating jitter*10^i, i = 0..max_tries-1."""
    L, info = torch.linalg.cholesky_ex(A)
    added = torch.zeros(A.shape[:-2], dtype=A.dtype)
    if not torch.any(info):
        return (L, added) if return_jitter else L
    if torch.isnan(A).any():
        raise NotPSDError("matrix contains NaNs")
    if jitter is None:
        jitter = 1e-6 if A.dtype == torch.float32 else 1e-8
    Ap = A.clone()
    prev = 0.0
    for i in range(max_tries):
        new = jitter * (10 ** i)
        fail = (info > 0)
        Ap.diagonal(dim1=-1, dim2=-2).add_((fail * (new - prev)).unsqueeze(-1).expand(*Ap.shape[:-1]))
        added = torch.where(fail, torch.as_tensor(new, dtype=A.dtype), added)
        prev = new
        warnings.warn(f"A not p.d., added jitter of {new:.1e} to the diagonal", RuntimeWarning)
        L, info = torch.linalg.cholesky_ex(Ap)
        if not torch.any(info):
            return (L, added) if return_jitter else L
    raise NotPSDError(f"Matrix not positive definite after repeatedly adding jitter up to {new:.1e}.")


# =============================================================================== exact MLL
def exact_mll(K, resid, noise):
    """[GPyTorch] ExactMarginalLogLikelihood(lh, model)(MVN(mean, K), y), Cholesky branch.
    Call sites: voltron/train_utils.py:80,89 / 127,136 / 240,249.

    A = K + noise*I; MLL = -1/2 (r^T A^-1 r + logdet A + T log 2pi) / T.  Batched over leading dims."""
    T = K.shape[-1]
    noise = torch.as_tensor(noise, dtype=K.dtype)
    A = K + noise.reshape(*noise.shape, 1, 1) * torch.eye(T, dtype=K.dtype) if noise.ndim else K + noise * torch.eye(T, dtype=K.dtype)
    L = psd_safe_cholesky(A)
    z = torch.linalg.solve_triangular(L, resid.unsqueeze(-1), upper=False).squeeze(-1)
    inv_quad = (z * z).sum(-1)
    logdet = 2.0 * torch.diagonal(L, dim1=-2, dim2=-1).log().sum(-1)
    return -0.5 * (inv_quad + logdet + T * LOG_2PI) / T


def exact_mll_and_grad(K, resid, noise):
    """MLL plus the analytic gradients autograd produces at train_utils.py:90,137,250
    (verified against autograd in tests/test_oracle.py).

    Returns dict: mll, dnoise = dMLL/dnoise = tr(G), dresid = dMLL/dr = -alpha/T, alpha = A^-1 r,
    logdet, inv_quad, tr_inv = tr(A^-1), G = 1/2 (alpha alpha^T - A^-1)/T (dense, = dMLL/dK)."""
    T = K.shape[-1]
    noise = torch.as_tensor(noise, dtype=K.dtype)
    eye = torch.eye(T, dtype=K.dtype)
    A = K + (noise.reshape(*noise.shape, 1, 1) if noise.ndim else noise) * eye
    L = psd_safe_cholesky(A)
    r = resid.unsqueeze(-1)
    alpha = torch.cholesky_solve(r, L)
    Ainv = torch.cholesky_inverse(L)
    inv_quad = (r * alpha).sum((-1, -2))
    logdet = 2.0 * torch.diagonal(L, dim1=-2, dim2=-1).log().sum(-1)
    mll = -0.5 * (inv_quad + logdet + T * LOG_2PI) / T
    G = 0.5 * (alpha @ alpha.transpose(-1, -2) - Ainv) / T
    tr_inv = torch.diagonal(Ainv, dim1=-2, dim2=-1).sum(-1)
    return dict(mll=mll, dnoise=torch.diagonal(G, dim1=-2, dim2=-1).sum(-1), dresid=-alpha.squeeze(-1) / T,
                alpha=alpha.squeeze(-1), logdet=logdet, inv_quad=inv_quad, tr_inv=tr_inv, G=G)


def volt_mll_and_grad(x, vol, resid, raw_noise):
    """One MLL+grad evaluation of the data model (train_utils.py:247-250 with the cached
    train_cov of VoltMagpie.py:46,123-124): returns mll and d mll / d raw_noise, d mll / d mean."""
    K = vol_kernel(x, vol)
    raw_noise = torch.as_tensor(raw_noise, dtype=K.dtype)
    out = exact_mll_and_grad(K, resid, noise_from_raw(raw_noise))
    out["draw_noise"] = out["dnoise"] * torch.sigmoid(raw_noise)
    out["dmean"] = -out["dresid"]
    return out


def bm_mll_and_grad(x, y, raw_vol, raw_noise):
    """One MLL+grad evaluation of the vol model BMGP (BMGP.py:20-28; train_utils.py:86-90):
    mean = -1/2 vol^2 x, K = vol*min(x,x'); gradients w.r.t. raw_vol and raw_noise by autograd."""
    raw_vol = torch.as_tensor(raw_vol, dtype=x.dtype).clone().requires_grad_(True)
    raw_noise = torch.as_tensor(raw_noise, dtype=x.dtype).clone().requires_grad_(True)
    vol = bm_vol_from_raw(raw_vol)
    mean = -0.5 * vol.pow(2.0) * x
    K = bm_kernel(x, x, vol)
    mll = exact_mll(K, y - mean, noise_from_raw(raw_noise))
    g = torch.autograd.grad(mll.sum(), [raw_vol, raw_noise])
    return dict(mll=mll.detach(), draw_vol=g[0], draw_noise=g[1])


# =============================================================================== vol-model posterior
def bmgp_posterior(train_x, train_y, test_x, vol, noise):
    """voltron/models/BMGP.py:18-28 in eval mode = [GPyTorch] exact prediction:
    mean* = m* + K*^T (K+noise I)^-1 (y - m); cov* = K** - K*^T (K+noise I)^-1 K*.
    Called at voltron/rollout_utils.py:66 and VoltMagpie.py:101-113."""
    n = train_x.numel()
    full_x = torch.cat([train_x.reshape(-1), test_x.reshape(-1)])
    full_mean = -0.5 * vol ** 2 * full_x
    Kf = bm_kernel(full_x, full_x, vol)
    A = Kf[:n, :n] + noise * torch.eye(n, dtype=Kf.dtype)
    L = psd_safe_cholesky(A)
    alpha = torch.cholesky_solve((train_y - full_mean[:n]).unsqueeze(-1), L).squeeze(-1)
    Kst = Kf[n:, :n]
    mean = full_mean[n:] + Kst @ alpha
    cov = Kf[n:, n:] - Kst @ torch.cholesky_solve(Kst.T, L)
    return mean, cov


def mvn_sample(mean, cov, eps):
    """[GPyTorch] MultivariateNormal.rsample: mean + chol(cov) @ eps, eps (H, S) -> (S, H)."""
    root = psd_safe_cholesky(cov)
    return (root @ eps).T + mean.unsqueeze(0)


# =============================================================================== prediction / rollouts
def generate_prediction(state, test_x, pred_vol, eps, latent_mean=None, theta=0.5, jitter=1e-4):
    """voltron/rollout_utils.py:6-53.

    state: dict(train_x (m,), train_y (m,) or (S,m) log-prices, log_vol_path (m,) or (S,m),
                k, mean_kind, mean_train_x, mean_train_y) -- the attributes the reference reads off `model`.
    test_x (H,), pred_vol (S,H), eps (S,H,1) base normals (reference: torch.randn at :47).
    Returns (S,H) samples (or (S,) squeezed by the caller), plus pred_mean, pred_cov for inspection."""
    vol = state["log_vol_path"].exp()
    train_x = state["train_x"]
    tx = test_x.unsqueeze(0).repeat(train_x.shape[0], 1) if train_x.ndim != test_x.ndim else test_x
    vs = vol.unsqueeze(0).repeat(pred_vol.shape[0], 1) if vol.ndim == 1 else vol
    full_x = torch.cat((train_x, tx), dim=-1)
    full_vol = torch.cat((vs, pred_vol), dim=-1)
    cut = train_x.shape[-1]
    cov = vol_kernel(full_x, full_vol)
    K_tr, K_tr_te, K_te = cov[..., :cut, :cut], cov[..., :cut, cut:], cov[..., cut:, cut:]
    mkind, k = state["mean_kind"], state["k"]
    mtx, mty = state["mean_train_x"], state["mean_train_y"]
    train_mean = ma_mean_forward(mkind, mtx, mty, k, train_x)
    diffs = state["train_y"].unsqueeze(-1) - train_mean.unsqueeze(-1)
    L = psd_safe_cholesky(K_tr, jitter=jitter)
    pred_mean = K_tr_te.transpose(-1, -2).matmul(torch.cholesky_solve(diffs, L))
    tm = ma_mean_forward(mkind, mtx, mty, k, test_x)
    pred_mean = pred_mean + (tm.T if tm.ndim == 2 else tm).unsqueeze(-1)
    if latent_mean is not None:
        pred_mean = pred_mean - theta * (pred_mean - latent_mean)
    pred_cov = K_te - K_tr_te.transpose(-1, -2).matmul(torch.cholesky_solve(K_tr_te, L))
    Lp = psd_safe_cholesky(pred_cov, jitter=jitter)
    samples = Lp @ eps
    if pred_mean.ndim == 1:
        return samples + pred_mean.unsqueeze(-1), pred_mean, pred_cov
    return (samples + pred_mean).squeeze(-1), pred_mean.squeeze(-1), pred_cov


def rollouts(train_x, train_y, log_vol_path, test_x, pred_vol, eps, k, mean_kind="ewma", theta=None):
    """voltron/rollout_utils.py:57-93 -- the autoregressive MC forecast.

    train_x (n,), train_y (n+1,) PRICES, log_vol_path (n,), test_x (H,), pred_vol (S,H) sigma draws
    (reference draws them at :66 from the vol model), eps (S,H) base normals (one per step, :47).
    Returns (S,H) log-price samples."""
    S, H = pred_vol.shape
    latent_mean = None if theta is None else train_y.log().mean()
    th = 0.5 if theta is None else theta
    samples = torch.zeros(S, H, dtype=train_y.dtype)
    logy = train_y[1:].log()
    state = dict(train_x=train_x, train_y=logy, log_vol_path=log_vol_path, k=k, mean_kind=mean_kind,
                 mean_train_x=train_x, mean_train_y=logy)
    s0, _, _ = generate_prediction(state, test_x[0:1], pred_vol[:, 0:1], eps[:, 0].reshape(S, 1, 1), latent_mean, th)
    samples[:, 0] = s0.squeeze()
    stack_y0 = logy.repeat(S, 1)
    stack_v0 = log_vol_path.repeat(S, 1)
    for idx in range(1, H):
        stack_y = torch.cat((stack_y0, samples[:, :idx]), -1)
        stack_vol = torch.cat((stack_v0, pred_vol[:, :idx].log()), -1)
        rolling_x = torch.cat((train_x, test_x[:idx]))
        state = dict(train_x=rolling_x, train_y=stack_y, log_vol_path=stack_vol, k=k, mean_kind=mean_kind,
                     mean_train_x=rolling_x, mean_train_y=stack_y)
        s, _, _ = generate_prediction(state, test_x[idx:idx + 1], pred_vol[:, idx:idx + 1],
                                      eps[:, idx].reshape(S, 1, 1), latent_mean, th)
        samples[:, idx] = s.squeeze()
    return samples


def rollout_closed_form(train_x, train_y, log_vol_path, test_x, pred_vol, eps, k):
    """KAT-3: with exact arithmetic and no jitter the noise-free predictor
    collapses to  next = m_test + (y_last - m_last) + sqrt(dx/2) sigma_test eps.  fp64 check only."""
    S, H = pred_vol.shape
    dx = train_x[1] - train_x[0]
    y = train_y[1:].log().repeat(S, 1)
    out = torch.zeros(S, H, dtype=train_y.dtype)
    for idx in range(H):
        m = ewma(y, k)
        nxt = m[:, -1] + (y[:, -1] - m[:, -2]) + (0.5 * dx).sqrt() * pred_vol[:, idx] * eps[:, idx]
        out[:, idx] = nxt
        y = torch.cat((y, nxt[:, None]), -1)
    return out


# =============================================================================== training loops
# =============================================================================== GPCV (section 8f-1)
# LearnGPCV (voltron/train_utils.py:15-67): a variational GP on the log-volatility f with inducing points = training
# inputs, UnwhitenedVariationalStrategy + CholeskyVariationalDistribution (models/single_task_variational_gp.py:86-107),
# BMKernel prior with ConstantMean, the "exp" VolatilityGaussianLikelihood (likelihoods/volatility_likelihood.py:44-52)
# and the 75-point Gauss-Hermite VariationalELBO.  [GPyTorch slice restated from memory of 1.6-1.8: PARITY UNPINNED]
#   * training-mode call with x == inducing points returns q(u) = N(m, L_S L_S^T) itself (UnwhitenedVariationalStrategy.forward)
#   * prior p(u) = N(c 1, K + 1e-3 I)  (prior_distribution: lazy_covariance_matrix.add_jitter(), default 1e-3)
#   * KL(q || p) = 0.5 [logdet K - logdet S + tr(K^-1 S) + (c - m)^T K^-1 (c - m) - n]   (kl_mvn_mvn, Cholesky branch)
#   * ELBO = (sum_i E_q[log p(y_i | f_i)] - KL) / n  (VariationalELBO, num_data = n, combine_terms=True)
#   * E_q[.] by GaussHermiteQuadrature1D: f = sqrt(2 S_ii) t_k + m_i, (1/sqrt(pi)) sum_k w_k log p(y_i | f)
GPCV_PRIOR_JITTER = 1e-3


def gauss_hermite(n=75, dtype=torch.float32):
    import numpy as np

    t, w = np.polynomial.hermite.hermgauss(n)
    return torch.as_tensor(t, dtype=dtype), torch.as_tensor(w, dtype=dtype)


def gpcv_scaled_returns(train_x, train_y):
    """train_utils.py:16-18."""
    dt = train_x[1] - train_x[0]
    return (train_y[1:] - train_y[:-1]) / (train_y[:-1]) / (dt ** 0.5)


def gpcv_loglik_exp(f, y):
    """volatility_likelihood.py:50-52 with param="exp": Normal(0, exp(f).clamp(min=1e-3)).log_prob(y)."""
    scale = f.exp().clamp(min=1e-3)
    return -scale.log() - 0.5 * LOG_2PI - 0.5 * (y / scale) ** 2


def gpcv_neg_elbo(x, y, var_mean, chol_var, raw_vol, constant, nq=75):
    """-VariationalELBO for one series (train_utils.py:44-56).  chol_var is the full (n,n) parameter; its lower triangle is
    used (CholeskyVariationalDistribution.forward masks it)."""
    n = x.numel()
    Ls = torch.tril(chol_var)
    vol = torch.sigmoid(raw_vol)                       # Interval(0,1) (BMKernel.py:10)
    K = vol * torch.minimum(x.view(-1, 1), x.view(1, -1)) + GPCV_PRIOR_JITTER * torch.eye(n, dtype=x.dtype)
    Lk = torch.linalg.cholesky(K)
    d = (constant - var_mean).unsqueeze(-1)
    A1 = torch.linalg.solve_triangular(Lk, torch.cat((d, Ls), -1), upper=False)
    logdet_k = 2.0 * torch.log(torch.diagonal(Lk)).sum()
    logdet_s = 2.0 * torch.log(torch.diagonal(Ls).abs()).sum()
    kl = 0.5 * (logdet_k - logdet_s + (A1 ** 2).sum() - n)
    t, w = gauss_hermite(nq, x.dtype)
    s_diag = (Ls ** 2).sum(-1)
    f = (2.0 * s_diag).sqrt().unsqueeze(-1) * t + var_mean.unsqueeze(-1)        # (n, nq)
    e = (gpcv_loglik_exp(f, y.unsqueeze(-1)) * w).sum(-1) / math.sqrt(math.pi)
    return -(e.sum() - kl) / n


def gpcv_init(x, y):
    """SingleTaskVariationalGP.initialize_variational_parameters with param="exp" (single_task_variational_gp.py:204-253)
    at the default BMKernel (vol = 0.2).  Returns (variational_mean, chol_variational_covar, constant): the last line of
    the reference routine (:254) also sets the ConstantMean to log(mean(running_std)), running_std AFTER the [:10] patch."""
    n = y.shape[0]
    running_std = torch.stack([y[:i].std(0) for i in range(n)])
    running_std[:10] = running_std[10]
    f = running_std.clamp(min=1e-4).log()
    inverse_hessian = torch.diag_embed(0.5 * y.pow(-2.0) * (f * 2.0).exp()).clamp(min=1e-4, max=1000.0)
    kuu = 0.2 * torch.minimum(x.view(-1, 1), x.view(1, -1))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        kuu_chol = psd_safe_cholesky(kuu)               # LazyTensor.cholesky() -> psd_safe_cholesky (x[0] = 0 => singular)
    inner = kuu_chol.t() @ inverse_hessian @ kuu_chol + torch.eye(n, dtype=x.dtype)       # add_jitter(1.0)
    S = kuu_chol @ torch.linalg.solve(inner, kuu_chol.t())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        S_root = torch.tril(psd_safe_cholesky(S)) * 10.0  # root_decomposition(method="cholesky").root.evaluate().tril() * 10
    return f, S_root, running_std.mean(0).log().reshape(1)


def learn_gpcv(train_x, train_y, train_iters=1000, eps=None, return_state=False, returns=None):
    """LearnGPCV (train_utils.py:15-67).  eps (n, 10): the base normals of the final `likelihood(predictive)` marginal
    (Likelihood.marginal draws settings.num_likelihood_samples = 10 function samples: mean + L_S eps).
    returns: fit these scaled returns directly instead of deriving them from prices (example.ipynb cells 5-8)."""
    x = train_x.reshape(-1)
    y = gpcv_scaled_returns(x, train_y) if returns is None else returns
    f0, s_root, c0 = gpcv_init(x, y)
    vm = f0.clone().requires_grad_(True)
    cv = s_root.clone().requires_grad_(True)
    raw_vol = torch.logit(torch.tensor([0.2], dtype=x.dtype)).requires_grad_(True)       # BMKernel(vol=0.2)
    const = c0.clone().requires_grad_(True)                                                 # ConstantMean, set by the init (:254)
    # model.parameters() order: variational_mean, chol_variational_covar, mean_module.constant, covar_module.raw_vol
    opt = torch.optim.Adam([vm, cv, const, raw_vol], lr=0.01)
    losses = []
    for _ in range(train_iters):
        opt.zero_grad()
        loss = gpcv_neg_elbo(x, y, vm, cv, raw_vol, const)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    with torch.no_grad():
        if eps is None:
            eps = torch.randn(x.numel(), 10, dtype=x.dtype)
        fs = (torch.tril(cv) @ eps).t() + vm                 # MultivariateNormal.rsample((10,))
        pred_scale = fs.exp().clamp(min=1e-3).mean(0)        # likelihood(...).scale.mean(0)
    if return_state:
        return pred_scale, dict(var_mean=vm.detach(), chol_var=cv.detach(), raw_vol=raw_vol.detach(), constant=const.detach(),
                                losses=losses, y=y)
    return pred_scale


# =============================================================================== evaluation reductions (section 8f-3)
def ecdf_logpx(sample_pxs, true_px):
    """voltron/option_utils.py:48-52 -- fraction of sampled prices whose log lies below the log of the realised price."""
    smp = sample_pxs.log().sort()[0]
    log_px = true_px.log()
    return (torch.sum(smp < log_px) / smp.shape[0]).item()


def call_valuation(mc_pxs, strike):
    """voltron/option_utils.py:37 -- np.mean(np.maximum(mc_pxs[:, e] - K, 0)) for one expiry column (numpy: fp32 mean)."""
    return torch.clamp(mc_pxs - strike, min=0).mean(0)


def rollout_stats(samples, truth=None, strike=None, exp=False):
    """Per (series, step) reductions over samples (B,S,H).  ecdf: calib_plotter notebook cell 2
    (`torch.sum(sample < truth, 0) / S`); nll: cell 15 (`-Normal(preds.mean(0), preds.std(0)).log_prob(truth)`);
    payoff: option_utils.py:37.  exp=True applies the notebooks' `preds = preds.exp()` first.  [notebook code is restated,
    it cannot be imported: unpinned, but it is three torch calls]"""
    v = samples.exp() if exp else samples
    out = dict(mean=v.mean(1), std=v.std(1))
    if truth is not None:
        out["ecdf"] = torch.sum(v < truth.unsqueeze(1), 1) / v.shape[1]
        out["nll"] = -torch.distributions.Normal(out["mean"], out["std"]).log_prob(truth)
    if strike is not None:
        out["payoff"] = torch.clamp(v - strike.unsqueeze(1), min=0).mean(1)
    return out


def _adam_loop(params, closure, iters, lr):
    opt = torch.optim.Adam(params, lr=lr)
    trace = []
    for _ in range(iters):
        opt.zero_grad()
        loss = closure()
        loss.backward()
        trace.append(float(loss.detach()))
        opt.step()
    return trace


def train_vol_model(train_x, vol_path, train_iters=1000):
    """voltron/train_utils.py:69-95 (BMGP, Adam lr 0.01; `vol_lh.noise.data=` at :71 is a no-op so
    raw_noise starts at 0; raw_vol starts at logit(0.2), BMKernel.py:8-21)."""
    raw_noise = torch.zeros(1, requires_grad=True)
    raw_vol = torch.logit(torch.tensor([0.2])).requires_grad_(True)
    y = vol_path.log()

    def closure():
        vol = bm_vol_from_raw(raw_vol)
        mean = -0.5 * vol.pow(2.0) * train_x
        return -exact_mll(bm_kernel(train_x, train_x, vol), y - mean, noise_from_raw(raw_noise))

    trace = _adam_loop([raw_noise, raw_vol], closure, train_iters, 0.01)
    return dict(raw_noise=raw_noise.detach(), raw_vol=raw_vol.detach(), loss=trace)


def train_voltmagpie_model(train_x, train_y, vol_path, train_iters=1000, k=25, theta=0.5, mean_func="ewma"):
    """voltron/train_utils.py:192-257 for the MA-mean families: the only trained parameter is the
    likelihood raw_noise (grad_flags :201-203), initialised to the RAW value 1e-5 (:222); Adam lr 0.1.
    train_y are PRICES aligned with train_x (the caller passes train_y[1:])."""
    logy = train_y.log()
    kind = {"ewma": "ewma", "dewma": "dewma", "tewma": "tewma", "meanrevert": "meanrevert"}[mean_func.lower()]
    K = vol_kernel(train_x, vol_path)
    raw_noise = torch.tensor([1e-5], requires_grad=True)

    def closure():
        mean = ma_mean_forward(kind, train_x, logy, k, train_x, theta)
        return -exact_mll(K, logy - mean, noise_from_raw(raw_noise))

    trace = _adam_loop([raw_noise], closure, train_iters, 0.1)
    return dict(raw_noise=raw_noise.detach(), loss=trace)


def train_data_model(train_x, train_y, vol_path, init_weights, train_iters=1000):
    """voltron/train_utils.py:98-144 -- TrainDataModel: VoltronGP (VoltronGP.py:12-50) with a LogLinearMean
    (loglinear_mean.py:5-21; bias initialised to mean(price), weights to the randn draw the caller passes), trainable
    [raw_noise := 1e-5 (RAW, :110), weights, bias] (grad_flags :114), Adam lr 0.1 (:125-127), cached train_cov.
    train_y are PRICES aligned with train_x."""
    logy = train_y.log()
    K = vol_kernel(train_x, vol_path)
    raw_noise = torch.tensor([1e-5], requires_grad=True)
    weights = init_weights.clone().reshape(1, 1).requires_grad_(True)
    bias = logy.exp().mean(-1, keepdim=True).clone().requires_grad_(True)

    def closure():
        return -exact_mll(K, logy - loglinear_mean(train_x, weights, bias), noise_from_raw(raw_noise))

    trace = _adam_loop([raw_noise, weights, bias], closure, train_iters, 0.1)
    with torch.no_grad():
        final = float(closure())
    return dict(raw_noise=raw_noise.detach(), weights=weights.detach(), bias=bias.detach(), loss=trace, final_loss=final)


def model_generate_prediction(train_x, logy, log_vol_path, train_mean, test_mean, test_x, pred_vol, eps):
    """The class-method GeneratePrediction (voltron/models/VoltronGP.py:62-95, twin VoltMagpie.py:67-99): joint draw at
    all H test points for ONE predicted vol path, psd_safe_cholesky with its DEFAULT jitter (no 1e-4 here), no
    likelihood noise, n_sample columns of base normals.  pred_vol (H,), eps (H, n_sample); train_mean (n,), test_mean
    (H,) are mean_module(train_inputs) / mean_module(test_x).  Returns (H, n_sample) ((H,) for n_sample == 1, the
    reference's trailing .squeeze(-1))."""
    full_x = torch.cat((train_x, test_x), dim=-1)
    full_vol = torch.cat((log_vol_path.exp(), pred_vol), dim=-1)
    cut = train_x.shape[-1]
    cov = vol_kernel(full_x, full_vol)
    K_tr, K_tr_te, K_te = cov[:cut, :cut], cov[:cut, cut:], cov[cut:, cut:]
    diffs = (logy - train_mean).unsqueeze(-1)
    L = psd_safe_cholesky(K_tr)
    pred_mean = K_tr_te.T.matmul(torch.cholesky_solve(diffs, L)) + test_mean.unsqueeze(-1)
    pred_cov = K_te - K_tr_te.T.matmul(torch.cholesky_solve(K_tr_te, L))
    samples = psd_safe_cholesky(pred_cov) @ eps
    return (samples + pred_mean).squeeze(-1)


# =============================================================================== synthetic data
def synth_series(B, T, dt=1.0 / 252, seed=2019, dtype=torch.float32):
    """Synthetic (B, T) workload: vol path sigma = exp(BM), g0 = log 0.2, increments N(0, 1.25^2 dt)
    (SABR-like alpha=1.25, V0=0.2, example.ipynb cell 2); log-price GBM y_t = y_{t-1} + sigma_{t-1} sqrt(dt) N(0,1),
    y_0 = log 10.  Per-series generator seed = seed + b.  Returns x (T,), vol (B,T), logy (B,T)."""
    x = (torch.arange(T, dtype=torch.float64) * dt).to(dtype)
    vol = torch.empty(B, T, dtype=torch.float64)
    logy = torch.empty(B, T, dtype=torch.float64)
    for b in range(B):
        g = torch.Generator().manual_seed(seed + b)
        z = torch.randn(2, T, generator=g, dtype=torch.float64)
        lv = math.log(0.2) + torch.cumsum(1.25 * math.sqrt(dt) * z[0], 0) - 1.25 * math.sqrt(dt) * z[0, 0]
        v = lv.exp()
        inc = v[:-1] * math.sqrt(dt) * z[1, 1:]
        vol[b] = v
        logy[b] = math.log(10.0) + torch.cat((torch.zeros(1, dtype=torch.float64), torch.cumsum(inc, 0)))
    return x, vol.to(dtype), logy.to(dtype)
