"""Tensor-level wrappers of the C ABI (device pointers come from torch tensors; torch is only the allocator/stream
provider here) and the autograd bridge for the exact marginal log-likelihood.

Every function accepts CPU or CUDA float tensors: CPU inputs are moved to the current CUDA device and results are
returned on the input's device, which is what lets the reference's CPU-tensor call sites run unchanged.
"""
import torch

from . import _lib
from ._lib import (MA_DEWMA, MA_EWMA, MA_GIVEN, MA_MEANREVERT, MA_TEWMA, S_ALAL, S_ALR, S_DNOISE, S_JITTER, S_MLL, S_TRINV,
                   VOL_LOGSIGMA, VOL_RAW, VOL_SIGMA, VOLT_NSCALARS)

MA_KINDS = {"ewma": MA_EWMA, "dewma": MA_DEWMA, "tewma": MA_TEWMA, "meanrevert": MA_MEANREVERT, "given": MA_GIVEN}


class NotPSDError(RuntimeError):
    """Raised when a matrix is not positive definite after the psd_safe_cholesky jitter retries
    ([GPyTorch] gpytorch.utils.errors.NotPSDError)."""


class NumericalWarning(RuntimeWarning):
    """[GPyTorch] gpytorch.utils.warnings.NumericalWarning."""


def _dev():
    _lib.require_device()
    return torch.device("cuda", torch.cuda.current_device())


def _f32(t, dev):
    """contiguous float32 tensor on `dev` (no copy when it already is)."""
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _back(t, like):
    return t if (torch.is_tensor(like) and like.is_cuda) else t.cpu()


def _empty(shape, dev, dtype=torch.float32):
    return torch.empty(shape, device=dev, dtype=dtype)


# ------------------------------------------------------------------------------------------------ covariance
def cumtrapz(y, x, vol_mode=VOL_RAW, half_last=True):
    """CumTrapz(y, x) -- voltron/kernels/VolKernel.py:4-10."""
    dev = _dev()
    lead = y.shape[:-1]
    T = y.shape[-1]
    yd = _f32(y, dev).reshape(-1, T)
    xd = _f32(x, dev)
    x_batched = int(xd.ndim > 1)
    if x_batched:
        xd = xd.expand(*lead, T).reshape(-1, T).contiguous()
    out = _empty(yd.shape, dev)
    _lib.check(_lib.load().volt_cumtrapz(_ptr(xd), x_batched, _ptr(yd), yd.shape[0], T, vol_mode, int(half_last), _ptr(out),
                                         _stream()), "volt_cumtrapz")
    return _back(out.reshape(*lead, T), y)


def vol_cov(x, vol, add_diag=None, vol_mode=VOL_SIGMA):
    """VolatilityKernel.forward: K[..., i, j] = V[..., min(i, j)] -- voltron/kernels/VolKernel.py:18-41."""
    dev = _dev()
    lead = vol.shape[:-1]
    T = vol.shape[-1]
    vd = _f32(vol, dev).reshape(-1, T)
    xd = _f32(x, dev)
    x_batched = int(xd.ndim > 1)
    if x_batched:
        xd = xd.expand(*lead, T).reshape(-1, T).contiguous()
    B = vd.shape[0]
    ad, astride = None, 0
    if add_diag is not None:
        ad = _f32(add_diag, dev).reshape(-1)
        astride = 0 if ad.numel() == 1 else 1
    K = _empty((B, T, T), dev)
    _lib.check(_lib.load().volt_vol_cov(_ptr(xd), x_batched, _ptr(vd), vol_mode, B, T, _ptr(ad), astride, _ptr(K), _stream()),
               "volt_vol_cov")
    return _back(K.reshape(*lead, T, T), vol)


def bm_cov(x1, x2, vol):
    """BMKernel.forward: vol * min(x1_i, x2_j) -- voltron/kernels/BMKernel.py:38-51."""
    dev = _dev()
    a, b, v = _f32(x1, dev).reshape(-1), _f32(x2, dev).reshape(-1), _f32(vol, dev).reshape(-1)
    K = _empty((a.numel(), b.numel()), dev)
    _lib.check(_lib.load().volt_bm_cov(_ptr(a), a.numel(), _ptr(b), b.numel(), _ptr(v), _ptr(K), _stream()), "volt_bm_cov")
    return _back(K, x1)


# ------------------------------------------------------------------------------------------------ moving averages
def ewma(y, k):
    """EWMA(y, k) -- voltron/means/EWMA.py:20-37.  (..., T) -> (..., T+1)."""
    dev = _dev()
    lead, T = y.shape[:-1], y.shape[-1]
    yd = _f32(y, dev).reshape(-1, T)
    out = _empty((yd.shape[0], T + 1), dev)
    _lib.check(_lib.load().volt_ewma(_ptr(yd), yd.shape[0], T, int(k), _ptr(out), _stream()), "volt_ewma")
    return _back(out.reshape(*lead, T + 1), y)


def ma_mean(kind, y, k, theta=0.5, latent=None, want_resid=False):
    """Full-length (T+1) path of EWMAMean / DEWMAMean / TEWMAMean / MeanRevertingEMAMean -- voltron/means/EWMA.py:39-135."""
    dev = _dev()
    lead, T = y.shape[:-1], y.shape[-1]
    yd = _f32(y, dev).reshape(-1, T)
    S = yd.shape[0]
    kid = MA_KINDS[kind.lower()]
    lat = None
    if kid == MA_MEANREVERT:
        lat = _f32(latent, dev).reshape(-1)
        if lat.numel() == 1:
            lat = lat.expand(S).contiguous()
    out = _empty((S, T + 1), dev)
    resid = _empty((S, T), dev) if want_resid else None
    _lib.check(_lib.load().volt_ma_mean(_ptr(yd), S, T, int(k), kid, float(theta), _ptr(lat), _ptr(out), None, None, _ptr(resid),
                                        _stream()), "volt_ma_mean")
    out = _back(out.reshape(*lead, T + 1), y)
    if want_resid:
        return out, _back(resid.reshape(*lead, T), y)
    return out


# ------------------------------------------------------------------------------------------------ exact MLL
def _check_info(info, jitter_used, what):
    """psd_safe_cholesky reporting: warn when jitter was needed, raise when the retries were exhausted."""
    import warnings

    bad = info.nonzero()
    if bad.numel():
        raise NotPSDError(f"{what}: matrix not positive definite after repeatedly adding jitter "
                          f"(first failing series {int(bad[0])}, leading minor {int(info[bad[0]])})")
    if jitter_used is not None and bool((jitter_used > 0).any()):
        warnings.warn(f"A not p.d., added jitter of {float(jitter_used.max()):.1e} to the diagonal", NumericalWarning)


def mll_grad(kind, x, gen, resid, noise, jitter=1e-6, max_tries=3, vol_mode=VOL_SIGMA, check=True, want_alpha=True):
    """One exact MLL + gradient evaluation per series (fused build + potrf + solves + trtri on the GPU).

    kind: "vol" (gen = vol path (B,T)), "bm" (gen = BM scale (B,) or scalar), "dense" (gen = K (B,T,T)).
    resid (B,T) = y - mean, noise (B,) or scalar.  Returns dict of CUDA tensors: scalars (B,16), alpha (B,T), info (B).
    Call sites replaced: voltron/train_utils.py:89-90,136-137,249-250."""
    dev = _dev()
    lib = _lib.load()
    r = _f32(resid, dev)
    T = r.shape[-1]
    r = r.reshape(-1, T)
    B = r.shape[0]
    nz = _f32(noise, dev).reshape(-1)
    nstride = 0 if nz.numel() == 1 else 1
    if nstride and nz.numel() != B:
        raise ValueError("noise must be a scalar or have one entry per series")
    scal = _empty((B, VOLT_NSCALARS), dev)
    alpha = _empty((B, T), dev) if want_alpha else None
    info = _empty((B,), dev, torch.int32)
    st = _stream()
    if kind == "vol":
        g = _f32(gen, dev).reshape(-1, T)
        if g.shape[0] != B:
            g = g.expand(B, T).contiguous()
        xd = _f32(x, dev)
        xb = int(xd.ndim > 1)
        if xb:
            xd = xd.reshape(-1, T)
            if xd.shape[0] == 1:
                xd = xd.expand(B, T).contiguous()
            elif xd.shape[0] != B:
                raise ValueError(f"a batched time grid needs one row per series: got {xd.shape[0]} rows for {B} series")
        _lib.check(lib.volt_mll_grad_vol(_ptr(xd), xb, _ptr(g), vol_mode, _ptr(r), _ptr(nz), nstride, B, T, float(jitter),
                                         int(max_tries), _ptr(scal), _ptr(alpha), _ptr(info), st), "volt_mll_grad_vol")
    elif kind == "bm":
        xd = _f32(x, dev).reshape(-1)
        sc = _f32(gen, dev).reshape(-1)
        sstride = 0 if sc.numel() == 1 else 1
        _lib.check(lib.volt_mll_grad_bm(_ptr(xd), _ptr(sc), sstride, _ptr(r), _ptr(nz), nstride, B, T, float(jitter),
                                        int(max_tries), _ptr(scal), _ptr(alpha), _ptr(info), st), "volt_mll_grad_bm")
    elif kind == "dense":
        Kd = _f32(gen, dev).reshape(-1, T, T)
        kb = T * T if Kd.shape[0] == B else 0
        _lib.check(lib.volt_mll_grad_dense(_ptr(Kd), kb, T, _ptr(r), _ptr(nz), nstride, B, T, float(jitter), int(max_tries),
                                           _ptr(scal), _ptr(alpha), _ptr(info), st), "volt_mll_grad_dense")
    else:
        raise ValueError(kind)
    if check:
        _check_info(info, scal[:, S_JITTER], "exact MLL")
    return dict(scalars=scal, alpha=alpha, info=info)


def mll_step(kind, x, gen, resid, raw_noise, jitter=1e-6, max_tries=3, vol_mode=VOL_SIGMA, check=True, want_alpha=True,
             exchange=None):
    """The training step of the data model in one launch (volt_mll_grad_vol_raw): like mll_grad(kind="vol") but takes the
    RAW likelihood noise (train_utils.py:222) -- softplus + 1e-4 is applied in-kernel -- and additionally returns
    scalars[:, S_DRAW] = dMLL/draw_noise and `loss` (1,) = -sum_b MLL_b of this batch (fixed summation order).

    exchange = (peer_slot_ptrs_dev, local_slots_ptr, prev_totals tensor | None, lag, world, rank, ring, seq): series-sharded job
    -- the kernel's last CTA also stores the partial into every rank's exchange buffer over peer memory and sums the
    previous step's slots (volt_mll_step_sharded); batched.LossExchange builds it."""
    if kind != "vol":
        raise ValueError("mll_step: only the Volatility-kernel data model has a fused training step")
    dev = _dev()
    lib = _lib.load()
    r = _f32(resid, dev)
    T = r.shape[-1]
    r = r.reshape(-1, T)
    B = r.shape[0]
    raw = _f32(raw_noise, dev).reshape(-1)
    rstride = 0 if raw.numel() == 1 else 1
    if rstride and raw.numel() != B:
        raise ValueError("raw_noise must be a scalar or have one entry per series")
    g = _f32(gen, dev).reshape(-1, T)
    if g.shape[0] != B:
        g = g.expand(B, T).contiguous()
    xd = _f32(x, dev)
    xb = int(xd.ndim > 1)
    if xb:
        xd = xd.reshape(-1, T)
        if xd.shape[0] == 1:
            xd = xd.expand(B, T).contiguous()
        elif xd.shape[0] != B:
            raise ValueError(f"a batched time grid needs one row per series: got {xd.shape[0]} rows for {B} series")
    scal = _empty((B, VOLT_NSCALARS), dev)
    alpha = _empty((B, T), dev) if want_alpha else None
    info = _empty((B,), dev, torch.int32)
    loss = _empty((1,), dev)
    if exchange is None:
        _lib.check(lib.volt_mll_grad_vol_raw(_ptr(xd), xb, _ptr(g), vol_mode, _ptr(r), _ptr(raw), rstride, B, T, float(jitter),
                                             int(max_tries), _ptr(scal), _ptr(alpha), _ptr(info), _ptr(loss), _stream()),
                   "volt_mll_grad_vol_raw")
    else:
        peers, mine, totals, lag, world, rank, ring, seq = exchange
        _lib.check(lib.volt_mll_step_sharded(_ptr(xd), xb, _ptr(g), vol_mode, _ptr(r), _ptr(raw), rstride, B, T, float(jitter),
                                             int(max_tries), _ptr(scal), _ptr(alpha), _ptr(info), _ptr(loss), int(peers) or None,
                                             int(mine) or None, _ptr(totals), int(lag), int(world), int(rank), int(ring),
                                             int(seq) & 0xFFFFFFFF, _stream()),
                   "volt_mll_step_sharded")
    if check:
        _check_info(info, scal[:, S_JITTER], "exact MLL")
    return dict(scalars=scal, alpha=alpha, info=info, loss=loss)


class _ExactMLL(torch.autograd.Function):
    """-> per-series MLL (GPyTorch normalisation, / T).  Analytic backward:
    dMLL/dresid = -alpha/T, dMLL/dnoise = 1/2 (alpha.alpha - tr A^-1)/T, and for the BM kernel
    dMLL/dscale = 1/2 [(alpha.r - noise alpha.alpha) - (T - noise tr A^-1)] / (scale T)."""

    @staticmethod
    def forward(ctx, kind, x, gen, resid, noise, jitter, vol_mode):
        if not torch.is_tensor(noise):
            noise = torch.as_tensor(noise, dtype=torch.float32)
        if not torch.is_tensor(resid):
            resid = torch.as_tensor(resid, dtype=torch.float32)
        out = mll_grad(kind, x, gen, resid, noise, jitter=jitter, vol_mode=vol_mode)
        sc = out["scalars"]
        ctx.kind, ctx.T, ctx.in_dev = kind, resid.shape[-1], resid.device
        ctx.rshape, ctx.nshape = tuple(resid.shape), tuple(noise.shape)
        ctx.gshape = tuple(gen.shape) if kind == "bm" else None
        nz = noise.detach().to(sc.device, torch.float32).reshape(-1)
        g = gen.detach().to(sc.device, torch.float32).reshape(-1) if kind == "bm" else sc.new_zeros(1)
        ctx.save_for_backward(sc, out["alpha"], nz, g)
        return sc[:, S_MLL].clone().to(resid.device).reshape(resid.shape[:-1])

    @staticmethod
    def backward(ctx, grad):
        sc, alpha, noise, scale = ctx.saved_tensors
        T = ctx.T
        gd = grad.to(sc.device, torch.float32).reshape(-1)

        def fold(v, shape):
            n = 1
            for d in shape:
                n *= d
            v = v.sum() if n == 1 else v
            return v.reshape(shape).to(ctx.in_dev)

        d_resid = ((-alpha / T) * gd[:, None]).reshape(ctx.rshape).to(ctx.in_dev)
        d_noise = fold(sc[:, S_DNOISE] * gd, ctx.nshape)
        d_gen = None
        if ctx.kind == "bm":
            ds = 0.5 * ((sc[:, S_ALR] - noise * sc[:, S_ALAL]) - (T - noise * sc[:, S_TRINV])) / (scale * T)
            d_gen = fold(ds * gd, ctx.gshape)
        return None, None, d_gen, d_resid, d_noise, None, None


def exact_mll(kind, x, gen, resid, noise, jitter=1e-6, vol_mode=VOL_SIGMA):
    """Differentiable exact marginal log-likelihood per series (already divided by T).  Gradients flow to resid (mean
    parameters), noise (raw_noise) and, for kind == "bm", gen = the BM scale (raw_vol).  A vol path / dense K is a
    constant, as in the reference where train_cov is detached (VoltMagpie.py:46)."""
    if kind != "bm" and torch.is_tensor(gen) and gen.requires_grad:
        raise NotImplementedError("gradients w.r.t. the covariance generator are only provided for the BM scale")
    return _ExactMLL.apply(kind, x, gen, resid, noise, jitter, vol_mode)


# ------------------------------------------------------------------------------------------------ Cholesky utilities
def potrf(A, jitter=None, max_tries=3, check=True):
    """psd_safe_cholesky(A, jitter) -- [GPyTorch]; call sites voltron/rollout_utils.py:35,46, VoltMagpie.py:87,92.
    Returns (L, info, jitter_used) with L lower-triangular (upper part zeroed), batched over leading dims."""
    dev = _dev()
    T = A.shape[-1]
    lead = A.shape[:-2]
    Ad = _f32(A, dev).reshape(-1, T, T)
    B = Ad.shape[0]
    L = _empty((B, T, T), dev)
    info = _empty((B,), dev, torch.int32)
    ju = _empty((B,), dev)
    if jitter is None:
        jitter = 1e-6
    _lib.check(_lib.load().volt_potrf(_ptr(Ad), T * T, T, None, 0, B, T, float(jitter), int(max_tries), _ptr(L), T * T, T,
                                      _ptr(ju), _ptr(info), _stream()), "volt_potrf")
    if check:
        _check_info(info, ju, "psd_safe_cholesky")
    return _back(L.reshape(*lead, T, T), A), info, ju


def potrs(L, rhs, forward_only=False):
    """torch.cholesky_solve(rhs, L) -- voltron/rollout_utils.py:36,44.  rhs (..., T, nrhs)."""
    dev = _dev()
    T = L.shape[-1]
    Ld = _f32(L, dev).reshape(-1, T, T)
    B = Ld.shape[0]
    R = _f32(rhs, dev).reshape(-1, T, rhs.shape[-1]).clone()
    if R.shape[0] != B:
        R = R.expand(B, T, rhs.shape[-1]).contiguous()
    nrhs = R.shape[-1]
    _lib.check(_lib.load().volt_potrs(_ptr(Ld), T * T, T, B, T, _ptr(R), T * nrhs, nrhs, int(forward_only), _stream()),
               "volt_potrs")
    return _back(R.reshape(*L.shape[:-2], T, nrhs), rhs)


def bmgp_posterior(x, y, xs, vol, noise):
    """BMGP eval-mode posterior at test points xs -- voltron/models/BMGP.py:18-28, rollout_utils.py:66.
    x (T), y (T) or (B,T), xs (H); vol / noise scalar or (B,).  Returns mean (B,H), cov (B,H,H) on the GPU."""
    dev = _dev()
    xd, xsd = _f32(x, dev).reshape(-1), _f32(xs, dev).reshape(-1)
    T, H = xd.numel(), xsd.numel()
    yd = _f32(y, dev).reshape(-1, T)
    B = yd.shape[0]
    v, nz = _f32(vol, dev).reshape(-1), _f32(noise, dev).reshape(-1)
    mean, cov = _empty((B, H), dev), _empty((B, H, H), dev)
    info = _empty((B,), dev, torch.int32)
    _lib.check(_lib.load().volt_bmgp_posterior(_ptr(xd), _ptr(yd), B, T, _ptr(xsd), H, _ptr(v), 0 if v.numel() == 1 else 1,
                                               _ptr(nz), 0 if nz.numel() == 1 else 1, _ptr(mean), _ptr(cov), _ptr(info),
                                               _stream()), "volt_bmgp_posterior")
    _check_info(info, None, "BMGP posterior")
    return mean, cov


def mvn_sample(mean, cov, eps, jitter=1e-6, exp_out=False):
    """MultivariateNormal.sample: mean + psd_safe_cholesky(cov) eps.  mean (B,H), cov (B,H,H), eps (B,H,S) -> (B,S,H)."""
    dev = _dev()
    m, c, e = _f32(mean, dev), _f32(cov, dev), _f32(eps, dev)
    H = m.shape[-1]
    m, c = m.reshape(-1, H), c.reshape(-1, H, H)
    B = m.shape[0]
    e = e.reshape(B, H, -1)
    S = e.shape[-1]
    out = _empty((B, S, H), dev)
    info = _empty((B,), dev, torch.int32)
    _lib.check(_lib.load().volt_mvn_sample(_ptr(m), _ptr(c), _ptr(e), B, H, S, float(jitter), int(exp_out), _ptr(out), _ptr(info),
                                           _stream()), "volt_mvn_sample")
    _check_info(info, None, "MultivariateNormal.sample")
    return out


# ------------------------------------------------------------------------------------------------ tensor-core product
def gemm_nt(A, B, out=None, subtract=False):
    """C (= | -=) A @ B^T per batch member on the tensor cores with fp32-equivalent accuracy (volt_gemm_nt: 3xTF32, TMA in /
    TMA out).  A (..., M, K), B (..., N, K) -> (..., M, N).  Rows are padded to a multiple of 4 floats when needed (the
    tensor maps want 16-byte row strides)."""
    dev = _dev()
    M, K = A.shape[-2:]
    N = B.shape[-2]
    lead = A.shape[:-2]
    Ad, Bd = _f32(A, dev).reshape(-1, M, K), _f32(B, dev).reshape(-1, N, K)
    nb = Ad.shape[0]
    if Bd.shape[0] != nb:
        raise ValueError("gemm_nt: A and B need the same batch shape")

    def pad4(t):
        k = t.shape[-1]
        return t if k % 4 == 0 else torch.nn.functional.pad(t, (0, 4 - k % 4)).contiguous()

    Ad, Bd = pad4(Ad), pad4(Bd)
    ldc = (N + 3) // 4 * 4
    if out is None:
        if subtract:
            raise ValueError("gemm_nt: subtract needs `out`")
        Cp = _empty((nb, M, ldc), dev)
    else:
        Cp = pad4(_f32(out, dev).reshape(nb, M, N))
    _lib.check(_lib.load().volt_gemm_nt(_ptr(Ad), Ad.shape[-1], M * Ad.shape[-1], _ptr(Bd), Bd.shape[-1], N * Bd.shape[-1], _ptr(Cp),
                                        ldc, M * ldc, M, N, K, nb, int(subtract), _stream()), "volt_gemm_nt")
    C = Cp[..., :N]
    if out is not None and C.data_ptr() != out.data_ptr():
        out.copy_(C.reshape(out.shape))
        return out
    return C.reshape(*lead, M, N)


# ------------------------------------------------------------------------------------------------ rollout
def rollout(x, logy, vol, pred_vol, eps=None, mean_kind="ewma", k=25, mr_theta=0.5, mr_latent=None, resid_given=None,
            mean_test=None, theta=None, latent=None, joint=False, jitter=1e-4, seed=0, vol_mode=VOL_SIGMA, check=True):
    """GeneratePrediction / Rollouts on the GPU -- voltron/rollout_utils.py:6-93.

    x (n,), logy (B,n) or (n,), vol (B,n) or (n,), pred_vol (B,S,H) or (S,H), eps like pred_vol or None (Philox).
    Returns samples (B,S,H) (CUDA), draw_info (B,S), series_info (B).  draw_info bits: 1 non-positive pivot in a draw's
    appended rows, 2 pred_cov needed jitter, 4 not PSD after the retries, 8 the draw was repaired by the per-draw
    psd_safe_cholesky fallback (its whole matrix re-factored with jitter, step by step, with the same base normals)."""
    dev = _dev()
    xd = _f32(x, dev).reshape(-1)
    n = xd.numel()
    ly = _f32(logy, dev).reshape(-1, n)
    B = ly.shape[0]
    vd = _f32(vol, dev).reshape(-1, n)
    if vd.shape[0] != B:
        vd = vd.expand(B, n).contiguous()
    pv = _f32(pred_vol, dev)
    H = pv.shape[-1]
    pv = pv.reshape(B, -1, H)
    S = pv.shape[1]
    ep = None if eps is None else _f32(eps, dev).reshape(B, S, H)
    kid = MA_KINDS[mean_kind.lower()]

    def per_series(t):
        if t is None:
            return None
        t = _f32(t, dev).reshape(-1)
        return t.expand(B).contiguous() if t.numel() == 1 else t

    mrl = per_series(mr_latent)
    lat = per_series(latent)
    rg = None if resid_given is None else _f32(resid_given, dev).reshape(B, n)
    mt = None if mean_test is None else _f32(mean_test, dev).reshape(-1, H)
    if mt is not None and mt.shape[0] != B:
        mt = mt.expand(B, H).contiguous()
    out = _empty((B, S, H), dev)
    dinfo = _empty((B, S), dev, torch.int32)
    sinfo = _empty((B,), dev, torch.int32)
    use_theta = int(theta is not None and lat is not None)
    _lib.check(_lib.load().volt_rollout(_ptr(xd), _ptr(ly), _ptr(vd), vol_mode, _ptr(pv), _ptr(ep), B, n, S, H, kid, int(k),
                                        float(mr_theta), _ptr(mrl), _ptr(rg), _ptr(mt), use_theta,
                                        float(theta if theta is not None else 0.0), _ptr(lat), int(joint), float(jitter),
                                        int(seed), _ptr(out), _ptr(dinfo), _ptr(sinfo), _stream()), "volt_rollout")
    if not joint and H > 1 and bool((dinfo & 1).any()):
        if ep is None:   # in-kernel Philox: regenerate the very normals the kernel drew
            ep = _empty((B, S, H), dev)
            _lib.check(_lib.load().volt_rollout_normals(int(seed), B, S, H, 0, _ptr(ep), _stream()), "volt_rollout_normals")
        _redo_flagged_draws(out, dinfo, xd, ly, vd, vol_mode, pv, ep, kid, mean_kind, int(k), mr_theta, mrl, rg, mt, theta, lat, jitter)
    if check:
        if bool(sinfo.any()):
            raise NotPSDError("rollout: training covariance not positive definite after jitter retries")
        if bool((dinfo & 5).any()):
            raise NotPSDError("rollout: a conditional covariance was not positive definite after jitter retries")
    return out, dinfo, sinfo


def _redo_flagged_draws(out, dinfo, xd, ly, vd, vol_mode, pv, ep, kid, mean_kind, k, mr_theta, mrl, rg, mt, theta, lat, jitter):
    """Per-draw psd_safe_cholesky fallback (voltron/rollout_utils.py:35 on a batch of per-draw matrices: jitter is added to
    the WHOLE diagonal of the failing members only).  The rollout kernel shares the n x n factor between the draws of a
    series, so a draw whose appended rows hit a non-positive pivot (draw_info bit 1) cannot be repaired in place; those
    draws are re-run the reference's way -- step by step, each as its own series whose conditioning set grows by one
    point per step, so that the batched potrf kernel's per-matrix jitter retry sees the draw's whole matrix.
    O(H (n+H)^3) per flagged draw; draws are flagged only on degenerate inputs (zero predicted volatility).
    Repaired draws get bit 1 cleared and bit 8 set; a draw that still fails keeps bit 1 and gets bit 4."""
    dev = out.device
    n, H = xd.numel(), pv.shape[-1]
    fb, fs = torch.nonzero(dinfo & 1, as_tuple=True)
    dx = xd[1] - xd[0]
    sig = vd[fb].exp() if vol_mode == VOL_LOGSIGMA else vd[fb].clone()
    yh = ly[fb].clone()
    rh = None if rg is None else rg[fb].clone()
    res = torch.empty(fb.numel(), H, device=dev)
    bad = torch.zeros(fb.numel(), dtype=torch.bool, device=dev)
    soft = torch.zeros(fb.numel(), dtype=torch.int32, device=dev)      # bit 2 of the re-run steps (pred_cov jitter)
    for idx in range(H):
        xh = torch.cat((xd, xd[-1] + dx * torch.arange(1, idx + 1, device=dev)))
        kw = dict(eps=ep[fb, fs, idx].reshape(-1, 1, 1), mean_kind=mean_kind, k=k, mr_theta=mr_theta,
                  mr_latent=None if mrl is None else mrl[fb], theta=theta, latent=None if lat is None else lat[fb],
                  joint=False, jitter=jitter, vol_mode=VOL_SIGMA, check=False)
        if rh is not None:
            kw.update(resid_given=rh, mean_test=mt[fb, idx].reshape(-1, 1))
        o, di, si = rollout(xh, yh, sig, pv[fb, fs, idx].reshape(-1, 1, 1), **kw)
        o = o.reshape(-1)
        bad |= (si != 0) | ((di.reshape(-1) & 4) != 0)
        soft |= di.reshape(-1) & 2
        res[:, idx] = o
        yh = torch.cat((yh, o.unsqueeze(-1)), -1)
        sig = torch.cat((sig, pv[fb, fs, idx].unsqueeze(-1)), -1)
        if rh is not None:
            rh = torch.cat((rh, (o - mt[fb, idx]).unsqueeze(-1)), -1)
    out[fb, fs] = res
    flags = dinfo[fb, fs]
    dinfo[fb, fs] = torch.where(bad, flags | 4, soft | 8)              # repaired: the first run's flags no longer apply


def rollout_stats(samples, truth=None, strike=None, exp=False):
    """Evaluation reductions over a rollout tensor in one pass.

    samples (B,S,H) or (S,H); truth / strike (B,H), (H,) or None.  Returns a dict of (B,H) CUDA tensors:
    `mean`, `std` (unbiased, torch.std), and when truth is given `ecdf` = sum(v < truth, 0) / S
    (voltron/option_utils.py:48-52; weather calibration notebook cell 2) and `nll` = -Normal(mean, std).log_prob(truth)
    (notebook cell 15); when strike is given `payoff` = mean(max(v - strike, 0)) (option_utils.py:37).
    exp=True evaluates v = exp(sample) (the notebooks' exp=True: samples are log values)."""
    dev = _dev()
    sm = _f32(samples, dev)
    if sm.dim() == 2:
        sm = sm.unsqueeze(0)
    B, S, H = sm.shape
    sm = sm.contiguous()

    def bh(t):
        if t is None:
            return None
        t = _f32(t, dev)
        if t.dim() == 0:
            t = t.reshape(1, 1)
        return t.reshape(-1, t.shape[-1]).expand(B, H).contiguous()

    tr, st = bh(truth), bh(strike)
    out = {"mean": _empty((B, H), dev), "std": _empty((B, H), dev)}
    if tr is not None:
        out["ecdf"] = _empty((B, H), dev)
        out["nll"] = _empty((B, H), dev)
    if st is not None:
        out["payoff"] = _empty((B, H), dev)
    _lib.check(_lib.load().volt_rollout_stats(_ptr(sm), B, S, H, _ptr(tr), _ptr(st), int(bool(exp)), _ptr(out.get("ecdf")),
                                              _ptr(out["mean"]), _ptr(out["std"]), _ptr(out.get("nll")), _ptr(out.get("payoff")),
                                              _stream()), "volt_rollout_stats")
    return out

