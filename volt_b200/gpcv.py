"""GPCV stage on the GPU -- LearnGPCV (voltron/train_utils.py:15-67), the step before the hot path that produces the vol
path.  B independent series are trained at once, device resident, with analytic gradients:

  * the T^3 pieces per iteration -- factorisation of K = vol min(x,x') + 1e-3 I, logdet K, tr K^-1, K^-1 (c - m) and the
    inverse factor (L^-1)^T -- come from ONE launch of the batched tensor-core MLL kernel (`volt_mll_grad_bm_inv`);
  * W = K^-1 L_S = U (U^T L_S) is two launches of this library's TMA-fed tensor-core product (`volt_gemm_nt`, 3xTF32);
  * `volt_gpcv_rows` turns them into the per-point Gauss-Hermite likelihood terms, the loss pieces and the gradient of the
    T x T variational factor; `volt_adam_step` is the optimiser over the flat parameter buffer.

[GPyTorch slice restated from memory, see oracle.volt_oracle; pinned in round 2 against the loss trace real GPyTorch printed
in the reference's example.ipynb (tests/notebook_data.py).]
"""
import math

import numpy as np
import torch

from . import _lib, ops
from ._lib import S_INVQUAD, S_LOGDET, S_TRINV

import os

_AB_BMM = os.environ.get("VOLT_GPCV_BMM") == "1"
PRIOR_JITTER = 1e-3      # [GPyTorch] UnwhitenedVariationalStrategy.prior_distribution: add_jitter() default
NUM_GH = 75              # train_utils.py:52
NUM_LIK_SAMPLES = 10     # [GPyTorch] settings.num_likelihood_samples (Likelihood.marginal)


def scaled_returns(train_x, train_y):
    """train_utils.py:16-18.  train_x (n,), train_y (..., n+1) prices -> (..., n)."""
    dt = train_x[1] - train_x[0]
    return (train_y[..., 1:] - train_y[..., :-1]) / train_y[..., :-1] / (dt ** 0.5)


def _running_std(y):
    """stack([y[:i].std(0) for i in range(n)]) along the last axis without the O(n^2) loop (single_task_variational_gp.py:216)."""
    n = y.shape[-1]
    yd = y.double()
    c1 = torch.cumsum(yd, -1)
    c2 = torch.cumsum(yd * yd, -1)
    cnt = torch.arange(1, n + 1, dtype=torch.float64, device=y.device)
    var = (c2 - c1 * c1 / cnt) / (cnt - 1)                     # unbiased variance of y[:i+1]; NaN (0/0) for a single value
    out = torch.full_like(yd, float("nan"))
    out[..., 2:] = var[..., 1:-1].clamp_min(0).sqrt()          # entry i uses y[:i]
    return out.to(y.dtype)


def init_variational(x, y, vol0=0.2):
    """SingleTaskVariationalGP.initialize_variational_parameters, param="exp" (single_task_variational_gp.py:204-253).
    x (n,), y (B,n) on the GPU -> (variational_mean (B,n), chol_variational_covar (B,n,n), constant (B,)): the routine's last
    line (:254) sets the ConstantMean of the prior to log(mean(running_std)), running_std after the [:10] patch."""
    B, n = y.shape
    rs = _running_std(y)
    rs[..., :10] = rs[..., 10:11]
    f = rs.clamp(min=1e-4).log()
    ih = (0.5 * y.pow(-2.0) * (f * 2.0).exp()).clamp(min=1e-4, max=1000.0)          # diagonal of the inverse Hessian
    kuu = vol0 * torch.minimum(x.view(-1, 1), x.view(1, -1))
    Lk, _, _ = ops.potrf(kuu, check=False)                     # LazyTensor.cholesky() -> psd_safe_cholesky (x[0] = 0: singular)
    # inner = Lk^T diag(ih) Lk + I (clamp(min=1e-4) fills the off-diagonal of the diag_embed'ed matrix with 1e-4 as well)
    H = torch.diag_embed(ih).clamp(min=1e-4, max=1000.0)
    inner = Lk.t().unsqueeze(0) @ H @ Lk.unsqueeze(0) + torch.eye(n, device=y.device)
    Li, _, _ = ops.potrf(inner, check=False)
    S = Lk.unsqueeze(0) @ ops.potrs(Li, Lk.t().unsqueeze(0).expand(B, n, n).contiguous())
    S_root, _, _ = ops.potrf(S, check=False)                   # root_decomposition(method="cholesky")
    return f, torch.tril(S_root) * 10.0, rs.mean(-1).log()


class _State:
    pass


def learn_gpcv(train_x, train_y, train_iters=1000, lr=0.01, eps=None, printing=False, return_state=False, use_graph=True,
               returns=None):
    """Batched LearnGPCV.  train_x (n,), train_y (B,n+1) or (n+1,) prices (or `returns` (B,n): the scaled returns themselves,
    as example.ipynb cells 5-8 fit them).  Returns pred_scale (B,n) on the GPU
    (train_utils.py:62-67: `likelihood(model(train_x), return_gaussian=False).scale.mean(0)` over 10 function samples;
    eps (B,n,10) fixes their base normals).  One Adam iteration (12 launches) is captured in a CUDA graph and replayed."""
    dev = ops._dev()
    lib = _lib.load()
    x = ops._f32(train_x, dev).reshape(-1)
    n = x.numel()
    if returns is None:
        y = scaled_returns(x, ops._f32(train_y, dev).reshape(-1, n + 1)).contiguous()
    else:
        y = ops._f32(returns, dev).reshape(-1, n).contiguous()
    B = y.shape[0]
    vm0, cv0, c0 = init_variational(x, y)
    # flat parameter buffer: [variational_mean | chol_variational_covar | constant | raw_vol]  (model.parameters() order)
    nm, nc = B * n, B * n * n
    P = torch.empty(nm + nc + 2 * B, device=dev)
    P[:nm] = vm0.reshape(-1)
    P[nm:nm + nc] = cv0.reshape(-1)
    P[nm + nc:nm + nc + B] = c0                                                    # ConstantMean, set by the init (:254)
    P[nm + nc + B:] = math.log(0.2 / 0.8)                                          # BMKernel(vol=0.2): logit
    G = torch.zeros_like(P)
    M1 = torch.zeros_like(P)
    M2 = torch.zeros_like(P)
    vm, cv = P[:nm].view(B, n), P[nm:nm + nc].view(B, n, n)
    const, raw_vol = P[nm + nc:nm + nc + B], P[nm + nc + B:]
    g_vm, g_cv = G[:nm].view(B, n), G[nm:nm + nc].view(B, n, n)
    g_const, g_raw = G[nm + nc:nm + nc + B], G[nm + nc + B:]
    t, w = np.polynomial.hermite.hermgauss(NUM_GH)
    gh_t = torch.as_tensor(t, dtype=torch.float32, device=dev)
    gh_w = torch.as_tensor(w, dtype=torch.float32, device=dev)
    jit = torch.full((B,), PRIOR_JITTER, device=dev)
    rows = torch.empty(B, n, 6, device=dev)
    U = torch.empty(B, n, n, device=dev)
    Qt = torch.empty(B, n, n, device=dev)
    W = torch.empty(B, n, n, device=dev)
    sc = torch.empty(B, 16, device=dev)
    alpha = torch.empty(B, n, device=dev)
    info = torch.empty(B, dtype=torch.int32, device=dev)
    inv_n = 1.0 / n
    losses = []
    t_dev = torch.zeros(1, device=dev)                                             # Adam step counter (device: graph replay)
    loss_dev = torch.zeros(B, device=dev)
    want_loss = printing or return_state

    def iteration():
        st = ops._stream()
        t_dev.add_(1.0)
        vol = torch.sigmoid(raw_vol)
        d = (const.unsqueeze(-1) - vm).contiguous()
        # one fused build + potrf + trtri of K = vol min(x,x') + 1e-3 I per series: logdet K, d^T K^-1 d, tr K^-1,
        # alpha = K^-1 d and the inverse factor U = (L^-1)^T
        _lib.check(lib.volt_mll_grad_bm_inv(x.data_ptr(), vol.data_ptr(), 1, d.data_ptr(), jit.data_ptr(), 1, B, n, 1e-6, 3,
                                            sc.data_ptr(), alpha.data_ptr(), info.data_ptr(), U.data_ptr(), st), "volt_mll_grad_bm_inv")
        # K^-1 L_S = U (U^T L_S) as two tensor-core products of this library (volt_gemm_nt: 3xTF32, TMA-fed; round 1 used
        # torch.bmm / cuBLAS).  The kernel contracts over the last index of both operands, so the first product is formed
        # transposed: Q^T = L_S^T X^T with X = U^T, then W = U Q.
        if _AB_BMM:      # developer A/B switch (VOLT_GPCV_BMM=1): the round-1 library GEMMs
            W.copy_(torch.bmm(U, torch.bmm(U.transpose(1, 2), torch.tril(cv))))
        else:
            ops.gemm_nt(torch.tril(cv).transpose(1, 2), U.transpose(1, 2), out=Qt)
            ops.gemm_nt(U, Qt, out=W)
        _lib.check(lib.volt_gpcv_rows(cv.data_ptr(), W.data_ptr(), vm.data_ptr(), y.data_ptr(), gh_t.data_ptr(), gh_w.data_ptr(),
                                      NUM_GH, B, n, inv_n, g_cv.data_ptr(), rows.data_ptr(), st), "volt_gpcv_rows")
        rs = rows.sum(1)                                                           # (B,6)
        sumE, trKS, WW, logdetS = rs[:, 0], rs[:, 2], rs[:, 3], 2.0 * rs[:, 4]
        q, logdetK, trKinv = sc[:, S_INVQUAD], sc[:, S_LOGDET], sc[:, S_TRINV]
        g_vm.copy_((-rows[:, :, 1] - alpha) * inv_n)
        g_const.copy_(alpha.sum(-1) * inv_n)
        dkl_dvol = (n - PRIOR_JITTER * trKinv - trKS - q + PRIOR_JITTER * (WW + (alpha * alpha).sum(-1))) / (2.0 * vol)
        g_raw.copy_(dkl_dvol * vol * (1.0 - vol) * inv_n)
        if want_loss:
            kl = 0.5 * (logdetK - logdetS + trKS + q - n)
            loss_dev.copy_(-(sumE - kl) * inv_n)
        _lib.check(lib.volt_adam_step(P.data_ptr(), G.data_ptr(), M1.data_ptr(), M2.data_ptr(), P.numel(), lr, 0.9, 0.999, 1e-8, 0,
                                      t_dev.data_ptr(), st), "volt_adam_step")

    def after(it):
        if return_state:
            losses.append(loss_dev.clone())
        if printing and (it - 1) % 50 == 0:
            print("Iter %d/%d - Loss: %.3f" % (it, train_iters, float(loss_dev.mean())))

    # warm up eagerly (workspace / cuBLAS handles are not capturable), capture one iteration in a CUDA graph, replay it
    done = 0
    for _ in range(min(2, train_iters)):
        iteration()
        done += 1
        after(done)
    graph = None
    if done < train_iters and use_graph:
        prev_stream = torch.cuda.current_stream()
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                iteration()
            graph = g
        except Exception as exc:  # noqa: BLE001  (capture refused: stay eager on the GPU, and say so)
            import warnings

            graph = None
            torch.cuda.set_stream(prev_stream)      # a failed capture_end leaves the side stream current
            torch.cuda.synchronize()
            warnings.warn(f"volt_b200.gpcv: CUDA-graph capture of the Adam iteration failed ({exc}); running eagerly", RuntimeWarning)
    while done < train_iters:
        if graph is not None:
            graph.replay()
        else:
            iteration()
        done += 1
        after(done)
    if eps is None:
        eps = torch.randn(B, n, NUM_LIK_SAMPLES, device=dev)
    else:
        eps = ops._f32(eps, dev).reshape(B, n, -1)
    fs = (torch.tril(cv) @ eps).transpose(-1, -2) + vm.unsqueeze(1)                 # (B,10,n): MultivariateNormal.rsample
    pred_scale = fs.exp().clamp(min=1e-3).mean(1)
    if return_state:
        s = _State()
        s.var_mean, s.chol_var, s.constant, s.raw_vol, s.y = vm.clone(), cv.clone(), const.clone(), raw_vol.clone(), y
        s.losses = torch.stack(losses, 0) if losses else None
        return pred_scale, s
    return pred_scale
