"""Device-resident Adam loops for the two training loops on the hot path.

`TrainVoltMagpieModel` with a moving-average mean trains ONE scalar (the likelihood raw_noise, train_utils.py:201-203)
against a fixed covariance and a fixed residual; `TrainVolModel` trains (raw_noise, raw_vol) of the BM vol model;
`TrainDataModel` and the constant / linear / loglinear branches of `TrainVoltMagpieModel` train raw_noise plus the
parameters of the mean (train_utils.py:98-144, 201-227).  All are `train_iters` repetitions of [transform parameters -> exact MLL + analytic gradient (one launch of the fused
CUDA kernel) -> Adam update].  Here one iteration is captured in a CUDA graph and replayed: no per-iteration host
synchronisation, H2D/D2H traffic or Python-side autograd.  The arithmetic is the reference's: GPyTorch's parameter
transforms, MLL / T, torch.optim.Adam defaults (betas 0.9 / 0.999, eps 1e-8, no weight decay, bias correction).
"""
import torch
import torch.nn.functional as F

from . import _lib, ops
from ._lib import S_ALAL, S_ALR, S_DNOISE, S_DRAW, S_MLL, S_TRINV, VOLT_NSCALARS
from .means import _MAMean

_B1, _B2, _EPS = 0.9, 0.999, 1e-8


class _Adam:
    """torch.optim.Adam (defaults) on a few device scalars, written so that a CUDA graph can capture it."""

    def __init__(self, params, lr):
        self.params, self.lr = params, lr
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.t = torch.zeros((), device=params[0].device)

    def step(self, grads):
        self.t += 1.0
        bc1 = 1.0 - torch.pow(torch.full_like(self.t, _B1), self.t)
        bc2 = 1.0 - torch.pow(torch.full_like(self.t, _B2), self.t)
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            m.mul_(_B1).add_(g, alpha=1.0 - _B1)
            v.mul_(_B2).addcmul_(g, g, value=1.0 - _B2)
            denom = v.sqrt() / bc2.sqrt() + _EPS
            p.sub_((self.lr / bc1) * (m / denom))


LARGE_PATH_T = 1536   # api.cu launch_mll_batched: a single series this long takes the multi-CTA path (chol_large.cu)


def _run(iteration, train_iters, printing, scal, capturable=True):
    """Warm up eagerly (workspace allocation is not capturable), capture one iteration, replay the rest.
    capturable=False (the multi-CTA long-series path reads a failure flag back per factorisation attempt, which is illegal
    during stream capture): every iteration is launched eagerly -- still device resident, no host synchronisation."""
    done = 0
    for _ in range(min(2, train_iters)):
        iteration()
        done += 1
        if printing and (done - 1) % 50 == 0:
            print("Iter %d/%d - Loss: %.3f" % (done, train_iters, -float(scal[0, S_MLL])))
    if done == train_iters:
        return
    graph = None
    if capturable:
        prev_stream = torch.cuda.current_stream()
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                iteration()
            graph = g
            done += 1  # capture does not execute; account for the replay below
            graph.replay()
        except Exception as exc:  # noqa: BLE001  (capture refused: stay eager on the GPU, and say so)
            import warnings

            graph = None
            torch.cuda.set_stream(prev_stream)      # a failed capture_end leaves the side stream current
            torch.cuda.synchronize()
            warnings.warn(f"volt_b200.fused: CUDA-graph capture of the Adam iteration failed ({exc}); running eagerly",
                          RuntimeWarning)
    while done < train_iters:
        if graph is not None:
            graph.replay()
        else:
            iteration()
        done += 1
        if printing and (done - 1) % 50 == 0:
            print("Iter %d/%d - Loss: %.3f" % (done, train_iters, -float(scal[0, S_MLL])))


def _fit_noise_ma(model, likelihood, train_x, target, lr, train_iters, printing):
    dev = ops._dev()
    lib = _lib.load()
    spec = model.train_cov.fused()
    if spec is None or spec[0] != "vol":
        return False
    _, x, vol = spec
    T = target.shape[-1]
    xd = ops._f32(x, dev).reshape(-1)
    vd = ops._f32(vol, dev).reshape(1, T)
    with torch.no_grad():
        resid = ops._f32(target - model.mean_module(train_x), dev).reshape(1, T)
    raw = ops._f32(likelihood.raw_noise, dev).reshape(1).clone()
    opt = _Adam([raw], lr)
    noise = torch.empty(1, device=dev)
    scal = torch.empty(1, VOLT_NSCALARS, device=dev)
    alpha = torch.empty(1, T, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    bad = torch.zeros(1, dtype=torch.int32, device=dev)

    def iteration():
        torch.add(F.softplus(raw), 1e-4, out=noise)
        _lib.check(lib.volt_mll_grad_vol(xd.data_ptr(), 0, vd.data_ptr(), ops.VOL_SIGMA, resid.data_ptr(), noise.data_ptr(), 0, 1, T,
                                         1e-6, 3, scal.data_ptr(), alpha.data_ptr(), info.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), "volt_mll_grad_vol")
        bad.add_(info.ne(0).to(torch.int32))
        opt.step([-(scal[0, S_DNOISE] * torch.sigmoid(raw))])   # d(-mll)/d raw_noise

    _run(iteration, train_iters, printing, scal, capturable=T < LARGE_PATH_T)
    if int(bad) != 0:
        raise ops.NotPSDError("training: covariance not positive definite after jitter retries")
    with torch.no_grad():
        likelihood.raw_noise.data = raw.to(likelihood.raw_noise.device).reshape(likelihood.raw_noise.shape)
    return True


def _fit_parametric_mean(model, likelihood, train_x, target, lr, train_iters, printing):
    """TrainDataModel (train_utils.py:98-144) and the constant / linear / loglinear branches of TrainVoltMagpieModel
    (:201-227): the trained tensors are raw_noise and the mean's parameters (ConstantMean: constant; LinearMean /
    LogLinearMean: weights, bias -- means/loglinear_mean.py:5-21).  Per iteration: evaluate the mean (elementwise, T
    values), one launch of the fused-step kernel (MLL, dMLL/draw_noise, alpha), the mean gradients as alpha-weighted sums
    (dMLL/dm = alpha / T, so d(-MLL)/dtheta = -(alpha / T) . dm/dtheta), one Adam update -- all on the device, captured
    once in a CUDA graph and replayed."""
    from . import gp
    from .means import LogLinearMean

    mm = model.mean_module
    dev = ops._dev()
    lib = _lib.load()
    spec = model.train_cov.fused()
    if spec is None or spec[0] != "vol":
        return False
    _, x, vol = spec
    T = target.shape[-1]
    xd = ops._f32(x, dev).reshape(-1)
    tx = ops._f32(train_x, dev).reshape(-1)
    vd = ops._f32(vol, dev).reshape(1, T)
    yd = ops._f32(target, dev).reshape(T)
    raw = ops._f32(likelihood.raw_noise, dev).reshape(1).clone()
    if isinstance(mm, gp.ConstantMean):
        kind, names = "constant", ["constant"]
    elif isinstance(mm, LogLinearMean):
        kind, names = "loglinear", ["weights", "bias"]
    else:
        kind, names = "linear", ["weights", "bias"]
    params = [ops._f32(getattr(mm, n), dev).reshape(1).clone() for n in names]
    opt = _Adam([raw] + params, lr)      # registration order: likelihood first (ExactGP), then mean_module
    resid = torch.empty(1, T, device=dev)
    scal = torch.empty(1, VOLT_NSCALARS, device=dev)
    alpha = torch.empty(1, T, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    bad = torch.zeros(1, dtype=torch.int32, device=dev)
    loss = torch.empty(1, device=dev)
    inv_T = 1.0 / T

    def iteration():
        if kind == "constant":
            torch.sub(yd, params[0], out=resid[0])
            dm = None
        else:
            lin = tx * params[0] + params[1]                       # x @ weights + bias
            if kind == "loglinear":
                inside = lin > 1e-6                                # clamp(min=1e-6): zero gradient below the clamp
                lin_c = lin.clamp(min=1e-6)
                torch.sub(yd, lin_c.log(), out=resid[0])
                dm = inside.to(lin.dtype) / lin_c                  # d log(clamp(.)) / d lin
            else:
                torch.sub(yd, lin, out=resid[0])
                dm = None
        _lib.check(lib.volt_mll_grad_vol_raw(xd.data_ptr(), 0, vd.data_ptr(), ops.VOL_SIGMA, resid.data_ptr(), raw.data_ptr(), 0, 1, T,
                                             1e-6, 3, scal.data_ptr(), alpha.data_ptr(), info.data_ptr(), loss.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream), "volt_mll_grad_vol_raw")
        bad.add_(info.ne(0).to(torch.int32))
        g_m = alpha[0] * (-inv_T)                                  # d(-MLL) / d mean
        if dm is not None:
            g_m = g_m * dm
        grads = [-scal[0, S_DRAW].reshape(1)]
        if kind == "constant":
            grads.append(g_m.sum().reshape(1))
        else:
            grads += [(g_m * tx).sum().reshape(1), g_m.sum().reshape(1)]
        opt.step(grads)

    _run(iteration, train_iters, printing, scal, capturable=T < LARGE_PATH_T)
    if int(bad) != 0:
        raise ops.NotPSDError("training: covariance not positive definite after jitter retries")
    with torch.no_grad():
        likelihood.raw_noise.data = raw.to(likelihood.raw_noise.device).reshape(likelihood.raw_noise.shape)
        for n, v in zip(names, params):
            t = getattr(mm, n)
            t.data = v.to(t.device).reshape(t.shape)
    return True


def _fit_bmgp(model, likelihood, train_x, target, lr, train_iters, printing):
    dev = ops._dev()
    lib = _lib.load()
    T = target.shape[-1]
    xd = ops._f32(train_x, dev).reshape(-1)
    yd = ops._f32(target, dev).reshape(1, T)
    cov = model.covar_module
    raw_noise = ops._f32(likelihood.raw_noise, dev).reshape(1).clone()
    raw_vol = ops._f32(cov.raw_vol, dev).reshape(1).clone()
    opt = _Adam([raw_noise, raw_vol], lr)   # registration order: likelihood first (ExactGP), then covar_module
    noise, vol = torch.empty(1, device=dev), torch.empty(1, device=dev)
    resid = torch.empty(1, T, device=dev)
    scal = torch.empty(1, VOLT_NSCALARS, device=dev)
    alpha = torch.empty(1, T, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    bad = torch.zeros(1, dtype=torch.int32, device=dev)

    def iteration():
        torch.add(F.softplus(raw_noise), 1e-4, out=noise)
        torch.sigmoid(raw_vol, out=vol)
        torch.add(yd, 0.5 * vol * vol * xd, out=resid)          # y - (-1/2 vol^2 x), BMGP.py:20-21
        _lib.check(lib.volt_mll_grad_bm(xd.data_ptr(), vol.data_ptr(), 0, resid.data_ptr(), noise.data_ptr(), 0, 1, T, 1e-6, 3,
                                        scal.data_ptr(), alpha.data_ptr(), info.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream), "volt_mll_grad_bm")
        bad.add_(info.ne(0).to(torch.int32))
        s = scal[0]
        d_noise = s[S_DNOISE]
        # kernel scale: dMLL/ds = 1/2 [(alpha.r - noise alpha.alpha) - (T - noise tr A^-1)] / (s T)
        d_scale = 0.5 * ((s[S_ALR] - noise * s[S_ALAL]) - (T - noise * s[S_TRINV])) / (vol * T)
        # mean path: dMLL/dresid = -alpha / T, dresid/dvol = vol x
        d_mean = (-(alpha[0] / T) * (vol * xd)).sum()
        g_vol = (d_scale + d_mean) * vol * (1.0 - vol)          # sigmoid'
        g_noise = d_noise * torch.sigmoid(raw_noise)
        opt.step([-g_noise.reshape(1), -g_vol.reshape(1)])

    _run(iteration, train_iters, printing, scal, capturable=T < LARGE_PATH_T)
    if int(bad) != 0:
        raise ops.NotPSDError("training: covariance not positive definite after jitter retries")
    with torch.no_grad():
        likelihood.raw_noise.data = raw_noise.to(likelihood.raw_noise.device).reshape(likelihood.raw_noise.shape)
        cov.raw_vol.data = raw_vol.to(cov.raw_vol.device).reshape(cov.raw_vol.shape)
    return True


def try_fused_loop(model, likelihood, train_x, target, lr, train_iters, printing):
    """Returns True when the loop was run on the device-resident path, False when the caller must use the generic one."""
    from . import gp
    from .means import LogLinearMean
    from .models import BMGP, _VoltBase

    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    if isinstance(model, _VoltBase) and isinstance(model.mean_module, _MAMean) and target.ndim == 1 \
            and trainable == ["likelihood.noise_covar.raw_noise"] and model.likelihood is likelihood:
        return _fit_noise_ma(model, likelihood, train_x, target, lr, train_iters, printing)
    if isinstance(model, BMGP) and target.ndim == 1 and model.likelihood is likelihood \
            and trainable == ["likelihood.noise_covar.raw_noise", "covar_module.raw_vol"]:
        return _fit_bmgp(model, likelihood, train_x, target, lr, train_iters, printing)
    if isinstance(model, _VoltBase) and target.ndim == 1 and model.likelihood is likelihood and train_x.ndim == 1 \
            and type(model.mean_module) in (gp.ConstantMean, gp.LinearMean, LogLinearMean):
        mean_names = ["mean_module." + n for n, _ in model.mean_module.named_parameters()]
        if trainable == ["likelihood.noise_covar.raw_noise"] + mean_names and all(p.numel() == 1 for p in model.mean_module.parameters()):
            return _fit_parametric_mean(model, likelihood, train_x, target, lr, train_iters, printing)
    return False
