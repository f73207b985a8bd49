"""voltron.rollout_utils on the B200 path: GeneratePrediction and Rollouts (voltron/rollout_utils.py:6-93).

The reference re-factorises an (S, n+idx, n+idx) batch of covariance matrices at every horizon step; here one call of
volt_rollout does the whole forecast: the shared n x n factor once per series (batched potrf kernel) and a bordered
per-draw update in the rollout kernel (see csrc/rollout.cu)."""
import torch

from . import ops
from .means import _MAMean


def _mean_spec(model, train_x_for_mean, train_y, test_x):
    """Translate model.mean_module into the rollout kernel's mean description."""
    mm = model.mean_module
    if isinstance(mm, _MAMean):
        spec = dict(mean_kind=mm.kind, k=mm.k)
        if mm.kind == "meanrevert":
            spec.update(mr_theta=mm.theta, mr_latent=mm.latent_mean)
        return spec
    # parametric mean (constant / linear / log-linear): evaluate it (tiny) and hand residual + test means over
    with torch.no_grad():
        resid = train_y - mm(train_x_for_mean)
        mt = mm(test_x)
    return dict(mean_kind="given", k=0, resid_given=resid, mean_test=mt)


def _generate_prediction(model, test_x, pred_vol, eps, latent_mean, theta, jitter, train_x, train_y, log_vol_path,
                         train_inputs_for_mean=None):
    """Common body: pred_vol (S,H), eps (S,H) -> samples (S,H) on the CPU/GPU device of test_x."""
    S, H = pred_vol.shape
    per_draw_state = train_y.ndim > 1 or log_vol_path.ndim > 1
    spec = _mean_spec(model, train_x if train_inputs_for_mean is None else train_inputs_for_mean, train_y, test_x)
    if H > 1 and spec["mean_kind"] != "given":
        raise RuntimeError("the moving-average means support a single test point per call (voltron/means/EWMA.py:48-54)")
    kw = dict(eps=None, joint=H > 1, jitter=jitter, vol_mode=ops.VOL_LOGSIGMA,
              theta=theta if latent_mean is not None else None, latent=latent_mean)
    kw.update(spec)
    if per_draw_state:
        # every draw has its own history: treat the S draws as S independent series with one draw each
        ly = train_y if train_y.ndim > 1 else train_y.unsqueeze(0).expand(S, -1)
        lv = log_vol_path if log_vol_path.ndim > 1 else log_vol_path.unsqueeze(0).expand(S, -1)
        if kw.get("mr_latent") is not None:
            kw["mr_latent"] = torch.as_tensor(kw["mr_latent"]).reshape(-1)[:1].expand(S)
        if latent_mean is not None:
            kw["latent"] = torch.as_tensor(latent_mean).reshape(-1)[:1].expand(S)
        kw["eps"] = eps.reshape(S, 1, H)
        out, _, _ = ops.rollout(train_x, ly, lv, pred_vol.reshape(S, 1, H), **kw)
        out = out.reshape(S, H)
    else:
        kw["eps"] = eps.reshape(1, S, H)
        out, _, _ = ops.rollout(train_x, train_y, log_vol_path, pred_vol.reshape(1, S, H), **kw)
        out = out.reshape(S, H)
    return out.to(test_x.device)


def GeneratePrediction(train_x, train_y, test_x, pred_vol, model, latent_mean=None, theta=0.5):
    """voltron/rollout_utils.py:6-53.  Reads model.train_x / train_y / log_vol_path / mean_module exactly like the
    reference (train_x / train_y arguments are unused there too); pred_vol (S, H); returns (S, H) ((S,) squeezed
    by the caller for H = 1)."""
    S, H = pred_vol.shape[0], test_x.shape[0]
    eps = torch.randn(S, H, 1).to(test_x.device)  # same draw shape/order as rollout_utils.py:47
    out = _generate_prediction(model, test_x, pred_vol, eps.reshape(S, H), latent_mean, theta, 1e-4,
                               model.train_x, model.train_y, model.log_vol_path)
    torch.cuda.empty_cache()
    return out


def Rollouts(train_x, train_y, test_x, model, nsample=50, method="volt", theta=None):
    """voltron/rollout_utils.py:57-93 -- autoregressive Monte-Carlo forecast; returns a CPU (nsample, ntest) tensor of
    log prices and leaves `model` in the state the reference leaves it in (grown train_x / train_y / log_vol_path)."""
    if method != "volt":
        raise NotImplementedError("nonvol_rollouts (BoTorch baselines) is outside the hot path (SURVEY.md 2.1 #12)")
    latent_mean = None if theta is None else train_y.log().mean()
    ntest = test_x.numel()
    pred_vol = model.vol_model(test_x).sample(torch.Size((nsample,))).exp()          # (S, H), rollout_utils.py:66
    # one base normal per (step, draw), drawn in the reference's order: step-major (rollout_utils.py:47 per call)
    eps = torch.randn(ntest, nsample).t().contiguous()
    logy = train_y[1:].log()
    spec = _mean_spec(model, train_x, logy, test_x)
    if spec["mean_kind"] == "given":
        raise RuntimeError("Rollouts needs a moving-average mean: parametric means return one value per test point "
                           "(use GeneratePrediction with all test points, GenerateMultiMeanPreds.py:114-118)")
    out, _, _ = ops.rollout(train_x, logy, model.log_vol_path, pred_vol.reshape(1, nsample, ntest),
                            eps=eps.reshape(1, nsample, ntest), joint=False, jitter=1e-4, vol_mode=ops.VOL_LOGSIGMA,
                            theta=theta, latent=latent_mean, **spec)
    samples = out.reshape(nsample, ntest).cpu()
    if ntest > 1:
        # reproduce the reference's side effects on the model (rollout_utils.py:75-86, state after the last step)
        idx = ntest - 1
        stack_y = torch.cat((logy.repeat(nsample, 1), samples[:, :idx].to(logy.device)), -1)
        stack_vol = torch.cat((model.log_vol_path.repeat(nsample, 1) if model.log_vol_path.ndim == 1 else model.log_vol_path,
                               pred_vol[:, :idx].to(model.log_vol_path.device).log()), -1)
        rolling_x = torch.cat((train_x, test_x[:idx]))
        model.mean_module.train_y = stack_y
        model.mean_module.train_x = rolling_x
        model.train_x = rolling_x
        model.train_y = stack_y
        model.log_vol_path = stack_vol
    torch.cuda.empty_cache()
    return samples
