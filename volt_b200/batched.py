"""Many independent series at once (new capability; semantics = the single-series reference path vmapped over series,
SURVEY.md section 0) and the series-sharded multi-GPU layout.

One process per GPU (torch.distributed, NCCL).  Series are independent, so each rank owns a contiguous block of
B / world series and the only data-path collective is one all-reduce of the scalar loss per MLL evaluation;
rollouts need none."""
import math

import torch
import torch.nn.functional as F

from . import ops
from ._lib import S_DNOISE, S_MLL


def shard_bounds(n_items, rank, world):
    """Contiguous block [lo, hi) of `n_items` owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def dist_info():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_reduce_sum(t):
    """Sum a (scalar) tensor over ranks: the single collective of the batched MLL path."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def noise_from_raw(raw_noise):
    """[GPyTorch] GaussianLikelihood: softplus(raw) + 1e-4."""
    return F.softplus(raw_noise) + 1e-4


def mll_and_grad(x, vol, resid, raw_noise, jitter=1e-6, check=False):
    """One MLL + gradient evaluation for each of the B local series (train_utils.py:247-250 per series).

    x (T,), vol (B,T), resid (B,T) = log y - mean, raw_noise (B,).  Returns dict of CUDA tensors:
    mll (B,), draw_noise (B,) = dMLL/draw_noise, alpha (B,T) (dMLL/dmean = alpha/T), info (B,), loss = -sum mll
    all-reduced over ranks (a 0-dim tensor)."""
    noise = noise_from_raw(raw_noise)
    out = ops.mll_grad("vol", x, vol, resid, noise, jitter=jitter, check=check)
    sc = out["scalars"]
    mll = sc[:, S_MLL]
    loss = all_reduce_sum(-mll.sum())
    return dict(mll=mll, draw_noise=sc[:, S_DNOISE] * torch.sigmoid(raw_noise.to(sc.device)), alpha=out["alpha"],
                info=out["info"], loss=loss, scalars=sc)


def train_noise(x, vol, logy, k=25, mean_func="ewma", train_iters=300, lr=0.1, raw_init=1e-5):
    """TrainVoltMagpieModel for B series at once (train_utils.py:192-257, MA-mean families: the only trained parameter
    is each series' raw_noise).  Adam (torch defaults, lr 0.1) is applied element-wise to the (B,) vector, which is
    exactly B independent scalar Adam optimisers.  Returns raw_noise (B,), final loss per series (B,)."""
    dev = ops._dev()
    B = logy.shape[0]
    _, resid = ops.ma_mean(mean_func, logy.to(dev), k, want_resid=True)
    raw = torch.full((B,), float(raw_init), device=dev)
    m = torch.zeros_like(raw)
    v = torch.zeros_like(raw)
    b1, b2, eps = 0.9, 0.999, 1e-8
    loss = None
    for it in range(1, train_iters + 1):
        out = mll_and_grad(x, vol, resid, raw)
        g = -out["draw_noise"]          # gradient of the loss -mll
        loss = -out["mll"]
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        mhat = m / (1 - b1 ** it)
        vhat = v / (1 - b2 ** it)
        raw = raw - lr * mhat / (vhat.sqrt() + eps)
    return raw, loss


def rollouts(x, logy, vol, pred_vol, eps=None, k=25, mean_func="ewma", theta=None, latent=None, seed=0, check=False):
    """Rollouts for B series x S draws x H steps (rollout_utils.py:57-93 per series).  Series-sharded callers pass
    their local block; no collective is involved.  Returns samples (B,S,H) on the GPU."""
    out, dinfo, sinfo = ops.rollout(x, logy, vol, pred_vol, eps=eps, mean_kind=mean_func, k=k, theta=theta, latent=latent,
                                    joint=False, jitter=1e-4, seed=seed, vol_mode=ops.VOL_SIGMA, check=check)
    return out, dinfo, sinfo


def rollout_stats(samples, truth=None, strike=None, exp=False):
    """Per-series, per-step evaluation of a (B,S,H) rollout tensor without leaving the GPU: sample percentile of the
    realised value (ECDF), moment-matched Gaussian NLL, mean / std and Monte-Carlo call payoff (ops.rollout_stats).
    Series-sharded callers pass their local block; no collective is involved."""
    return ops.rollout_stats(samples, truth=truth, strike=strike, exp=exp)


def synth_series(B, T, dt=1.0 / 252, seed=2019, start=0):
    """Synthetic workload of SURVEY.md section 8d (same generator as oracle.volt_oracle.synth_series, restated so the
    product never imports the oracle): vol = exp(BM) with V0 = 0.2, alpha = 1.25; log price a GBM from log 10.
    Series b uses generator seed `seed + start + b`, so ranks can build their own block."""
    x = (torch.arange(T, dtype=torch.float64) * dt).to(torch.float32)
    vol = torch.empty(B, T, dtype=torch.float64)
    logy = torch.empty(B, T, dtype=torch.float64)
    sq = math.sqrt(dt)
    for b in range(B):
        g = torch.Generator().manual_seed(seed + start + b)
        z = torch.randn(2, T, generator=g, dtype=torch.float64)
        lv = math.log(0.2) + torch.cumsum(1.25 * sq * z[0], 0) - 1.25 * sq * z[0, 0]
        v = lv.exp()
        vol[b] = v
        logy[b] = math.log(10.0) + torch.cat((torch.zeros(1, dtype=torch.float64), torch.cumsum(v[:-1] * sq * z[1, 1:], 0)))
    return x, vol.to(torch.float32), logy.to(torch.float32)
