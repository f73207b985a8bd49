"""Many independent series at once (new capability; semantics = the single-series reference path vmapped over series,
SURVEY.md section 0) and the series-sharded multi-GPU layout.

One process per GPU (torch.distributed, NCCL).  Series are independent, so each rank owns a contiguous block of
B / world series and the only data-path collective is one all-reduce of the scalar loss per MLL evaluation;
rollouts need none."""
import math

import torch
import torch.nn.functional as F

from . import ops
from ._lib import S_DRAW, S_MLL


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
    """Sum a (scalar) tensor over ranks: the single collective of the batched MLL path (blocking form; the training step
    uses the asynchronous one below)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def noise_from_raw(raw_noise):
    """[GPyTorch] GaussianLikelihood: softplus(raw) + 1e-4."""
    return F.softplus(raw_noise) + 1e-4


class PendingLoss:
    """The all-reduced scalar loss of one step.  The collective runs on a side stream so that the next step's kernels are
    not held back by it (nor by the slowest rank); `wait()` orders the CURRENT stream after it and returns the 0-dim
    tensor; `float(loss)` / `loss.item()` do that and read it back."""

    def __init__(self, tensor, event):
        self._t, self._ev = tensor, event

    def wait(self):
        if self._ev is not None:
            torch.cuda.current_stream().wait_event(self._ev)
            self._ev = None
        return self._t

    def item(self):
        return self.wait().item()

    def __float__(self):
        return float(self.wait())


_side = {}   # device index -> (side stream, ring of loss buffers, cursor)
_RING = 8


def _async_loss_all_reduce(partial):
    """partial: 0-dim / (1,) CUDA tensor produced on the current stream.  World size 1: returned as is."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return PendingLoss(partial.reshape(()), None)
    dev = partial.device.index
    if dev not in _side:
        _side[dev] = [torch.cuda.Stream(device=dev), torch.empty(_RING, device=partial.device), 0]
    side, ring, cur = _side[dev]
    _side[dev][2] = (cur + 1) % _RING
    buf = ring[cur:cur + 1]
    produced = torch.cuda.Event()
    produced.record()
    with torch.cuda.stream(side):
        side.wait_event(produced)
        buf.copy_(partial.reshape(1))          # the step's own buffer is free to be overwritten by the next step
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        done = torch.cuda.Event()
        done.record()
    return PendingLoss(buf.reshape(()), done)


def mll_and_grad(x, vol, resid, raw_noise, jitter=1e-6, check=False):
    """One MLL + gradient evaluation for each of the B local series (train_utils.py:247-250 per series): ONE launch
    (plus the prefix-sum kernel) -- the likelihood's softplus transform, dMLL/draw_noise and the rank-local partial of the
    loss are produced by the kernel's epilogue (volt_mll_grad_vol_raw); the 4-byte all-reduce runs on a side stream.

    x (T,), vol (B,T), resid (B,T) = log y - mean, raw_noise (B,).  Returns dict of CUDA tensors:
    mll (B,), draw_noise (B,) = dMLL/draw_noise, alpha (B,T) (dMLL/dmean = alpha/T), info (B,), scalars (B,16), and
    loss = -sum mll over ALL ranks as a PendingLoss (float(loss) / loss.wait())."""
    out = ops.mll_step("vol", x, vol, resid, raw_noise, jitter=jitter, check=check)
    sc = out["scalars"]
    return dict(mll=sc[:, S_MLL], draw_noise=sc[:, S_DRAW], alpha=out["alpha"], info=out["info"],
                loss=_async_loss_all_reduce(out["loss"]), scalars=sc)


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
