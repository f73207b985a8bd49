"""Many independent series at once (new capability; semantics = the single-series reference path vmapped over series,
SURVEY.md section 0) and the series-sharded multi-GPU layout.

One process per GPU (torch.distributed, NCCL).  Series are independent, so each rank owns a contiguous block of
B / world series and the only data-path exchange is the sum of the scalar loss per MLL evaluation -- pushed to the peers
by the MLL kernel itself (LossExchange), an NCCL all-reduce on a side stream where peer memory is unavailable;
rollouts need none."""
import math
import os
import warnings
import weakref

import torch
import torch.nn.functional as F

from . import _lib, ops
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
    if os.environ.get("VOLT_LOSS_EXCHANGE") == "none":   # measurement only: every rank keeps its own partial (no exchange at all)
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


class _ExchangedLoss(PendingLoss):
    """Loss of a step whose partials were pushed to every rank by the MLL kernel itself (LossExchange)."""

    def __init__(self, exchange, seq):
        self._ex, self._seq, self._t, self._ev = exchange, seq, None, None

    def wait(self):
        if self._t is None:
            self._t = self._ex._total(self)
        if self._ev is not None:                       # summed by a later step's kernel: order after that launch
            self._ex._wait_event(self._ev)
            self._ev = None
        return self._t


class LossExchange:
    """The cross-rank sum of the step loss WITHOUT a collective launch: every rank owns a symmetric-memory
    buffer of RING x world 64-bit slots; the MLL kernel's last CTA stores {step number, partial} into slot
    [step % RING][rank] of every rank's buffer over NVLink and adds up the slots of step - LAG of its own buffer
    (volt_mll_step_sharded), so in a training loop the total of step s is simply there once step s + LAG has run; only the
    newest LAG steps need a tiny kernel that waits for their `world` slots (volt_loss_gather).  A separate NCCL kernel
    cannot overlap the next step here -- the persistent MLL kernel leaves it no SM to run on -- a peer store can.

    Slots and totals are reused after RING steps: `next()` copies out any loss still un-waited RING - LAG - 1 steps later."""

    RING = 8
    # The kernel of step s sums step s - LAG.  With 1 the ranks stay in lock step (a kernel cannot end before every rank has
    # finished the step before it); 2 would let them drift by a whole step.  Measured at 8 GPUs: lag 1
    # 0.962 of linear in both runs, lag 2 between 0.84 and 0.96 over four -- the in-kernel wait is 2 us per step either way,
    # so the default is the one with the steadier result.  VOLT_LOSS_LAG overrides it (same value on every rank).
    LAG = max(1, min(4, int(os.environ.get("VOLT_LOSS_LAG", "1"))))

    def __init__(self, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(device)
        self.slots = symm.empty(self.RING * self.world, dtype=torch.int64, device=self.device)
        self.slots.zero_()                       # step numbers start at 1: a zeroed slot never matches
        torch.cuda.synchronize(self.device)
        self.handle = symm.rendezvous(self.slots, group)   # collective; every rank has zeroed its slots when it returns
        self.handle.barrier()
        torch.cuda.synchronize(self.device)
        self.peers = int(self.handle.buffer_ptrs_dev)
        self.totals = torch.zeros(self.RING + 1, device=self.device)   # (+1: wait-time counter of a -DVOLT_EXCHANGE_DEBUG build)
        self.seq = 0
        # "kernel" (default): the step kernel's last CTA issues the remote stores itself.  "side": a one-warp kernel on a side
        # stream does (volt_loss_push) -- measured at 8 GPUs: no difference (1.72 vs 1.78 ms per step, inside the run-to-run spread)
        self.push_mode = os.environ.get("VOLT_LOSS_PUSH", "kernel")
        self.side = torch.cuda.Stream(device=self.device) if self.push_mode == "side" else None
        self._keep = []          # partial-loss buffers of the steps whose push may still be in flight
        self._pending = []       # weak references to the losses handed out, oldest first
        self._launched = None    # event recorded after the newest step's launch

    def next(self):
        """-> the `exchange` tuple for ops.mll_step for the next step, and the loss object to hand to the caller;
        call launched() right after the step has been enqueued."""
        self.seq += 1
        while self._pending and (self._pending[0]() is None or self._pending[0]()._seq <= self.seq - (self.RING - self.LAG - 1)):
            old = self._pending.pop(0)()
            if old is not None:
                old._t = old.wait().clone()      # its totals entry is about to be reused
        loss = _ExchangedLoss(self, self.seq)
        self._pending.append(weakref.ref(loss))
        return (self.peers if self.side is None else 0, self.slots.data_ptr(), self.totals if self.seq > self.LAG else None, self.LAG,
                self.world, self.rank, self.RING, self.seq), loss

    def _record_event(self):
        ev = torch.cuda.Event()
        ev.record()
        return ev

    def _wait_event(self, ev):
        torch.cuda.current_stream().wait_event(ev)

    def launched(self, partial=None):
        """The step of `next()` is enqueued on the current stream: the total of the step LAG before it is ordered behind it."""
        ev = self._record_event()
        if self.side is not None:                # publish this step's partial from the side stream
            self._keep = self._keep[-(self.RING - 1):] + [partial]
            self.side.wait_event(ev)
            _lib.check(_lib.load().volt_loss_push(partial.data_ptr(), self.peers, self.world, self.rank, self.RING,
                                                  self.seq & 0xFFFFFFFF, self.side.cuda_stream), "volt_loss_push")
        for ref in self._pending:
            l = ref()
            if l is not None and l._seq == self.seq - self.LAG and l._t is None:
                l._t, l._ev = self.totals[l._seq % self.RING], ev

    def _total(self, loss):
        """Newest step (no later kernel has summed it): gather on the current stream."""
        self._pending = [r for r in self._pending if r() is not None and r() is not loss]
        out = torch.empty(1, device=self.device)
        _lib.check(_lib.load().volt_loss_gather(self.slots.data_ptr(), self.world, self.RING, loss._seq & 0xFFFFFFFF,
                                                out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream),
                   "volt_loss_gather")
        return out.reshape(())


_exchange = {}   # device index -> LossExchange, or None once setting it up failed (NCCL all-reduce is used instead)


def _loss_exchange(device):
    """The peer-memory exchange for this device when the job is multi-rank and VOLT_LOSS_EXCHANGE is not "nccl"."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return None
    if os.environ.get("VOLT_LOSS_EXCHANGE", "peer") in ("nccl", "none") or dist.get_backend() != "nccl":
        return None
    idx = torch.device(device).index
    if idx not in _exchange:
        try:
            _exchange[idx] = LossExchange(device)
        except Exception as e:   # no peer access / symmetric memory unavailable: all ranks fail alike
            warnings.warn(f"volt_b200: peer-memory loss exchange unavailable ({e}); using an NCCL all-reduce on a side stream")
            _exchange[idx] = None
    return _exchange[idx]


def mll_and_grad(x, vol, resid, raw_noise, jitter=1e-6, check=False):
    """One MLL + gradient evaluation for each of the B local series (train_utils.py:247-250 per series): ONE launch
    (plus the prefix-sum kernel) -- the likelihood's softplus transform, dMLL/draw_noise and the rank-local partial of the
    loss are produced by the kernel's epilogue (volt_mll_grad_vol_raw); on more than one GPU the kernel also exchanges the
    partial with the other ranks over peer memory (LossExchange; NCCL all-reduce on a side stream as the fallback).

    x (T,), vol (B,T), resid (B,T) = log y - mean, raw_noise (B,).  Returns dict of CUDA tensors:
    mll (B,), draw_noise (B,) = dMLL/draw_noise, alpha (B,T) (dMLL/dmean = alpha/T), info (B,), scalars (B,16), and
    loss = -sum mll over ALL ranks as a PendingLoss (float(loss) / loss.wait())."""
    ex = _loss_exchange(resid.device if torch.is_tensor(resid) and resid.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    if ex is not None:
        desc, loss = ex.next()
        out = ops.mll_step("vol", x, vol, resid, raw_noise, jitter=jitter, check=False, exchange=desc)
        ex.launched(out["loss"])
        if check:
            ops._check_info(out["info"], out["scalars"][:, ops.S_JITTER], "exact MLL")
    else:
        out = ops.mll_step("vol", x, vol, resid, raw_noise, jitter=jitter, check=check)
        loss = _async_loss_all_reduce(out["loss"])
    sc = out["scalars"]
    return dict(mll=sc[:, S_MLL], draw_noise=sc[:, S_DRAW], alpha=out["alpha"], info=out["info"], loss=loss, scalars=sc,
                partial_loss=out["loss"])


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
