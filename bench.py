#!/usr/bin/env python
"""bench.py -- MLL+grad evaluations / second on the BASELINE.json headline workload (config c2:
1024 synthetic stock-return series x T=512, exact batched MLL + gradients) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5                      # our arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus 1 --steps 20 --warmup 5     # reference arm: CPU oracle port on host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...               # one rank per GPU, series-sharded (weak scaling)

A "step" is one exact MLL + gradient evaluation for every series of the local batch (train_utils.py:247-250 per
series): cumtrapz -> ONE kernel (fused build + potrf + forward substitution + trtri -> tr A^-1, alpha; likelihood
transform, dMLL/draw_noise and the rank-local partial of the loss in its epilogue; at N > 1 the same kernel pushes that
partial into every rank's slot buffer over NVLink peer memory and sums the previous step's slots: no collective launch,
DESIGN.md section 4).  The residual y - EWMA_k(y) is computed once outside the timed region in BOTH arms (it does not depend on the
trained parameter; the reference recomputes it every iteration, 0.1 % of its step).
`value` is timed with inputs resident in HBM (CUDA events per step, L2 flushed between steps, max over ranks);
`e2e` is the same metric through the host-buffer C-ABI call on pinned host buffers (the kernel reads x / vol / resid /
noise from them over PCIe and writes the per-series results back, all inside the timed region; at N > 1 also an
all-reduce of the loss); at N > 1 `per_rank` lists every rank's own step, kernel and enqueue time.  Extra objects on the c2 line: `rollout` (c4 per-GPU
share), `long_series` (c5: one series of T = 8192), each with its own roofline and CPU baseline; `gpu_torch_baseline`
(stock torch on the same GPU: what the reference's `.cuda()` path runs); `peaks` (incl. a TF32 cuBLAS GEMM measured in
this run).  Other workloads (--workload c1|c3) are for profiling, not bench lines.
"""
import argparse
import json
import os
import platform
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (series per GPU, T, dt, description)
    "c1": (1, 256, 1.0 / 252, "c1: single synthetic GBM series T=256, Volatility kernel + EWMA(k=25) mean"),
    "c2": (1024, 512, 1.0 / 252, "c2: 1024 synthetic stock-return series x T=512, batched exact MLL+grad"),
    "c3": (256, 1024, 1.0 / 365, "c3: weather shape, 256 stations x T=1024, MA mean + Volatility kernel"),
}
K_EWMA = 25
RAW_NOISE = 1e-5  # train_utils.py:222 sets the RAW noise to 1e-5 (noise = softplus(1e-5) + 1e-4 ~ 0.6933)
L2_NOTE = "GPU arm: L2 flushed between timed iterations (256 MB write); CPU arm: working set (GBs) exceeds every cache"


def config_of(desc, B, T):
    """Identical in both arms (the driver compares them)."""
    return dict(workload=desc, series_per_gpu=B, T=T, mean=f"ewma k={K_EWMA} (residual precomputed outside the timed region)",
                raw_noise=RAW_NOISE, l2=L2_NOTE)


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or platform.machine()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf=p.get("bf16_tflops_sustained", p["bf16_tflops"]), tf_burst=p["bf16_tflops"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML every 100 ms while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if mask & bit:
                        self.reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
            self._stop_evt.wait(0.1)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(s))


# ---------------------------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_inputs(n_series, T, dt):
    """Cached train_cov (VoltMagpie.py:46: built once per series, not part of a step) and the precomputed residual."""
    from oracle import volt_oracle as O
    from volt_b200 import batched

    import torch

    x, vol, logy = batched.synth_series(n_series, T, dt)
    K = O.vol_kernel(x, vol)
    resid = logy - O.ewma(logy, K_EWMA)[..., :-1]
    raw = torch.full((n_series,), RAW_NOISE)
    return K, resid, raw


def cpu_mll_grad_step(K, resid, raw, threads, chunk=256):
    """The reference's CPU path for one MLL+grad step per series (cached train_cov, Cholesky MLL through
    psd_safe_cholesky, autograd backward) as restated by the oracle; batched over the series in chunks of `chunk` (bounds
    the host memory of the autograd graph).  Returns seconds."""
    import torch

    from oracle import volt_oracle as O

    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    for lo in range(0, K.shape[0], chunk):
        r = raw[lo:lo + chunk].clone().requires_grad_(True)
        mll = O.exact_mll(K[lo:lo + chunk], resid[lo:lo + chunk], O.noise_from_raw(r))
        (-mll.sum()).backward()
    return time.perf_counter() - t0


# ---------------------------------------------------------------------------------------------- GPU comparators
def measure_tf32_peak(dev, n=8192):
    """cuBLAS TF32 GEMM (fp32 in, TF32 tensor-core math) on this GPU, in this run: the tensor-pipe peak for kind::tf32."""
    import torch

    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        best = float("inf")
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 60
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        sustained = e0.elapsed_time(e1) / reps
        del a, b
        fl = 2.0 * n ** 3
        return dict(burst=fl / (best * 1e-3) / 1e12, sustained=fl / (sustained * 1e-3) / 1e12,
                    how=f"torch.matmul fp32 {n}^3 with allow_tf32 (cuBLAS), best of 10 / {reps} back to back, this run")
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def gpu_torch_baseline(xd, vd, resid, raw, T, reps=3):
    """Stock torch on the same GPU -- what the reference runs when its drivers move the tensors with .cuda()
    (GenerateMultiMeanPreds.py:92-96): dense cached K, torch.linalg.cholesky (cuSOLVER / MAGMA batched potrf),
    triangular solve, log-det, autograd backward.  Same inputs, same step definition; ms per step (best of reps)."""
    import math

    import torch

    B = vd.shape[0]
    dx = xd[1] - xd[0]
    w = dx * torch.ones_like(xd)
    w[0] *= 0.5
    w[-1] *= 0.5
    V = torch.cumsum(w * vd * vd, -1)
    idx = torch.arange(T, device=xd.device)
    K = V[:, torch.minimum(idx[:, None], idx[None, :])]        # cached train_cov (VoltMagpie.py:46)
    eye = torch.eye(T, device=xd.device)
    best = float("inf")
    for it in range(reps + 1):
        r = raw.clone().requires_grad_(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        noise = torch.nn.functional.softplus(r) + 1e-4
        A = K + noise[:, None, None] * eye
        L = torch.linalg.cholesky(A)
        z = torch.linalg.solve_triangular(L, resid.unsqueeze(-1), upper=False).squeeze(-1)
        mll = -0.5 * ((z * z).sum(-1) + 2.0 * torch.diagonal(L, dim1=-2, dim2=-1).log().sum(-1) + T * math.log(2 * math.pi)) / T
        (-mll.sum()).backward()
        e1.record()
        torch.cuda.synchronize()
        if it > 0:
            best = min(best, e0.elapsed_time(e1))
        loss = float(-mll.sum())
        del A, L, z, mll, r
    del K
    torch.cuda.empty_cache()
    return dict(ms_per_step=best, value=B / (best * 1e-3), unit="evals/s", loss=loss,
                what="stock torch on this GPU (reference's .cuda() path): cached dense K, torch.linalg.cholesky + "
                     "solve_triangular + autograd backward, fp32, best of %d" % reps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=0, help="series per CPU step (0 = the whole per-GPU batch on the reference arm, 64 for the in-line cpu_baseline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rollout", action="store_true", help="skip the secondary rollout-throughput measurement")
    ap.add_argument("--no-long", action="store_true", help="skip the secondary long-series (c5) measurement")
    ap.add_argument("--no-torch-baseline", action="store_true", help="skip the stock-torch-on-GPU comparator")
    args = ap.parse_args()

    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    B, T, dt, desc = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm (CPU oracle port, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        n = args.cpu_sample if args.cpu_sample > 0 else B
        K, resid, raw = cpu_inputs(n, T, dt)
        for _ in range(max(args.warmup, 1)):
            cpu_mll_grad_step(K, resid, raw, threads)
        times = [cpu_mll_grad_step(K, resid, raw, threads) for _ in range(args.steps)]
        el = sum(times)
        value = n * args.steps / el
        sample = (f"{n} of {B} series x T={T} per step, batched torch CPU (chunks of 256), {threads} threads on {cpu_model()}; "
                  f"mean of {args.steps} steps (best step {n / min(times):.0f} evals/s, worst {n / max(times):.0f})")
        line = dict(impl="reference", metric="MLL+grad evals/sec", value=value, unit="evals/s", n_gpus=args.gpus,
                    steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * el / args.steps, higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", config=config_of(desc, B, T),
                    series_per_step=n, cpu_model=cpu_model(),
                    cpu_baseline=dict(value=value, unit="evals/s", cores=threads, kind="port", sample=sample,
                                      best_step_value=n / min(times)),
                    e2e=dict(value=value, unit="evals/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ our arm
    import torch.distributed as dist

    from volt_b200 import _lib, batched, ops

    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL announces its version on stdout when the first communicator is created; stdout must carry exactly one JSON
        # line, so file descriptor 1 points at stderr while the process group comes up (init + one collective)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.all_reduce(torch.zeros(1, device=torch.device("cuda", local_rank)))
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    _lib.require_device()
    dev = torch.device("cuda", local_rank)

    # weak scaling: every rank owns B series; series b of rank r is global series r*B + b
    x, vol, logy = batched.synth_series(B, T, dt, start=rank * B)
    xd, vd, yd = x.to(dev), vol.to(dev), logy.to(dev)
    _, resid = ops.ma_mean("ewma", yd, K_EWMA, want_resid=True)     # outside the timed region in both arms
    raw = torch.full((B,), RAW_NOISE, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)  # > 126 MB L2

    def step():
        return batched.mll_and_grad(xd, vd, resid, raw)

    for _ in range(max(args.warmup, 3)):
        out = step()
    float(out["loss"])
    torch.cuda.synchronize()
    assert int(out["info"].abs().sum()) == 0, "Cholesky failure on the synthetic workload"

    lib = _lib.load()
    noise = batched.noise_from_raw(raw)
    scal = torch.empty(B, 16, device=dev)
    alpha = torch.empty(B, T, device=dev)
    info = torch.empty(B, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # K timed steps.  The cross-rank sum of step s completes under step s + 1 (pushed to the peers by the kernel of step s
    # and added up at the end of the kernel of step s + 1; with VOLT_LOSS_EXCHANGE=nccl an all-reduce on a side stream);
    # the end event of a step is recorded after the current stream has waited for the total of the step `lag` before it (the
    # last step waits for everything still pending), so every exchange completes inside a timed region.
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = _lib.launch_count()
    lag = batched.LossExchange.LAG if world > 1 else 1
    pending = []
    host_t0 = time.perf_counter()
    for s in range(args.steps):
        flush.zero_()  # evict L2 between timed iterations (not timed)
        ev[s][0].record()
        out = step()
        pending.append(out["loss"])
        if len(pending) > lag:
            pending.pop(0).wait()
        if s == args.steps - 1:
            for l in pending:
                l.wait()
        ev[s][1].record()
    host_ms = (time.perf_counter() - host_t0) * 1e3 / args.steps     # enqueue time per step (the host runs ahead of the GPU)
    torch.cuda.synchronize()
    untimed_ms = sum(ev[s - 1][1].elapsed_time(ev[s][0]) for s in range(1, args.steps)) / max(args.steps - 1, 1)   # the L2 flush
    ex_dbg = batched._exchange.get(local_rank)
    wait_us = float(ex_dbg.totals[-1]) / (args.steps + max(args.warmup, 3)) if ex_dbg is not None else 0.0   # debug builds only
    launches = _lib.launch_count() - l0
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    loss = float(out["loss"])

    # dominant-kernel duration, measured live (same stream, CUDA events, L2 flushed): the fused-step kernel itself
    loss_dev = torch.empty(1, device=dev)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s in range(args.steps):
        flush.zero_()
        kev[s][0].record()
        _lib.check(lib.volt_mll_grad_vol_raw(xd.data_ptr(), 0, vd.data_ptr(), 1, resid.data_ptr(), raw.data_ptr(), 1, B, T, 1e-6, 3,
                                             scal.data_ptr(), alpha.data_ptr(), info.data_ptr(), loss_dev.data_ptr(), st),
                   "volt_mll_grad_vol_raw")
        kev[s][1].record()
    torch.cuda.synchronize()
    kern_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps

    # end-to-end through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region); at N > 1 the
    # step also all-reduces the loss (host scalar -> device -> NCCL -> host), like the device-resident step
    hx, hv, hr = x.pin_memory(), vol.pin_memory(), resid.cpu().pin_memory()
    hn = noise.cpu().pin_memory()
    hs = torch.empty(B, 16).pin_memory()
    hi = torch.empty(B, dtype=torch.int32).pin_memory()
    e2e_loss = torch.zeros(1, device=dev)

    e2e_direct = os.environ.get("VOLT_E2E_DIRECT", "1") != "0"

    def e2e_step():
        _lib.check(lib.volt_mll_grad_vol_host(hx.data_ptr(), hv.data_ptr(), hr.data_ptr(), hn.data_ptr(), 1, B, T, 1e-6, 3,
                                              hs.data_ptr(), None, hi.data_ptr()), "volt_mll_grad_vol_host")
        if world > 1:
            e2e_loss.fill_(-float(hs[:, 0].sum()))
            dist.all_reduce(e2e_loss)
            return float(e2e_loss)
        return -float(hs[:, 0].sum())

    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    e2e_t = 0.0
    for s in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_total = e2e_step()
        e2e_t += time.perf_counter() - t0
    clocks = sampler.finish()
    if os.environ.get("VOLT_LOSS_EXCHANGE") != "none":   # ("none": measurement-only mode, every rank keeps its own partial)
        assert abs(e2e_total - loss) < 1e-2 * abs(loss) + 1e-3, (e2e_total, loss)

    pk = peaks()
    tf32 = measure_tf32_peak(dev) if rank == 0 else None

    # secondary metric of BASELINE.json's north_star: Monte-Carlo rollout throughput on the c4 per-GPU share
    # (512 series x 512 draws x 30 steps, T=256, EWMA k=25, Philox normals in-kernel); reported, not the headline
    roll = None
    rB, rT, rS, rH = 512, 256, 512, 30
    if args.workload == "c2" and not args.no_rollout:
        rx, rvol, rlogy = batched.synth_series(rB, rT, dt, start=rank * rB)
        g = torch.Generator().manual_seed(1 + rank)
        rpv = (rvol[:, -1:, None] * torch.exp(0.1 * torch.randn(rB, rS, rH, generator=g))).to(dev)
        rxd, rvd, ryd = rx.to(dev), rvol.to(dev), rlogy.to(dev)
        for _ in range(3):
            ops.rollout(rxd, ryd, rvd, rpv, eps=None, k=K_EWMA, seed=3, check=False)
        torch.cuda.synchronize()
        rev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, b in rev:
            flush.zero_()
            a.record()
            ops.rollout(rxd, ryd, rvd, rpv, eps=None, k=K_EWMA, seed=3, check=False)
            b.record()
        torch.cuda.synchronize()
        roll_ms = sum(a.elapsed_time(b) for a, b in rev) / len(rev)
    else:
        roll_ms = 0.0

    # secondary: config c5 of BASELINE.json -- ONE series of T = 8192 through the multi-CTA long-series path (the
    # tensor-pipe roofline point); reported, never allowed to break the headline line
    long_ms = 0.0
    lT = 8192
    if args.workload == "c2" and not args.no_long:
        try:
            lx, lvol, llogy = batched.synth_series(1, lT, dt, start=rank)
            _, lres = ops.ma_mean("ewma", llogy.to(dev), K_EWMA, want_resid=True)
            lnoise = batched.noise_from_raw(torch.full((1,), RAW_NOISE, device=dev))
            lxd, lvd = lx.to(dev), lvol.to(dev)
            for _ in range(2):   # rank-local calls only (no collective inside a try block)
                lout = ops.mll_grad("vol", lxd, lvd, lres, lnoise, check=False)
            torch.cuda.synchronize()
            lev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
            for a, b in lev:
                a.record()
                lout = ops.mll_grad("vol", lxd, lvd, lres, lnoise, check=False)
                b.record()
            torch.cuda.synchronize()
            assert int(lout["info"].abs().sum()) == 0
            long_ms = sum(a.elapsed_time(b) for a, b in lev) / len(lev)
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] long-series measurement skipped: {exc}", file=sys.stderr)
            long_ms = 0.0

    # stock torch on the same GPU (SURVEY 2.4's second bar), rank 0 at N = 1 only
    torch_base = None
    if rank == 0 and world == 1 and not args.no_torch_baseline:
        try:
            torch_base = gpu_torch_baseline(xd, vd, resid, raw, T)
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] stock-torch GPU baseline skipped: {exc}", file=sys.stderr)

    t = torch.tensor([total_ms, e2e_t * 1e3, kern_ms, roll_ms, host_ms, untimed_ms, wait_us], device=dev, dtype=torch.float64)
    per_rank = None
    if world > 1:
        # every rank's own numbers next to the max: the spread between GPUs of one box is what weak scaling loses here
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = dict(ms_per_step=[round(float(a[0]) / args.steps, 4) for a in allt],
                        kernel_ms=[round(float(a[2]), 4) for a in allt],
                        host_enqueue_ms_per_step=[round(float(a[4]), 4) for a in allt],
                        untimed_flush_ms_per_step=[round(float(a[5]), 4) for a in allt],
                        exchange_wait_us_per_step=[round(float(a[6]), 1) for a in allt])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kern_ms, roll_ms = (float(v) for v in t[:4])
    value = B * world * args.steps / (total_ms * 1e-3)
    e2e_value = B * world * args.steps / (e2e_ms * 1e-3)

    traffic, traffic_src = None, None  # measured DRAM bytes per launch of the dominant kernel (one ncu --set full capture, committed)
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f).get(args.workload)
        if tj:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            traffic_src = tj.get("source", "profiles/traffic.json (one ncu --set full capture of this kernel)")
    # algorithmic work per eval: bytes 4 T^2 + 12 T (build fused into the factorisation),
    # flops 2 T^3 / 3 (potrf + trtri) + 4 T^2
    bytes_per_eval = 4.0 * T * T + 12.0 * T
    flops_per_eval = 2.0 * T ** 3 / 3.0 + 4.0 * T * T
    ach_gbs = bytes_per_eval * B / (kern_ms * 1e-3) / 1e9
    ach_tf = flops_per_eval * B / (kern_ms * 1e-3) / 1e12
    tf32_peak = tf32["sustained"] if tf32 else None
    line = dict(
        metric="MLL+grad evals/sec", value=value, unit="evals/s", n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
        ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
        data="synthetic", config=config_of(desc, B, T), loss=loss,
        # direct form of the host entry (pinned buffers, default): every series reads x, vol and resid from mapped host memory
        e2e=dict(value=e2e_value, unit="evals/s",
                 h2d_bytes_per_step=int(((3 * B * T + B) if e2e_direct else (T + 2 * B * T + B)) * 4 + (4 if world > 1 else 0)),
                 d2h_bytes_per_step=int(B * 16 * 4 + B * 4 + (4 if world > 1 else 0)),
                 transfer=("kernel reads / writes the caller's pinned host buffers (mapped) over PCIe" if e2e_direct
                           else "staged: cudaMemcpyAsync H2D under the kernel + D2H"),
                 collective=("loss all-reduce inside the timed step" if world > 1 else "none (one rank)")),
        gpu_launches=int(launches),
        **({"per_rank": per_rank, "loss_exchange": type(out["loss"]).__name__} if world > 1 else {}),
        clocks=clocks,
        peaks=dict(hbm_gbs=pk["hbm"], bf16_tflops_sustained=pk["tf"], source=pk["source"], tf32_tflops=tf32),
        roofline=dict(bound="hbm", achieved=ach_gbs, peak=pk["hbm"], unit="GB/s", frac=ach_gbs / pk["hbm"], traffic=traffic,
                      traffic_source=traffic_src,
                      kernel="mll_batched_tc_kernel (+ cumtrapz_kernel, <1% of the time)", ms_per_launch=kern_ms,
                      bytes_per_eval=bytes_per_eval, peak_source=pk["source"]),
        roofline_tensor=dict(bound="tensor", achieved=ach_tf, peak=tf32_peak, unit="TFLOP/s",
                             frac=(ach_tf / tf32_peak if tf32_peak else None), issued_frac=(3.0 * ach_tf / tf32_peak if tf32_peak else None),
                             flops_per_eval=flops_per_eval,
                             note="peak = the TF32 cuBLAS GEMM measured in this run (sustained); algorithmic flops (potrf + trtri); "
                                  "the tcgen05 kind::tf32 3-pass split (hi*hi + hi*lo + lo*hi) issues 3x of them (issued_frac)"),
        gpu_torch_baseline=torch_base,
    )
    if rank == 0:
        from oracle import volt_oracle as O   # CPU baselines only (the checker / comparator, never the product path)

        want_cpu = world == 1 and not args.no_cpu_baseline
        if roll_ms > 0:
            n_ss = rB * rS * rH
            # 12 B per step-sample (pred_vol in, base normal in, sample out; the normals are generated in-kernel
            # here: 8 B) + the per-series shared factor (4 T^2: build fused, written once to scratch)
            r_bytes = 8.0 * n_ss + rB * 4.0 * rT * rT
            r_gbs = r_bytes / (roll_ms * 1e-3) / 1e9
            roll = dict(metric="rollout step-samples/sec", value=n_ss * world / (roll_ms * 1e-3), unit="step-samples/s",
                        ms_per_call=roll_ms, path_samples_per_s=rB * rS * world / (roll_ms * 1e-3),
                        config="c4 per-GPU share: 512 series x 512 draws x 30 steps, T=256, ewma k=25, Philox in-kernel, L2 flushed",
                        roofline=dict(bound="hbm", achieved=r_gbs, peak=pk["hbm"], unit="GB/s", frac=r_gbs / pk["hbm"],
                                      bytes_per_call=r_bytes, stream_only_gbs=8.0 * n_ss / (roll_ms * 1e-3) / 1e9,
                                      note="algorithmic bytes: 8 B per step-sample (pred_vol in, sample out, normals in-kernel) "
                                           "+ 4 T^2 per series for the shared factor; whole call (prep + rollout kernel)"))
            if want_cpu:
                cS, cB = 16, 2
                cx, cvol, clogy = batched.synth_series(cB, rT, dt)
                g = torch.Generator().manual_seed(1)
                cpv = cvol[:, -1:, None] * torch.exp(0.1 * torch.randn(cB, cS, rH, generator=g))
                ceps = torch.randn(cB, cS, rH, generator=g)
                ctx = cx[-1] + cx[1] * torch.arange(1, rH + 1)
                torch.set_num_threads(threads)
                t0 = time.perf_counter()
                for b in range(cB):
                    px = torch.cat((clogy[b, :1], clogy[b])).exp()
                    O.rollouts(cx, px, cvol[b].log(), ctx, cpv[b], ceps[b], K_EWMA)
                cel = time.perf_counter() - t0
                roll["cpu_baseline"] = dict(value=cB * cS * rH / cel, unit="step-samples/s", cores=threads, kind="port",
                                            sample=f"{cB} series x {cS} draws x {rH} steps, T={rT} (re-factorisation per step, "
                                                   f"rollout_utils.py:35), {cel:.1f} s on {cpu_model()}")
        line["rollout"] = roll
        long_obj = None
        if long_ms > 0:
            lfl = 2.0 * lT ** 3 / 3.0 + 4.0 * lT ** 2
            l_tf = lfl / (long_ms * 1e-3) / 1e12
            long_obj = dict(metric="MLL+grad evals/sec, one series of T=8192 (config c5, multi-CTA path)", ms_per_eval=long_ms,
                            value=1e3 / long_ms, unit="evals/s", n_gpus=1,
                            roofline=dict(bound="tensor", achieved=l_tf, peak=tf32_peak, unit="TFLOP/s",
                                          frac=(l_tf / tf32_peak if tf32_peak else None),
                                          issued_frac=(3.0 * l_tf / tf32_peak if tf32_peak else None),
                                          note="algorithmic flops (potrf + trtri); 3x are issued (3xTF32); peak = TF32 cuBLAS GEMM measured in this run"))
            if want_cpu:
                lx, lvol, llogy = batched.synth_series(1, lT, dt)
                lK = O.vol_kernel(lx, lvol)
                lres = llogy - O.ewma(llogy, K_EWMA)[..., :-1]
                lraw = torch.full((1,), RAW_NOISE)
                cpu_mll_grad_step(lK, lres, lraw, threads)
                cel = min(cpu_mll_grad_step(lK, lres, lraw, threads) for _ in range(2))
                long_obj["cpu_baseline"] = dict(value=1.0 / cel, unit="evals/s", cores=threads, kind="port",
                                                sample=f"the same series, torch CPU (Cholesky MLL + autograd backward), best of 2 "
                                                       f"({cel:.2f} s) on {cpu_model()}")
        line["long_series"] = long_obj
        if want_cpu:
            n = args.cpu_sample if args.cpu_sample > 0 else 64
            cK, cres, craw = cpu_inputs(n, T, dt)
            cpu_mll_grad_step(cK, cres, craw, threads)
            times = [cpu_mll_grad_step(cK, cres, craw, threads) for _ in range(5)]
            line["cpu_baseline"] = dict(value=n / min(times), unit="evals/s", cores=threads, kind="port", cpu_model=cpu_model(),
                                        sample=f"{n} of {B} series x T={T}, batched torch CPU, best of 5 ({min(times) * 1e3:.0f} ms per "
                                               f"pass; mean {n * 5 / sum(times):.0f} evals/s)")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
