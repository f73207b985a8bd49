#!/usr/bin/env python
"""bench.py -- MLL+grad evaluations / second on the BASELINE.json headline workload (config c2:
1024 synthetic stock-return series x T=512, exact batched MLL + gradients) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5                      # our arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1      # reference arm: CPU oracle port on host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...               # one rank per GPU, series-sharded (weak scaling)

A "step" is one exact MLL + gradient evaluation for every series of the local batch (train_utils.py:247-250 per
series): cumtrapz -> fused build + potrf + forward substitution + trtri (tr A^-1, alpha) -> scalar loss all-reduce.
`value` is timed with inputs resident in HBM (CUDA events per step, L2 flushed between steps, max over ranks);
`e2e` is the same metric through the host-buffer C-ABI call (H2D of x / vol / resid / noise and D2H of the per-series
results inside the timed region).  Two secondary objects ride on the c2 line: `rollout` (c4 per-GPU share) and
`long_series` (c5: one series of T = 8192, tensor-pipe roofline point).  Other workloads (--workload c1|c3) are for
profiling, not bench lines.
"""
import argparse
import json
import math
import os
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


def cpu_train_cov(x, vol, sample):
    """VoltMagpie.py:46: the reference builds train_cov once per series and caches it (not part of a step)."""
    from oracle import volt_oracle as O

    return O.vol_kernel(x, vol[:sample])


def cpu_mll_grad_arm(K, logy, raw, sample, reps, threads):
    """The reference's CPU path for one MLL+grad step per series (cached train_cov, EWMA mean recomputed every
    iteration, Cholesky MLL, autograd backward) as restated by the oracle; batched over `sample` series."""
    import torch

    from oracle import volt_oracle as O

    torch.set_num_threads(threads)
    ys = logy[:sample]
    best = float("inf")
    for _ in range(reps):
        r = raw[:sample].clone().requires_grad_(True)
        t0 = time.perf_counter()
        mean = O.ewma(ys, K_EWMA)[..., :-1]
        mll = O.exact_mll(K, ys - mean, O.noise_from_raw(r))
        (-mll.sum()).backward()
        best = min(best, time.perf_counter() - t0)
    return sample / best, best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=32, help="series timed on the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rollout", action="store_true", help="skip the secondary rollout-throughput measurement")
    ap.add_argument("--no-long", action="store_true", help="skip the secondary long-series (c5) measurement")
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
        from volt_b200 import batched

        x, vol, logy = batched.synth_series(args.cpu_sample, T, dt)
        raw = torch.full((args.cpu_sample,), RAW_NOISE)
        K = cpu_train_cov(x, vol, args.cpu_sample)
        for _ in range(max(args.warmup, 2)):
            cpu_mll_grad_arm(K, logy, raw, args.cpu_sample, 1, threads)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_mll_grad_arm(K, logy, raw, args.cpu_sample, 1, threads)
        el = time.perf_counter() - t0
        value = args.cpu_sample * args.steps / el
        sample = f"{args.cpu_sample} series x T={T} per step (of {B} per GPU), batched torch CPU"
        line = dict(impl="reference", metric="MLL+grad evals/sec", value=value, unit="evals/s", n_gpus=args.gpus,
                    steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * el / args.steps, higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=desc, series_per_step=args.cpu_sample, T=T, mean="ewma k=25", raw_noise=RAW_NOISE),
                    cpu_baseline=dict(value=value, unit="evals/s", cores=threads, kind="port", sample=sample),
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
    _, resid = ops.ma_mean("ewma", yd, K_EWMA, want_resid=True)
    raw = torch.full((B,), RAW_NOISE, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)  # > 126 MB L2

    def step():
        return batched.mll_and_grad(xd, vd, resid, raw)

    for _ in range(max(args.warmup, 3)):
        out = step()
    torch.cuda.synchronize()
    assert int(out["info"].abs().sum()) == 0, "Cholesky failure on the synthetic workload"

    # kernel-only timing of the dominant kernel (mll_batched_kernel) through the device-pointer C-ABI entry
    noise = batched.noise_from_raw(raw)
    lib = _lib.load()
    scal = torch.empty(B, 16, device=dev)
    alpha = torch.empty(B, T, device=dev)
    info = torch.empty(B, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = _lib.launch_count()
    for s in range(args.steps):
        flush.zero_()  # evict L2 between timed iterations (not timed)
        ev[s][0].record()
        out = step()
        ev[s][1].record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - l0
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    loss = float(out["loss"])

    # dominant-kernel duration, measured live (same stream, CUDA events, L2 flushed)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s in range(args.steps):
        flush.zero_()
        kev[s][0].record()
        _lib.check(lib.volt_mll_grad_vol(xd.data_ptr(), 0, vd.data_ptr(), 1, resid.data_ptr(), noise.data_ptr(), 1, B, T, 1e-6, 3,
                                         scal.data_ptr(), alpha.data_ptr(), info.data_ptr(), st), "volt_mll_grad_vol")
        kev[s][1].record()
    torch.cuda.synchronize()
    kern_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps

    # end-to-end through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region)
    hx, hv, hr = x.pin_memory(), vol.pin_memory(), resid.cpu().pin_memory()
    hn = noise.cpu().pin_memory()
    hs = torch.empty(B, 16).pin_memory()
    hi = torch.empty(B, dtype=torch.int32).pin_memory()

    def e2e_step():
        _lib.check(lib.volt_mll_grad_vol_host(hx.data_ptr(), hv.data_ptr(), hr.data_ptr(), hn.data_ptr(), 1, B, T, 1e-6, 3,
                                              hs.data_ptr(), None, hi.data_ptr()), "volt_mll_grad_vol_host")

    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    e2e_t = 0.0
    for s in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_step()
        e2e_t += time.perf_counter() - t0
    clocks = sampler.finish()
    assert abs(float(hs[:, 0].sum()) + loss) < 1e-2 * abs(loss) + 1e-3 or world > 1

    # secondary metric of BASELINE.json's north_star: Monte-Carlo rollout throughput on the c4 per-GPU share
    # (512 series x 512 draws x 30 steps, T=256, EWMA k=25, Philox normals in-kernel); reported, not the headline
    roll = None
    if args.workload == "c2" and not args.no_rollout:
        rB, rT, rS, rH = 512, 256, 512, 30
        rx, rvol, rlogy = batched.synth_series(rB, rT, dt, start=rank * rB)
        g = torch.Generator().manual_seed(1 + rank)
        rpv = (rvol[:, -1:, None] * torch.exp(0.1 * torch.randn(rB, rS, rH, generator=g))).to(dev)
        rxd, rvd, ryd = rx.to(dev), rvol.to(dev), rlogy.to(dev)
        for _ in range(3):
            ops.rollout(rxd, ryd, rvd, rpv, eps=None, k=K_EWMA, seed=3, check=False)
        torch.cuda.synchronize()
        rev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, b in rev:
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
    if args.workload == "c2" and not args.no_long:
        try:
            lT = 8192
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

    t = torch.tensor([total_ms, e2e_t * 1e3, kern_ms, roll_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kern_ms, roll_ms = (float(v) for v in t)
    if roll_ms > 0:
        roll = dict(metric="rollout step-samples/sec", value=512 * 512 * 30 * world / (roll_ms * 1e-3), unit="step-samples/s",
                    ms_per_call=roll_ms, config="c4 per-GPU share: 512 series x 512 draws x 30 steps, T=256, ewma k=25, Philox in-kernel",
                    path_samples_per_s=512 * 512 * world / (roll_ms * 1e-3))
    value = B * world * args.steps / (total_ms * 1e-3)
    e2e_value = B * world * args.steps / (e2e_ms * 1e-3)

    pk = peaks()
    traffic = None  # measured DRAM bytes per launch of the dominant kernel (one ncu --set full capture, committed)
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f).get(args.workload)
        if tj:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    # algorithmic work per eval: bytes 4 T^2 + 12 T (build fused into the factorisation),
    # flops 2 T^3 / 3 (potrf + trtri) + 4 T^2
    bytes_per_eval = 4.0 * T * T + 12.0 * T
    flops_per_eval = 2.0 * T ** 3 / 3.0 + 4.0 * T * T
    ach_gbs = bytes_per_eval * B / (kern_ms * 1e-3) / 1e9
    ach_tf = flops_per_eval * B / (kern_ms * 1e-3) / 1e12
    line = dict(
        metric="MLL+grad evals/sec", value=value, unit="evals/s", n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
        ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
        data="synthetic",
        config=dict(workload=desc, series_per_gpu=B, T=T, mean=f"ewma k={K_EWMA}", raw_noise=RAW_NOISE,
                    l2="flushed between timed iterations (256 MB write)", loss=loss),
        e2e=dict(value=e2e_value, unit="evals/s", h2d_bytes_per_step=int((T + 2 * B * T + B) * 4),
                 d2h_bytes_per_step=int(B * 16 * 4 + B * 4)),
        gpu_launches=int(launches),
        clocks=clocks,
        rollout=roll,
        long_series=(dict(metric="MLL+grad evals/sec, one series of T=8192 (config c5, multi-CTA path)", ms_per_eval=long_ms,
                          value=1e3 / long_ms, unit="evals/s", n_gpus=1,
                          roofline=dict(bound="tensor", achieved=(2.0 * 8192 ** 3 / 3.0 + 4.0 * 8192 ** 2) / (long_ms * 1e-3) / 1e12,
                                        peak=peaks()["tf"], unit="TFLOP/s",
                                        frac=(2.0 * 8192 ** 3 / 3.0 + 4.0 * 8192 ** 2) / (long_ms * 1e-3) / 1e12 / peaks()["tf"],
                                        note="algorithmic flops (potrf + trtri); 3x are issued (3xTF32); peak = measured bf16 dense GEMM"))
                     if long_ms > 0 else None),
        roofline=dict(bound="hbm", achieved=ach_gbs, peak=pk["hbm"], unit="GB/s", frac=ach_gbs / pk["hbm"], traffic=traffic,
                      kernel="mll_batched_tc_kernel (+ cumtrapz_kernel, <1% of the time)", ms_per_launch=kern_ms, bytes_per_eval=bytes_per_eval, peak_source=pk["source"]),
        roofline_tensor=dict(bound="tensor", achieved=ach_tf, peak=pk["tf"], unit="TFLOP/s", frac=ach_tf / pk["tf"],
                             flops_per_eval=flops_per_eval, note="tcgen05 kind::tf32 3-pass split (hi*hi + hi*lo + lo*hi): 3x the algorithmic flops are issued on the tensor pipe; peak is the measured bf16 dense GEMM figure (nominal TF32 peak is half of bf16)"),
    )
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            cx, cv, cy = batched.synth_series(args.cpu_sample, T, dt)
            craw = torch.full((args.cpu_sample,), RAW_NOISE)
            cK = cpu_train_cov(cx, cv, args.cpu_sample)
            cpu_mll_grad_arm(cK, cy, craw, args.cpu_sample, 2, threads)
            v, best = cpu_mll_grad_arm(cK, cy, craw, args.cpu_sample, 5, threads)
            line["cpu_baseline"] = dict(value=v, unit="evals/s", cores=threads, kind="port",
                                        sample=f"{args.cpu_sample} of {B} series x T={T}, batched torch CPU, best of 5 "
                                               f"({best * 1e3:.0f} ms per pass)")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
