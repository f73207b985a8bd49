"""Series-in-flight probe of mll_batched_tc_kernel: time B = 148 / 296 / 444 / 888 / 1024 series of length T (one, two, three
resident CTAs per SM, then full waves) to separate per-series latency from per-SM throughput."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from volt_b200 import _lib, batched, ops  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lib = _lib.load()
dev = torch.device("cuda")
Bmax = 1024
x, vol, logy = batched.synth_series(Bmax, T)
xd, vd = x.to(dev), vol.to(dev)
_, resid = ops.ma_mean("ewma", logy.to(dev), 25, want_resid=True)
noise = batched.noise_from_raw(torch.full((Bmax,), 1e-5, device=dev))
scal = torch.empty(Bmax, 16, device=dev)
alpha = torch.empty(Bmax, T, device=dev)
info = torch.empty(Bmax, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 * 1024 * 1024, device=dev)
for B in (1, 74, 148, 296, 444, 592, 888, 1024):
    def run():
        _lib.check(lib.volt_mll_grad_vol(xd.data_ptr(), 0, vd.data_ptr(), 1, resid.data_ptr(), noise.data_ptr(), 1, B, T, 1e-6, 3,
                                         scal.data_ptr(), alpha.data_ptr(), info.data_ptr(), st), "mll")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"T={T} B={B:5d}: median {ts[len(ts) // 2]:.4f} ms  min {ts[0]:.4f} ms  -> {B / ts[len(ts) // 2]:.0f} series/ms")
