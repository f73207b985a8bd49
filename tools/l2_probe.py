"""Kernel time of one wave of series (one CTA per SM or fewer) at T=512: does the chain of one series depend on how many
scratch squares compete for the L2?  (B x 1 MB against 126 MB.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from volt_b200 import batched, ops  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 512
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
for B in (32, 64, 100, 120, 148, 200, 296):
    x, vol, logy = batched.synth_series(B, T)
    _, resid = ops.ma_mean("ewma", logy.cuda(), 25, want_resid=True)
    raw = torch.full((B,), 1e-5).cuda()
    xd, vd = x.cuda(), vol.cuda()
    for _ in range(3):
        batched.mll_and_grad(xd, vd, resid, raw)
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        batched.mll_and_grad(xd, vd, resid, raw)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"B={B:4d}  median {ts[len(ts) // 2]:.4f} ms  min {ts[0]:.4f} ms")
