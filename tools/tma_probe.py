"""A/B of the TMA-fed batched MLL instance against the register-staged ones: parity (vs the default instance and the fp64
oracle) and kernel time.  Usage (GPU box):  VOLT_TC_TMA=2 python tools/tma_probe.py check|time [B T]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import volt_oracle as O  # noqa: E402
from volt_b200 import batched, ops  # noqa: E402


def run(B, T, raw_val=1e-5):
    x, vol, logy = batched.synth_series(B, T)
    _, resid = ops.ma_mean("ewma", logy.cuda(), 25, want_resid=True)
    raw = torch.full((B,), raw_val).cuda()
    return x, vol, resid, raw, batched.mll_and_grad(x.cuda(), vol.cuda(), resid, raw, check=False)


if sys.argv[1] == "check":
    for B, T in [(3, 64), (2, 100), (4, 192), (5, 256), (3, 512), (2, 576), (2, 900), (2, 1024), (300, 128), (600, 512)]:
        x, vol, resid, raw, out = run(B, T)
        torch.cuda.synchronize()
        worst = 0.0
        for b in (0, B - 1):
            ref = O.volt_mll_and_grad(x.double(), vol[b].double(), resid[b].cpu().double(), raw[b].cpu().double())
            e1 = abs(float(out["mll"][b]) - float(ref["mll"])) / abs(float(ref["mll"]))
            e2 = abs(float(out["draw_noise"][b]) - float(ref["draw_noise"])) / abs(float(ref["draw_noise"]))
            e3 = float((out["alpha"][b].cpu().double() - ref["alpha"]).abs().max() / ref["alpha"].abs().max())
            worst = max(worst, e1, e2, e3)
            assert e1 < 1e-4 and e2 < 2e-3 and e3 < 2e-3, (B, T, b, e1, e2, e3)
        print(f"B={B} T={T} ok  worst rel err {worst:.2e}  info {int(out['info'].abs().sum())}", flush=True)
else:
    B, T = int(sys.argv[2]), int(sys.argv[3])
    x, vol, resid, raw, out = run(B, T)
    xd, vd = x.cuda(), vol.cuda()
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    for _ in range(3):
        batched.mll_and_grad(xd, vd, resid, raw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        batched.mll_and_grad(xd, vd, resid, raw)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"VOLT_TC_TMA={os.environ.get('VOLT_TC_TMA')} VOLT_TC_CTAS={os.environ.get('VOLT_TC_CTAS')} B={B} T={T}: "
          f"{sum(ts) / len(ts):.4f} ms (min {min(ts):.4f})", flush=True)
