"""Timing of the evaluation reductions on the c4 per-GPU share (512 x 512 x 30) and a larger tensor (HBM-bound check)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from volt_b200 import ops

def ev_time(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]

out = {}
for name, B, S, H in (("c4_share", 512, 512, 30), ("big", 4096, 512, 30)):
    smp = torch.randn(B, S, H, device="cuda")
    tr = torch.zeros(B, H, device="cuda")
    ms = ev_time(lambda: ops.rollout_stats(smp, truth=tr, strike=tr))
    ms_t = ev_time(lambda: (smp.mean(1), smp.std(1), (smp < tr.unsqueeze(1)).float().mean(1)))
    out[name] = dict(B=B, S=S, H=H, ms=ms, GBs=4.0 * B * S * H / ms / 1e6, torch_eager_ms=ms_t)
print(json.dumps(out, indent=1))
