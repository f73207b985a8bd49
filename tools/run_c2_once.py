"""Two c2-sized MLL+grad launches (B x T from argv, default 1024 x 512) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from volt_b200 import batched, ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 512
x, vol, logy = batched.synth_series(B, T)
_, resid = ops.ma_mean("ewma", logy.cuda(), 25, want_resid=True)
raw = torch.full((B,), 1e-5).cuda()
for _ in range(3):
    out = batched.mll_and_grad(x.cuda(), vol.cuda(), resid, raw)
torch.cuda.synchronize()
print(float(out["loss"]))
