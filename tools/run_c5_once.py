import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from volt_b200 import batched, ops
T = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
x, vol, logy = batched.synth_series(1, T)
_, resid = ops.ma_mean("ewma", logy.cuda(), 25, want_resid=True)
raw = torch.full((1,), 1e-5).cuda()
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    o = batched.mll_and_grad(x.cuda(), vol.cuda(), resid, raw)
torch.cuda.synchronize()
print(float(o["mll"][0]))
