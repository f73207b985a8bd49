import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from volt_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libvolt_prof.so")
from volt_b200 import batched, ops
B, T = 1024, 512
x, vol, logy = batched.synth_series(B, T)
_, resid = ops.ma_mean("ewma", logy.cuda(), 25, want_resid=True)
raw = torch.full((B,), 1e-5).cuda()
for _ in range(3): out = batched.mll_and_grad(x.cuda(), vol.cuda(), resid, raw)
torch.cuda.synchronize()
a = out["alpha"].reshape(-1)[:12].cpu()
names = ["non-diag", "S0 potrf11(+zero)", "S1 trsm||Li11", "S2 syrk", "S3 potrf22||WT", "S4 Li22", "S5 Li21", "(after diag)", "rest"]
for n, v in zip(names, a.tolist()): print(f"{n:22s} {v/1e3:9.0f}k cycles  ({v/32/1e3:6.1f}k per diag block)")
