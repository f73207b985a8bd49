import os, sys
sys.path.insert(0, "/root/repo")
import torch
from volt_b200 import batched, ops
for B, T in ((3, 100), (2, 256), (1, 512), (2, 700)):
    x, vol, logy = batched.synth_series(B, T)
    _, resid = ops.ma_mean("ewma", logy.cuda(), 25, want_resid=True)
    raw = torch.full((B,), 1e-5).cuda()
    o = batched.mll_and_grad(x.cuda(), vol.cuda(), resid, raw)
    torch.cuda.synchronize()
    print(B, T, float(o["mll"][0]))
pv = torch.rand(2, 8, 5).cuda() * 0.1 + 0.1
out, di, si = batched.rollouts(x.cuda(), logy[:2].cuda(), vol[:2].cuda(), pv, eps=torch.randn(2, 8, 5).cuda(), k=25)
print(out.shape, int(di.sum()), int(si.sum()))
st = ops.rollout_stats(out, truth=out.mean(1))
print({k: float(v.sum()) for k, v in st.items()})
