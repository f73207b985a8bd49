import os, sys
sys.path.insert(0, "/root/repo")
import torch
from volt_b200 import batched, ops
rB, rT, rS, rH = 512, 256, 512, 30
rx, rvol, rlogy = batched.synth_series(rB, rT)
g = torch.Generator().manual_seed(1)
rpv = (rvol[:, -1:, None] * torch.exp(0.1 * torch.randn(rB, rS, rH, generator=g))).cuda()
rxd, rvd, ryd = rx.cuda(), rvol.cuda(), rlogy.cuda()
for _ in range(4):
    ops.rollout(rxd, ryd, rvd, rpv, eps=None, k=25, seed=3, check=False)
torch.cuda.synchronize()
