import sys, os, ctypes
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
a = out["alpha"].reshape(-1)[:48].cpu().reshape(4, 12)
names = ["other/storeprev", "gemmA", "epiA(ld+gen)", "stash", "diag64", "diagpost(L,dinv,linv,z)", "trsmA", "storeB+loop", "gemmB", "trsmB", "phaseB tail", "final"]
tot = a.sum(1)
for i, n in enumerate(names):
    print(f"{n:28s} " + "  ".join(f"{a[c, i]/1e3:9.0f}k ({100*a[c,i]/tot[c]:4.1f}%)" for c in range(4)))
print("total cycles", tot.tolist())
g = out["alpha"].reshape(-1)[48:54].cpu()
print("gemm_tc producer thread 32 (CTA 0, kcycles): other %.0f wait_stage %.0f st_split(+load wait) %.0f gload-issue %.0f fence %.0f arrive %.0f" % tuple((g / 1e3).tolist()))
