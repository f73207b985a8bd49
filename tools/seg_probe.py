"""Per-segment cycle breakdown of mll_batched_tc_kernel on the c2 workload (needs tools/build.sh --profile)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from volt_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.environ.get("VOLT_PROF_LIB") or os.path.join(os.path.dirname(_lib.LIB_PATH), "libvolt_prof.so")
from volt_b200 import batched, ops  # noqa: E402

B, T = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024), (int(sys.argv[2]) if len(sys.argv) > 2 else 400)
x, vol, logy = batched.synth_series(B, T)
_, resid = ops.ma_mean("ewma", logy.cuda(), 25, want_resid=True)
raw = torch.full((B,), 1e-5).cuda()
for _ in range(3):
    out = batched.mll_and_grad(x.cuda(), vol.cuda(), resid, raw)
torch.cuda.synchronize()
a = out["alpha"].reshape(-1)[:12].cpu()
dg = out["alpha"].reshape(-1)[12:20].cpu().tolist()
names = ["store A / loop", "GEMM A", "epilogue A (tmem ld + generator)", "stash", "diagonal block", "L_jj, dinv, Linv operand, z",
         "TRSM A", "store B, tr, alpha", "GEMM B", "TRSM B", "phase B tail", "final"]
tot = float(a.sum())
for n, v in zip(names, a.tolist()):
    print(f"{n:34s} {v / 1e3:9.0f}k cycles ({100 * v / tot:4.1f}%)")
print(f"CTA 0 total: {tot / 1e6:.2f} M cycles for {-(-B // 444)} series")
print("diagonal block (thread 0 = pivot warp): " + ", ".join(f"{n} {v / 1e3:.0f}k" for n, v in zip(
    ["pivot16 / LiT clear", "barrier after pivot", "panel solve", "trailing update", "inverse blocks"], dg)))
tm = out["alpha"].reshape(-1)[20:28].cpu().tolist()
print("W2 worker (thread 0): " + ", ".join(f"{n} {v / 1e3:.0f}k" for n, v in zip(
    ["call start (dep fence + barrier)", "wait full, first tile", "wait full, other tiles", "split + st + fence + arrive", "wait last done",
     "calls (x1000)", "k-tiles (x1000)"], tm)))
cw = out["alpha"].reshape(-1)[28:36].cpu().tolist()
print("W2 control warps (CTA 0): " + ", ".join(f"{n} {v / 1e3:.0f}k" for n, v in zip(
    ["MMA wait ready", "MMA issue + commit", "TMA wait ringfree", "TMA wait depready", "TMA wait done", "TMA issue"], cw)))
ex = out["alpha"].reshape(-1)[36:40].cpu().tolist()
print("phase B detail: " + ", ".join(f"{n} {v / 1e3:.0f}k" for n, v in zip(["preamble (Dinv -> LiT, stage, tr, alpha)", "store of the call"], ex)))
