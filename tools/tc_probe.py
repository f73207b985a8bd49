"""A/B probe: tcgen05 vs SIMT batched MLL kernel on the same inputs (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from volt_b200 import _lib, batched, ops
from oracle import volt_oracle as O

lib = _lib.load()
cases = [(2, 64), (2, 128), (3, 192), (4, 512), (2, 1024), (3, 399)] if len(sys.argv) < 2 else [tuple(map(int, a.split("x"))) for a in sys.argv[1:]]
for B, T in cases:
    x, vol, logy = batched.synth_series(B, T)
    resid = logy - logy.mean(-1, keepdim=True)
    raw = torch.linspace(-4, 1e-5, B)
    res = {}
    for impl in (0, 1):
        lib.volt_set_mll_impl(impl)
        out = batched.mll_and_grad(x.cuda(), vol.cuda(), resid.cuda(), raw.cuda())
        torch.cuda.synchronize()
        res[impl] = {k: v.detach().cpu() for k, v in out.items() if torch.is_tensor(v)}
    a, b = res[0], res[1]
    ref = [O.volt_mll_and_grad(x.double(), vol[i].double(), resid[i].double(), raw[i].double()) for i in range(B)]
    rm = torch.stack([r["mll"] for r in ref]).float()
    rg = torch.stack([r["draw_noise"] for r in ref]).float()
    ra = torch.stack([r["alpha"] for r in ref]).float()
    def rel(u, v): return float((u - v).abs().max() / v.abs().max())
    print(f"B={B} T={T}: info simt={a['info'].tolist()} tc={b['info'].tolist()} | mll rel simt={rel(a['mll'], rm):.2e} tc={rel(b['mll'], rm):.2e}"
          f" | dnoise rel simt={rel(a['draw_noise'], rg):.2e} tc={rel(b['draw_noise'], rg):.2e} | alpha rel simt={rel(a['alpha'], ra):.2e} tc={rel(b['alpha'], ra):.2e}", flush=True)
# timing c2
B, T = 1024, 512
x, vol, logy = batched.synth_series(B, T)
_, resid = ops.ma_mean("ewma", logy.cuda(), 25, want_resid=True)
raw = torch.full((B,), 1e-5).cuda()
xd, vd = x.cuda(), vol.cuda()
for impl in (0, 1):
    lib.volt_set_mll_impl(impl)
    for _ in range(3): batched.mll_and_grad(xd, vd, resid, raw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): batched.mll_and_grad(xd, vd, resid, raw)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"impl={impl} c2 {dt*1e3:.3f} ms/batch -> {B/dt:.0f} evals/s", flush=True)
