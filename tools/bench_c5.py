"""Config c5 (one series, T=8192) and c3 (256 x 1024): MLL+grad time through the C ABI."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from volt_b200 import batched, ops
out = {}
for name, B, T in (("c5", 1, 8192), ("T4096", 1, 4096), ("T2048x4", 4, 2048), ("c3", 256, 1024)):
    x, vol, logy = batched.synth_series(B, T, dt=1 / 365 if name == "c3" else 1 / 252)
    _, resid = ops.ma_mean("ewma", logy.cuda(), 25, want_resid=True)
    raw = torch.full((B,), 1e-5).cuda(); xd, vd = x.cuda(), vol.cuda()
    for _ in range(2): o = batched.mll_and_grad(xd, vd, resid, raw)
    torch.cuda.synchronize(); t0 = time.perf_counter(); n = 3
    for _ in range(n): o = batched.mll_and_grad(xd, vd, resid, raw)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    fl = B * (2 * T ** 3 / 3 + 4 * T * T)
    out[name] = dict(B=B, T=T, ms=dt * 1e3, evals_per_s=B / dt, algorithmic_TFLOPs=fl / dt / 1e12, mll=float(o["mll"][0]))
# CPU oracle (fp32, reference semantics: Cholesky MLL + autograd backward) on the same c5 input, host cores of this box
if "--cpu" in sys.argv:
    from oracle import volt_oracle as O
    torch.set_num_threads(os.cpu_count())
    x, vol, logy = batched.synth_series(1, 8192)
    mean = O.ewma(logy, 25)[..., :-1]
    K = O.vol_kernel(x, vol)
    r = torch.full((1,), 1e-5, requires_grad=True)
    t0 = time.perf_counter()
    mll = O.exact_mll(K, logy - mean, O.noise_from_raw(r))
    (-mll.sum()).backward()
    out["c5_cpu_oracle"] = dict(seconds=time.perf_counter() - t0, cores=os.cpu_count(), mll=float(mll))
print(json.dumps(out, indent=1))
