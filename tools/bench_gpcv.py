"""GPCV stage timing: ms per Adam iteration for one series and for a batch, next to the CPU oracle."""
import json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from volt_b200 import batched, gpcv
warnings.simplefilter("ignore")
out = {}
for name, B, n, iters in (("single_n400", 1, 400, 100), ("batch64_n400", 64, 400, 50), ("batch148_n256", 148, 256, 50)):
    x, vol, logy = batched.synth_series(B, n + 1)
    px = logy.exp().cuda()
    gpcv.learn_gpcv(x[:n].cuda(), px, train_iters=3)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pred = gpcv.learn_gpcv(x[:n].cuda(), px, train_iters=iters)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out[name] = dict(B=B, n=n, iters=iters, ms_per_iter=dt / iters * 1e3, series_iters_per_s=B * iters / dt)
if "--cpu" in sys.argv:
    from oracle import volt_oracle as O
    torch.set_num_threads(os.cpu_count())
    x, vol, logy = batched.synth_series(1, 401)
    t0 = time.perf_counter()
    O.learn_gpcv(x[:400], logy[0].exp(), train_iters=10)
    out["cpu_oracle_n400"] = dict(ms_per_iter=(time.perf_counter() - t0) / 10 * 1e3, cores=os.cpu_count())
print(json.dumps(out, indent=1))
