"""GPCV per-iteration time without the one-off initialisation: (t(110 iterations) - t(10 iterations)) / 100."""
import os, sys, time, warnings, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from volt_b200 import batched, gpcv
warnings.simplefilter("always")
for B, n in ((1, 400), (148, 256), (64, 400)):
    x, vol, logy = batched.synth_series(B, n + 1)
    px = logy.exp().cuda()
    gpcv.learn_gpcv(x[:n].cuda(), px, train_iters=3)
    ts = {}
    for iters in (10, 110):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        gpcv.learn_gpcv(x[:n].cuda(), px, train_iters=iters)
        torch.cuda.synchronize(); ts[iters] = time.perf_counter() - t0
    print(f"VOLT_GPCV_BMM={os.environ.get('VOLT_GPCV_BMM')} B={B} n={n}: {(ts[110] - ts[10]) / 100 * 1e3:.3f} ms/iter (init + 10 iters {ts[10] * 1e3:.1f} ms)")
