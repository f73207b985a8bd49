"""Reference usage pattern (config c1): ONE series, TrainVolModel + TrainVoltMagpieModel + Rollouts through the voltron API,
timed on the GPU path and on the CPU oracle's training loop."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import volt_b200 as vb
from oracle import volt_oracle as O

T, iters = 256, 100
x, vol, logy = vb.batched.synth_series(1, T)
vol, logy = vol[0], logy[0]
px = torch.cat((logy[:1], logy)).exp()
out = {}
for name, dev in (("cpu_tensors", "cpu"), ("cuda_tensors", "cuda")):
    xx, vv, pp = x.to(dev), vol.to(dev), px.to(dev)
    vb.TrainVoltMagpieModel(xx, pp[1:], *vb.TrainVolModel(xx, vv, train_iters=2), vv, train_iters=2, k=25)  # warm-up
    torch.cuda.synchronize(); t0 = time.perf_counter()
    vmod, vlh = vb.TrainVolModel(xx, vv, train_iters=iters)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    volt, lh = vb.TrainVoltMagpieModel(xx, pp[1:], vmod, vlh, vv, train_iters=iters, k=25)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    vmod.eval()
    test_x = xx[-1] + xx[1] * torch.arange(1, 31, device=dev)
    import copy
    vb.Rollouts(xx, pp, test_x, copy.deepcopy(volt), nsample=1000)       # first call: workspace / attribute set-up (Rollouts mutates the model)
    volt2 = copy.deepcopy(volt)
    torch.cuda.synchronize(); t2b = time.perf_counter()
    s = vb.Rollouts(xx, pp, test_x, volt2, nsample=1000)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    t3 = t2 + (t3 - t2b)
    out[name] = dict(train_vol_ms_per_iter=(t1 - t0) / iters * 1e3, train_volt_ms_per_iter=(t2 - t1) / iters * 1e3,
                     rollouts_1000x30_ms=(t3 - t2) * 1e3, raw_noise=float(lh.raw_noise.detach()))
torch.set_num_threads(os.cpu_count())
t0 = time.perf_counter(); r = O.train_vol_model(x, vol, train_iters=iters); t1 = time.perf_counter()
r2 = O.train_voltmagpie_model(x, px[1:], vol, train_iters=iters, k=25); t2 = time.perf_counter()
out["cpu_oracle"] = dict(train_vol_ms_per_iter=(t1 - t0) / iters * 1e3, train_volt_ms_per_iter=(t2 - t1) / iters * 1e3,
                         raw_noise=float(r2["raw_noise"]), cores=os.cpu_count())
print(json.dumps(out, indent=1))
