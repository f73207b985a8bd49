"""Rollout throughput (BASELINE config c4 per-GPU share: 512 series x 512 draws x 30 steps, T=256, EWMA k=25) and
covariance-build bandwidth (c2 shape).  Not the bench.py headline; numbers go to profiles/ and DESIGN.md."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from volt_b200 import batched, ops, _lib

def ev_time(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]

out = {}
# ---- rollout c4 (per-GPU share)
B, T, S, H, k = 512, 256, 512, 30, 25
x, vol, logy = batched.synth_series(B, T)
g = torch.Generator().manual_seed(1)
pv = (vol[:, -1:, None] * torch.exp(0.1 * torch.randn(B, S, H, generator=g))).cuda()
eps = torch.randn(B, S, H, generator=g).cuda()
xd, vd, yd = x.cuda(), vol.cuda(), logy.cuda()
ms_eps = ev_time(lambda: ops.rollout(xd, yd, vd, pv, eps=eps, k=k, check=False))
ms_phx = ev_time(lambda: ops.rollout(xd, yd, vd, pv, eps=None, k=k, seed=3, check=False))
n = B * S * H
out["rollout_c4_share"] = dict(series=B, draws=S, horizon=H, T=T, ms_given_eps=ms_eps, ms_philox=ms_phx,
                               step_samples_per_s_given_eps=n / ms_eps * 1e3, step_samples_per_s_philox=n / ms_phx * 1e3,
                               path_samples_per_s_philox=B * S / ms_phx * 1e3,
                               algorithmic_GBs_given_eps=12.0 * n / ms_eps / 1e6, algorithmic_GBs_philox=8.0 * n / ms_phx / 1e6)
# CPU oracle on a small sample for the ratio (reference semantics: full re-factorisation every step)
from oracle import volt_oracle as O
bs, ss = 1, 32
px = torch.cat((logy[:bs, :1], logy[:bs]), -1).exp()
test_x = x[-1] + x[1] * torch.arange(1, H + 1)
t0 = time.perf_counter()
O.rollouts(x, px[0], vol[0].log(), test_x, pv[0, :ss].cpu(), eps[0, :ss].cpu(), k)
dt = time.perf_counter() - t0
out["rollout_cpu_oracle"] = dict(sample=f"{bs} series x {ss} draws x {H} steps, T={T}", seconds=dt, step_samples_per_s=bs * ss * H / dt,
                                 cores=os.cpu_count())
# ---- covariance build bandwidth (c2: 1024 x 512 x 512 fp32 = 1.07 GB written)
B2, T2 = 1024, 512
x2, vol2, _ = batched.synth_series(B2, T2)
x2d, v2d = x2.cuda(), vol2.cuda()
K = torch.empty(B2, T2, T2, device="cuda")
lib = _lib.load(); st = torch.cuda.current_stream().cuda_stream
ms_cov = ev_time(lambda: _lib.check(lib.volt_vol_cov(x2d.data_ptr(), 0, v2d.data_ptr(), 1, B2, T2, None, 0, K.data_ptr(), st), "cov"))
out["vol_cov_c2"] = dict(B=B2, T=T2, ms=ms_cov, GBs=4.0 * B2 * T2 * T2 / ms_cov / 1e6)
B3, T3 = 4, 8192
x3, vol3, _ = batched.synth_series(B3, T3)
K3 = torch.empty(B3, T3, T3, device="cuda")
x3d, v3d = x3.cuda(), vol3.cuda()
ms_cov3 = ev_time(lambda: _lib.check(lib.volt_vol_cov(x3d.data_ptr(), 0, v3d.data_ptr(), 1, B3, T3, None, 0, K3.data_ptr(), st), "cov"))
out["vol_cov_T8192"] = dict(B=B3, T=T3, ms=ms_cov3, GBs=4.0 * B3 * T3 * T3 / ms_cov3 / 1e6)
print(json.dumps(out, indent=1))
