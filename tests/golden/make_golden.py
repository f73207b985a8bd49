"""Generate golden vectors by running the reference's OWN files (unchanged) under the GPyTorch stub.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Writes tests/golden/volt_golden.pt (inputs + outputs, float32, < 1 MB).  The fixtures are data
derived by executing the reference; no reference source is copied.
"""
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _gpytorch_stub as stub  # noqa: E402

warnings.simplefilter("ignore")


class RandnRecorder:
    """Wrap torch.randn so every base-normal draw the reference makes is recorded in call order."""

    def __init__(self):
        self.calls = []
        self._orig = torch.randn

    def __enter__(self):
        def rec(*a, **k):
            out = self._orig(*a, **k)
            self.calls.append(out.detach().clone())
            return out
        torch.randn = rec
        return self

    def __exit__(self, *exc):
        torch.randn = self._orig
        return False


def synth(T, seed, dt=1.0 / 252):
    g = torch.Generator().manual_seed(seed)
    x = torch.arange(T) * dt
    lv = torch.log(torch.tensor(0.2)) + torch.cumsum(1.25 * dt ** 0.5 * torch.randn(T, generator=g), 0)
    vol = lv.exp()
    ret = vol * dt ** 0.5 * torch.randn(T, generator=g)
    px = (torch.log(torch.tensor(10.0)) + torch.cat((torch.zeros(1), torch.cumsum(ret, 0)))).exp()  # T+1 prices
    return x, vol, px


def main():
    ref = stub.load_reference()
    G = {}

    # ---- G1 covariance build (VolKernel.py)
    x, vol, _ = synth(33, 1)
    kern = ref.kernels.VolatilityKernel()
    G["volk_1d"] = dict(x=x, vol=vol, K=kern(x, vol).evaluate(), diag=kern(x, vol, diag=True),
                        cumtrapz=ref.kernels.CumTrapz(vol * vol, x))
    xb = torch.arange(17) / 365.0
    volb = torch.stack([synth(17, 10 + b)[1] for b in range(3)])
    G["volk_batched"] = dict(x=xb, vol=volb,
                             K=kern(xb.unsqueeze(0).repeat(3, 1).unsqueeze(-1), volb.unsqueeze(-1)).evaluate())
    # last_dim_is_batch (VolKernel.py:24-26,36-37, marked "TODO: check this") is shape-inconsistent under
    # GPyTorch's (d, n, 1) convention and never exercised by the shipped drivers: not pinned.

    # ---- G2 BM kernel (BMKernel.py)
    bmk = ref.kernels.BMKernel()
    x1, x2 = torch.arange(16) / 252.0, torch.arange(16, 21) / 252.0
    G["bmk"] = dict(x1=x1, x2=x2, vol=bmk.vol.detach().clone(), raw_vol=bmk.raw_vol.detach().clone(),
                    K11=bmk(x1).evaluate().detach(), K12=bmk(x1, x2).evaluate().detach())

    # ---- G3 moving-average means (EWMA.py)
    _, _, px = synth(64, 3)
    y = px[1:].log()
    yb = torch.stack([synth(64, 20 + b)[2][1:].log() for b in range(3)])
    G["ewma"] = dict(y=y, yb=yb)
    for k in (5, 25, 100):
        G["ewma"][f"k{k}"] = ref.means.EWMA(y, k)
        G["ewma"][f"kb{k}"] = ref.means.EWMA(yb, k)
    tx = torch.arange(64) / 252.0
    other = torch.arange(65) / 252.0
    means = {}
    for name, cls in (("ewma", ref.means.EWMAMean), ("dewma", ref.means.DEWMAMean), ("tewma", ref.means.TEWMAMean),
                      ("meanrevert", ref.means.MeanRevertingEMAMean)):
        m = cls(tx, y, 10)
        means[name] = dict(train=m(tx), one=m(tx[-1:] + 1.0), other=m(other))
    G["means"] = dict(train_x=tx, train_y=y, k=10, theta=0.5, out=means)

    # ---- G5 training loops (train_utils.py) + G6 vol-model posterior + G4 rollouts
    n = 48
    x, vol, px = synth(n, 7)
    torch.manual_seed(11)
    vmod, vlh = ref.train_utils.TrainVolModel(x, vol, train_iters=10)
    G["train_vol"] = dict(x=x, vol=vol, iters=10, raw_noise=vlh.raw_noise.detach().clone(),
                          raw_vol=vmod.covar_module.raw_vol.detach().clone())
    test_x = torch.arange(5) / 252.0 + x[-1] + x[1]
    vmod.eval()
    post = vmod(test_x)
    with RandnRecorder() as rr:
        vs = post.sample(torch.Size((6,)))
    G["bmgp_post"] = dict(train_x=x, train_y=vol.log(), test_x=test_x, vol=vmod.covar_module.vol.detach().clone(),
                          noise=vlh.noise.detach().clone(), mean=post.mean.detach(),
                          cov=post.covariance_matrix.detach(), eps=rr.calls[0], samples=vs)
    vmod.train()

    for mean_func in ("ewma", "dewma", "tewma"):
        torch.manual_seed(5)
        volt, lh = ref.train_utils.TrainVoltMagpieModel(x, px[1:], vmod, vlh, vol, train_iters=10, k=10,
                                                        mean_func=mean_func)
        mll = sys.modules["gpytorch"].mlls.ExactMarginalLogLikelihood(lh, volt)
        loss = -mll(volt(x), px[1:].log())
        G[f"train_volt_{mean_func}"] = dict(x=x, px=px, vol=vol, k=10, iters=10,
                                            raw_noise=lh.raw_noise.detach().clone(), final_loss=loss.detach(),
                                            param_names=[n_ for n_, _ in volt.named_parameters()],
                                            requires_grad=[p.requires_grad for p in volt.parameters()])
        vmod.eval()
        for theta in (None, 0.5):
            with RandnRecorder() as rr:
                torch.manual_seed(123)
                out = ref.rollout_utils.Rollouts(x, px, test_x, volt, nsample=6, theta=theta)
            # calls[0] = vol-model base normals (H,S); calls[1:] = one (S,1,1) draw per step
            vol_eps = rr.calls[0]
            pv_post = vmod(test_x)
            pred_vol = (sys.modules["gpytorch"].utils.cholesky.psd_safe_cholesky(pv_post.covariance_matrix)
                        @ vol_eps).T.add(pv_post.mean.unsqueeze(0)).exp().detach()
            eps = torch.cat([c.reshape(6, 1) for c in rr.calls[1:]], dim=1)
            G[f"rollout_{mean_func}_{'none' if theta is None else 'th'}"] = dict(
                train_x=x, train_y=px, log_vol_path=vol.log(), test_x=test_x, pred_vol=pred_vol, eps=eps,
                k=10, theta=theta, samples=out.detach())
            # Rollouts mutated the model; rebuild for the next case
            vmod.train()
            torch.manual_seed(5)
            volt, lh = ref.train_utils.TrainVoltMagpieModel(x, px[1:], vmod, vlh, vol, train_iters=10, k=10,
                                                            mean_func=mean_func)
            vmod.eval()
        vmod.train()

    # one GeneratePrediction call of rollout_utils (single step, S draws) with mean reversion
    torch.manual_seed(5)
    volt, lh = ref.train_utils.TrainVoltMagpieModel(x, px[1:], vmod, vlh, vol, train_iters=1, k=10)
    pv = (vol[-1] * torch.exp(0.1 * torch.randn(6, 1))).detach()
    with RandnRecorder() as rr:
        s = ref.rollout_utils.GeneratePrediction(x, px, test_x[0:1], pv, volt, latent_mean=px.log().mean(), theta=0.3)
    G["genpred_step"] = dict(train_x=x, train_y=px, log_vol_path=vol.log(), test_x=test_x[0:1], pred_vol=pv,
                             eps=rr.calls[0], k=10, latent_mean=px.log().mean(), theta=0.3, samples=s.detach())

    # ---- G7 multi-point prediction with parametric means (the "VOLT + STANDARD MEAN" branch,
    #      experiments/stocks/GenerateMultiMeanPreds.py:114-118; the MA means only support one test point)
    for mean_func in ("constant", "loglinear"):
        vmod.train()
        torch.manual_seed(9)
        voltc, lhc = ref.train_utils.TrainVoltMagpieModel(x, px[1:], vmod, vlh, vol, train_iters=10, k=10,
                                                          mean_func=mean_func)
        mllc = sys.modules["gpytorch"].mlls.ExactMarginalLogLikelihood(lhc, voltc)
        lossc = -mllc(voltc(x), px[1:].log())
        mean_params = {n_: p.detach().clone() for n_, p in voltc.mean_module.named_parameters()}
        pvSH = (vol[-1] * torch.exp(0.1 * torch.randn(6, 5))).detach()
        with RandnRecorder() as rr:
            s = ref.rollout_utils.GeneratePrediction(x, px, test_x, pvSH, voltc)
        G[f"genpred_multi_{mean_func}"] = dict(
            train_x=x, train_y=px, vol=vol, test_x=test_x, pred_vol=pvSH, eps=rr.calls[0], k=10, iters=10,
            raw_noise=lhc.raw_noise.detach().clone(), mean_params=mean_params, final_loss=lossc.detach(),
            param_names=[n_ for n_, _ in voltc.named_parameters()],
            requires_grad=[p.requires_grad for p in voltc.parameters()], samples=s.detach())
        if mean_func == "constant":
            pvH = (vol[-1] * torch.exp(0.1 * torch.randn(5))).detach()
            with RandnRecorder() as rr:
                s = voltc.GeneratePrediction(test_x, pvH, n_sample=4)
            G["genpred_method"] = dict(train_x=x, train_y=px, vol=vol, test_x=test_x, pred_vol=pvH, eps=rr.calls[0],
                                       raw_noise=lhc.raw_noise.detach().clone(), mean_params=mean_params,
                                       samples=s.detach())

    # ---- one MLL + gradient evaluation of the data model at a second noise level
    volt.likelihood.raw_noise.data = torch.tensor([-4.0])
    volt.zero_grad()
    mll = sys.modules["gpytorch"].mlls.ExactMarginalLogLikelihood(lh, volt)
    val = mll(volt(x), px[1:].log())
    val.backward()
    G["mll_point"] = dict(x=x, vol=vol, logy=px[1:].log(), k=10, raw_noise=torch.tensor([-4.0]),
                          mll=val.detach(), draw_noise=lh.raw_noise.grad.detach().clone())

    out = os.path.join(HERE, "volt_golden.pt")
    torch.save(G, out)
    print("wrote", out, os.path.getsize(out), "bytes;", len(G), "groups")


if __name__ == "__main__":
    main()
