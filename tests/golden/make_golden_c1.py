"""Golden vectors for BASELINE config c1 -- the example.ipynb path (cells 2-3, 11-15): SABR-style synthetic series,
TrainVolModel -> TrainDataModel (VoltronGP + LogLinearMean) -> dmod.vol_model(test_x).sample() ->
dmod.GeneratePrediction(test_x, vol_pred, npx), produced by running the reference's OWN files (unchanged) under the
GPyTorch stub.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_c1.py
Writes tests/golden/c1_golden.pt (inputs + outputs, float32, a few KB).  Data derived by executing the reference; no
reference source is copied.  The GPyTorch slice (noise transform, MLL / T, psd_safe_cholesky, exact prediction) is the
stub's restatement, so these fixtures pin the reference's own arithmetic and glue (VoltronGP.py:12-95,
train_utils.py:98-144, loglinear_mean.py:5-21), not GPyTorch's.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _gpytorch_stub as stub  # noqa: E402
from make_golden import RandnRecorder  # noqa: E402

warnings.simplefilter("ignore")


def sabr_series(steps, seed=2019):
    """example.ipynb cells 2-3 with `steps` grid points (the notebook uses 400): returns train_x (steps-1,),
    prices F[1:], true vol V[1:], test_x (about steps/2 points after the training grid), dt."""
    np.random.seed(seed)
    F0, V0, alpha, beta, rho, T = 10.0, 0.2, 1.25, 0.9, -0.2, 1
    dt = T / steps
    dW = np.random.normal(0, np.sqrt(dt), steps * T)
    dZ = rho * dW + np.sqrt(1 - rho ** 2) * np.random.normal(0, np.sqrt(dt), steps * T)
    F, V = np.zeros(steps * T), np.zeros(steps * T)
    F[0], V[0] = F0, V0
    for t in range(1, steps * T):
        F[t] = F[t - 1] + V[t - 1] * (F[t - 1]) ** beta * dW[t]
        V[t] = V[t - 1] + alpha * V[t - 1] * dZ[t]
    train_x = torch.FloatTensor(np.linspace(0, T, steps - 1)) + dt
    test_x = torch.linspace(T + dt, 1.5 * T, int(.5 * steps) - 1) + dt
    return train_x, torch.FloatTensor(F)[1:], torch.FloatTensor(V)[1:], test_x, dt


def main():
    ref = stub.load_reference()
    G = {}
    steps, iters_vol, iters_data = 257, 20, 20          # n = 256 training points: BASELINE config c1
    train_x, px, vol, test_x, dt = sabr_series(steps)
    test_x = test_x[:32]
    G["data"] = dict(train_x=train_x, px=px, vol=vol, test_x=test_x, steps=steps)

    torch.manual_seed(2019)
    vmod, vlh = ref.train_utils.TrainVolModel(train_x, vol, train_iters=iters_vol)
    G["train_vol"] = dict(iters=iters_vol, raw_noise=vlh.raw_noise.detach().clone(),
                          raw_vol=vmod.covar_module.raw_vol.detach().clone())

    torch.manual_seed(7)
    dmod, dlh = ref.train_utils.TrainDataModel(train_x, px, vmod, vlh, vol, train_iters=0)
    init = {n: p.detach().clone() for n, p in dmod.mean_module.named_parameters()}
    torch.manual_seed(7)
    dmod, dlh = ref.train_utils.TrainDataModel(train_x, px, vmod, vlh, vol, train_iters=iters_data)
    mll = sys.modules["gpytorch"].mlls.ExactMarginalLogLikelihood(dlh, dmod)
    loss = -mll(dmod(train_x), px.log())
    G["train_data"] = dict(seed=7, iters=iters_data, init_mean_params=init,
                           raw_noise=dlh.raw_noise.detach().clone(),
                           mean_params={n: p.detach().clone() for n, p in dmod.mean_module.named_parameters()},
                           final_loss=loss.detach(), param_names=[n for n, _ in dmod.named_parameters()],
                           requires_grad=[p.requires_grad for p in dmod.parameters()])

    # example.ipynb cell 15
    dmod.eval()
    dlh.eval()
    dmod.vol_model.eval()
    post = dmod.vol_model(test_x)
    G["vol_post"] = dict(mean=post.mean.detach(), cov=post.covariance_matrix.detach())
    preds = []
    for npx in (1, 3):
        with RandnRecorder() as rr:
            vol_pred = dmod.vol_model(test_x).sample().exp()
        vol_eps = rr.calls[0]
        with RandnRecorder() as rr:
            px_pred = dmod.GeneratePrediction(test_x, vol_pred, npx)
        preds.append(dict(npx=npx, vol_eps=vol_eps, vol_pred=vol_pred.detach(), eps=rr.calls[0],
                          px_pred=px_pred.detach()))
    G["predict"] = preds

    out = os.path.join(HERE, "c1_golden.pt")
    torch.save(G, out)
    print("wrote", out, os.path.getsize(out), "bytes")
    print("raw_noise", float(dlh.raw_noise), "loss", float(loss), {n: p.flatten().tolist() for n, p in dmod.mean_module.named_parameters()})
    for p in preds:
        print("npx", p["npx"], "px_pred", tuple(p["px_pred"].shape), "vol_eps", tuple(p["vol_eps"].shape), "eps", tuple(p["eps"].shape))


if __name__ == "__main__":
    main()
