"""voltron.train_utils on the B200 path: TrainVolModel, TrainDataModel, TrainVoltMagpieModel
(voltron/train_utils.py:69-144,192-257).  Same signatures, same quirks (raw_noise := 1e-5, positional grad_flags,
the no-op `vol_lh.noise.data = ...`), same Adam loops; every `mll(output, y)` + `loss.backward()` is one launch of the
fused CUDA MLL kernel (volt_mll_grad_*), with analytic gradients."""
import torch

from . import gp
from .means import DEWMAMean, LogLinearMean, MeanRevertingEMAMean, TEWMAMean
from .models import BMGP, VoltMagpie, VoltronGP


def _adam_mll_loop(model, likelihood, train_x, target, lr, train_iters, printing):
    """`for i in range(train_iters): loss = -mll(model(train_x), y); loss.backward(); optimizer.step()`
    (train_utils.py:84-94, 131-141, 243-254).  The two loops the shipped drivers spend their time in -- the noise-only
    fit of the MA-mean data model and the (raw_noise, raw_vol) fit of the BM vol model -- run as a device-resident loop
    (fused.py: one CUDA-graph replay per Adam iteration, no host round trips); every other configuration (parametric
    means) takes the generic autograd path below."""
    from . import fused

    if train_iters > 0 and fused.try_fused_loop(model, likelihood, train_x, target, lr, train_iters, printing):
        return
    optimizer = torch.optim.Adam([{"params": model.parameters()}], lr=lr)
    mll = gp.ExactMarginalLogLikelihood(likelihood, model)
    print_every = 50
    for i in range(train_iters):
        optimizer.zero_grad()
        output = model(train_x)
        loss = -mll(output, target)
        loss.backward()
        if printing and i % print_every == 0:
            print("Iter %d/%d - Loss: %.3f" % (i + 1, train_iters, loss.item()))
        optimizer.step()


def TrainVolModel(train_x, vol_path, train_iters=1000, printing=False, kernel="bm"):
    """voltron/train_utils.py:69-95."""
    vol_lh = gp.GaussianLikelihood().to(train_x.device)
    vol_lh.noise.data = torch.tensor([1e-2])  # writes to a temporary, exactly like the reference (:71): a no-op
    vol_model = BMGP(train_x, vol_path.log(), vol_lh, kernel=kernel).to(train_x.device)
    _adam_mll_loop(vol_model, vol_lh, train_x, vol_path.log(), 0.01, train_iters, printing)
    return vol_model, vol_lh


def _set_flags(model, grad_flags):
    for idx, p in enumerate(model.parameters()):
        p.requires_grad = grad_flags[idx]


def _train_mode(model, lh):
    model.train()
    lh.train()
    model.vol_lh.train()
    model.vol_model.train()


def TrainDataModel(train_x, train_y, vol_model, vol_lh, vol_path, train_iters=1000, printing=False):
    """voltron/train_utils.py:98-144 -- VoltronGP + LogLinearMean, trains noise / weights / bias."""
    voltron_lh = gp.GaussianLikelihood()
    voltron = VoltronGP(train_x, train_y.log(), voltron_lh, vol_path)
    voltron.mean_module = LogLinearMean(1)
    voltron.mean_module.initialize_from_data(train_x, train_y.log())
    voltron.likelihood.raw_noise.data = torch.tensor([1e-5])
    voltron.vol_lh = vol_lh
    voltron.vol_model = vol_model
    _set_flags(voltron, [True, True, True, False, False, False])
    _train_mode(voltron, voltron_lh)
    _adam_mll_loop(voltron, voltron_lh, train_x, train_y.log(), 0.1, train_iters, printing)
    return voltron, voltron_lh


def TrainVoltMagpieModel(train_x, train_y, vol_model, vol_lh, vol_path, train_iters=1000, printing=False, k=25, theta=0.5,
                         mean_func="ewma"):
    """voltron/train_utils.py:192-257."""
    voltron_lh = gp.GaussianLikelihood().to(train_x.device)
    voltron = VoltMagpie(train_x, train_y.log(), voltron_lh, vol_path, k=k).to(train_x.device)
    mf = mean_func.lower()
    if mf in ["ewma", "dewma", "tewma", "meanrevert"]:
        grad_flags = [True, False, False, False]
        if mf == "dewma":
            voltron.mean_module = DEWMAMean(train_x, train_y.log(), k).to(train_x.device)
        elif mf == "tewma":
            voltron.mean_module = TEWMAMean(train_x, train_y.log(), k).to(train_x.device)
        elif mf == "meanrevert":
            voltron.mean_module = MeanRevertingEMAMean(train_x, train_y.log(), k, theta).to(train_x.device)
    elif mf == "constant":
        voltron.mean_module = gp.ConstantMean().to(train_x.device)
        grad_flags = [True, True, False, False, False]
    elif mf == "loglinear":
        voltron.mean_module = LogLinearMean(1).to(train_x.device)
        voltron.mean_module.initialize_from_data(train_x, train_y.log())
        grad_flags = [True, True, True, False, False, False]
    elif mf == "linear":
        voltron.mean_module = gp.LinearMean(1).to(train_x.device)
        grad_flags = [True, True, True, False, False, False]
    else:
        raise ValueError(mean_func)
    voltron.likelihood.raw_noise.data = torch.tensor([1e-5]).to(train_x.device)
    voltron.vol_lh = vol_lh.to(train_x.device)
    voltron.vol_model = vol_model.to(train_x.device)
    _set_flags(voltron, grad_flags)
    _train_mode(voltron, voltron_lh)
    _adam_mll_loop(voltron, voltron_lh, train_x, train_y.log(), 0.1, train_iters, printing)
    return voltron, voltron_lh


def LearnGPCV(train_x, train_y, train_iters=1000, printing=False, early_stopping=False, kernel="bm"):
    """voltron/train_utils.py:15-67 -- fit the GPCV variational GP to the scaled returns of one price series and return
    the predicted volatility path (n,) on train_x's device.  Runs device resident with analytic gradients
    (volt_b200.gpcv); many series at once: volt_b200.gpcv.learn_gpcv.  Only the Brownian-motion prior of the hot path is
    provided (kernel="fbm" is outside it)."""
    from . import gpcv

    if kernel != "bm":
        raise NotImplementedError("LearnGPCV: only kernel='bm' is provided (FBMKernel is outside the hot path)")
    x = torch.as_tensor(train_x)
    pred = gpcv.learn_gpcv(x.reshape(-1), torch.as_tensor(train_y).reshape(1, -1), train_iters=train_iters, printing=printing)
    return pred[0].to(x.device)

