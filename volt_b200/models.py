"""voltron.models on the B200 path (voltron/models/__init__.py:1-6): BMGP, VoltMagpie, VoltronGP."""
import torch

from . import gp, ops
from .gp import ExactGP, GaussianLikelihood, MultivariateNormal
from .kernels import BMKernel, VolatilityKernel
from .means import EWMAMean


class BMGP(ExactGP):
    """voltron/models/BMGP.py:9-28 -- exact GP on log-vol: mean -1/2 vol^2 x, Brownian-motion kernel."""

    def __init__(self, train_x, train_y, likelihood, kernel="bm"):
        super().__init__(train_x, train_y, likelihood)
        if kernel != "bm":
            raise NotImplementedError("only the Brownian-motion vol kernel is on the hot path (FBM: SURVEY.md 2.1 #3)")
        self.covar_module = BMKernel()
        self.scaling = train_x[1] - train_x[0]

    def mean_module(self, x):
        return -0.5 * self.covar_module.vol.pow(2.0) * x.squeeze()

    def forward(self, x):
        return MultivariateNormal(self.mean_module(x), self.covar_module(x))

    def _posterior(self, test_x):
        """Fused eval-mode posterior (volt_bmgp_posterior): build + potrf + solves in CUDA."""
        x = self.train_inputs[0][..., 0]
        mean, cov = ops.bmgp_posterior(x, self.train_targets, test_x[..., 0], self.covar_module.vol.detach(),
                                       self.likelihood.noise.detach())
        dev = x.device
        return MultivariateNormal(mean[0].to(dev), cov[0].to(dev))


class _VoltBase(ExactGP):
    """Shared body of VoltMagpie (VoltMagpie.py:15-127) and VoltronGP (VoltronGP.py:11-122)."""

    def _init_common(self, train_x, train_y, vol_path):
        if train_y.ndim > 1:
            raise NotImplementedError("the batched constructor routes to BoTorch's MultitaskBMGP in the reference "
                                      "(VoltMagpie.py:51-55); use volt_b200.batched for many independent series")
        self.covar_module = VolatilityKernel().to(train_x.device)
        self.train_x = train_x
        self.train_y = train_y
        if vol_path is None:
            self.log_vol_path = -1 * torch.ones(train_x.shape[0])
        else:
            self.log_vol_path = vol_path.log()
        self.train_cov = self.covar_module(self.train_x.unsqueeze(-1), self.log_vol_path.exp().unsqueeze(-1)).detach()
        self.vol_lh = GaussianLikelihood()
        self.vol_model = BMGP(train_x, self.log_vol_path, self.vol_lh)

    def UpdateVolPath(self, vol_path):
        self.log_vol_path = vol_path.log()
        self.train_cov = self.covar_module(self.train_x, self.log_vol_path.exp())

    def VolMLL(self):
        vol_mll = gp.ExactMarginalLogLikelihood(self.vol_lh, self.vol_model)
        return vol_mll(self.vol_model(self.train_x), self.log_vol_path)

    def GeneratePrediction(self, test_x, pred_vol, n_sample=1):
        """VoltMagpie.py:67-99 -- joint draw at all test points, psd_safe_cholesky default jitter, n_sample columns."""
        from .rollout_utils import _generate_prediction

        eps = torch.randn(test_x.shape[0], n_sample)
        pv = pred_vol.reshape(1, -1).expand(n_sample, -1)
        out = _generate_prediction(self, test_x, pv, eps.t().contiguous(), None, 0.5, jitter=1e-6,
                                   train_x=self.train_x, train_y=self.train_y, log_vol_path=self.log_vol_path,
                                   train_inputs_for_mean=self.train_inputs[0])
        return out.t().squeeze(-1)  # (H, n_sample); (H,) for n_sample == 1 -- the reference's trailing .squeeze(-1) (:97-99)

    def SamplePrediction(self, test_x, n_sample=1, return_vol=False):
        self.vol_model.eval()
        pred_vol = self.vol_model(test_x).sample().exp()
        prediction = self.GeneratePrediction(test_x, pred_vol, n_sample)
        return (prediction, pred_vol) if return_vol else prediction

    def MeanPrediction(self, test_x, n_sample=1, return_vol=False):
        self.vol_model.eval()
        pred_vol = self.vol_model(test_x).mean.exp()
        prediction = self.GeneratePrediction(test_x, pred_vol, n_sample)
        return (prediction, pred_vol) if return_vol else prediction

    def forward(self, x):
        mean_x = self.mean_module(x)
        if torch.equal(x, self.train_inputs[0]):
            covar_x = self.train_cov
        else:
            covar_x = self.covar_module(x, self.log_vol_path.exp())
        return MultivariateNormal(mean_x, covar_x)


class VoltMagpie(_VoltBase):
    """voltron/models/VoltMagpie.py:15-127 -- EWMA ("Magpie") mean + Volatility kernel, cached detached train_cov."""

    def __init__(self, train_x, train_y, likelihood, vol_path=None, k=25):
        super().__init__(train_x, train_y, likelihood)
        self.mean_module = EWMAMean(train_x, train_y, k).to(train_x.device)
        self._init_common(train_x, train_y, vol_path)


class VoltronGP(_VoltBase):
    """voltron/models/VoltronGP.py:11-122 -- same model with a LinearMean default (TrainDataModel swaps in LogLinearMean)."""

    def __init__(self, train_x, train_y, likelihood, vol_path=None):
        super().__init__(train_x, train_y, likelihood)
        self.mean_module = gp.LinearMean(1)
        self._init_common(train_x, train_y, vol_path)
