"""voltron.kernels on the B200 path: VolatilityKernel, BMKernel, CumTrapz (voltron/kernels/__init__.py:1-5)."""
import torch

from . import ops
from .gp import Interval, Kernel


def CumTrapz(y, x):
    """voltron/kernels/VolKernel.py:4-10 -- trapezoid-weighted cumulative sum (CUDA: cumtrapz_kernel)."""
    return ops.cumtrapz(y, x, vol_mode=ops.VOL_RAW, half_last=True)


class VolatilityKernel(Kernel):
    """voltron/kernels/VolKernel.py:12-41.  K[..., i, j] = V[..., min(i, j)], V = CumTrapz(vol^2, x).
    The second argument carries the VOLATILITY PATH, not a second set of inputs (the reference's convention)."""
    has_lengthscale = False

    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    @staticmethod
    def _squeeze(x, vol_path):
        if x.shape[-1] == 1:
            x = x.squeeze()
        if vol_path.shape[-1] == 1:
            vol_path = vol_path.squeeze()
        return x, vol_path

    def forward(self, x, vol_path, diag=False, **params):
        x, vol_path = self._squeeze(x, vol_path)
        last_dim_is_batch = params.get("last_dim_is_batch", False)
        if last_dim_is_batch:
            vol_path = vol_path.transpose(-1, -2)
        if diag:
            # diagonal of V[min(i,i)] is V itself (VolKernel.py:39-40)
            return ops.cumtrapz(vol_path, x, vol_mode=ops.VOL_SIGMA, half_last=True)
        res = ops.vol_cov(x, vol_path)
        if last_dim_is_batch:
            res = res.permute(1, 2, 0)
        return res

    def fused_spec(self, x1, x2):
        """(kind, grid, generator) consumed by the fused MLL kernel: the matrix is never materialised."""
        x, vol = self._squeeze(x1, x2)
        return "vol", x, vol


class BMKernel(Kernel):
    """voltron/kernels/BMKernel.py:6-51.  K = vol * min(x, x'), vol = sigmoid(raw_vol) (Interval(0, 1))."""
    has_lengthscale = False

    def __init__(self, vol=0.2, batch_shape=None, vol_constraint=None, **kwargs):
        vol_constraint = Interval(0.0, 1.0) if not vol_constraint else vol_constraint
        if batch_shape is None:
            batch_shape = torch.Size()
            vol_size = [1]
        else:
            vol_size = [*batch_shape, 1]
        super().__init__(batch_shape=batch_shape, lengthscale_constraint=vol_constraint, **kwargs)
        self.register_parameter("raw_vol", torch.nn.Parameter(torch.zeros(*vol_size)))
        self.register_constraint("raw_vol", vol_constraint)
        self.vol = vol

    def _set_vol(self, value):
        if not torch.is_tensor(value):
            value = torch.as_tensor(value).to(self.raw_vol)
        self.initialize(raw_vol=self.raw_vol_constraint.inverse_transform(value))

    @property
    def vol(self):
        return self.raw_vol_constraint.transform(self.raw_vol)

    @vol.setter
    def vol(self, value):
        return self._set_vol(value)

    def forward(self, x1s, x2s, **kwargs):
        if self.batch_shape != torch.Size():
            raise NotImplementedError("batched BMKernel belongs to MultitaskBMGP, outside the hot path (SURVEY.md 2.1 #8)")
        cov = ops.bm_cov(x1s[:, 0], x2s[:, 0], self.vol)
        if kwargs.pop("diag", False):
            return cov.diag()
        return cov

    def fused_spec(self, x1, x2):
        if x1.shape != x2.shape or not torch.equal(x1, x2):
            return None
        return "bm", x1[..., 0], self.vol
