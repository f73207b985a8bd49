"""voltron.means on the B200 path (voltron/means/__init__.py:1-3): moving-average "Magpie" means and LogLinearMean."""
import torch

from . import ops
from .gp import LinearMean, Mean


def EWMA(y, k):
    """voltron/means/EWMA.py:20-37 -- (..., T) -> (..., T+1) causal k-tap weighted mean (CUDA: ewma_kernel)."""
    return ops.ewma(y, k)


class _MAMean(Mean):
    """Shared forward(x) selection rule of the MA means (EWMA.py:46-54 and twins): last element for a single test
    point, [..., :-1] on the training grid, the full T+1 path otherwise.  No learnable parameters."""
    kind = "ewma"

    def __init__(self, train_x, train_y, k=20):
        super().__init__()
        self.k = k
        self.train_x = train_x
        self.train_y = train_y

    def _path(self):
        return ops.ma_mean(self.kind, self.train_y, self.k)

    def forward(self, x):
        path = self._path().to(self.train_x.device)
        if x.numel() == 1:
            return path[..., -1].unsqueeze(0)
        if torch.equal(x.squeeze(), self.train_x.squeeze()):
            return path[..., :-1]
        return path


class EWMAMean(_MAMean):
    """voltron/means/EWMA.py:39-54."""
    kind = "ewma"


class DEWMAMean(_MAMean):
    """voltron/means/EWMA.py:74-91 -- 2 e - EWMA(e)[:-1]."""
    kind = "dewma"


class TEWMAMean(_MAMean):
    """voltron/means/EWMA.py:94-113 -- 3 e - 3 ee + eee."""
    kind = "tewma"

    def __init__(self, train_x, train_y, k=20):
        super().__init__(train_x, train_y, k)
        self.alpha = 2.0 / (self.k + 1)


class MeanRevertingEMAMean(_MAMean):
    """voltron/means/EWMA.py:116-135 -- e[j] - theta (e[j-1] - latent_mean)."""
    kind = "meanrevert"

    def __init__(self, train_x, train_y, k=20, theta=0.5):
        super().__init__(train_x, train_y, k)
        self.theta = theta
        self.alpha = 2.0 / (self.k + 1)
        self.latent_mean = train_y.mean()

    def _path(self):
        return ops.ma_mean(self.kind, self.train_y, self.k, theta=self.theta, latent=self.latent_mean)


class LogLinearMean(LinearMean):
    """voltron/means/loglinear_mean.py:5-21 -- log(clamp(w x + b, 1e-6)); three scalars, evaluated by torch
    (the gradient reaches them through dMLL/dresid = -alpha/T from the CUDA kernel)."""

    def __init__(self, input_size, batch_shape=None, bias=True):
        if batch_shape is None:
            batch_shape = torch.Size()
        super().__init__(input_size=input_size, batch_shape=batch_shape, bias=bias)

    def initialize_from_data(self, x, y):
        with torch.no_grad():
            self.bias.data = y.exp().mean(-1, keepdim=True)

    def forward(self, x):
        return super().forward(x).clamp(min=1e-6).log()
