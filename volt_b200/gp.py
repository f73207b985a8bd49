"""The slice of the GPyTorch protocol the Volt hot path uses, re-created on top of the volt_b200 CUDA ops.

GPyTorch is not a dependency of this package: the reference only needs Kernel.__call__ -> lazy tensor
(.evaluate()/.detach()), Mean, GaussianLikelihood, MultivariateNormal (log_prob / sample), ExactGP (train / eval
__call__), ExactMarginalLogLikelihood and psd_safe_cholesky.  The heavy lifting of each --
covariance build, Cholesky, solves, log-det, MLL gradients, posterior, sampling -- runs in libvolt_b200.so; the
classes here only carry parameters and shapes.  Behaviour restated from GPyTorch 1.6-1.8 (not installable offline).
"""
import math
import warnings

import torch
from torch import nn
from torch.nn.functional import softplus

from . import ops
from .ops import NotPSDError, NumericalWarning  # noqa: F401  (re-exported)


# ------------------------------------------------------------------------------------------------ constraints
class Interval(nn.Module):
    """gpytorch.constraints.Interval: lower + (upper - lower) * sigmoid(raw)."""

    def __init__(self, lower_bound, upper_bound, transform=None, inv_transform=None, initial_value=None):
        super().__init__()
        self.lower_bound = torch.as_tensor(float(lower_bound))
        self.upper_bound = torch.as_tensor(float(upper_bound))

    def transform(self, t):
        return self.lower_bound.to(t) + (self.upper_bound - self.lower_bound).to(t) * torch.sigmoid(t)

    def inverse_transform(self, t):
        u = (t - self.lower_bound.to(t)) / (self.upper_bound - self.lower_bound).to(t)
        return torch.log(u) - torch.log1p(-u)


class GreaterThan(Interval):
    """gpytorch.constraints.GreaterThan: softplus(raw) + lower."""

    def __init__(self, lower_bound, **kw):
        super().__init__(lower_bound, math.inf)

    def transform(self, t):
        return softplus(t) + self.lower_bound.to(t)

    def inverse_transform(self, t):
        u = t - self.lower_bound.to(t)
        return u + torch.log(-torch.expm1(-u))


class Positive(GreaterThan):
    def __init__(self, **kw):
        super().__init__(0.0)


# ------------------------------------------------------------------------------------------------ module base
class Module(nn.Module):
    def register_constraint(self, param_name, constraint, replace=True):
        self.add_module(param_name + "_constraint", constraint)

    def register_prior(self, name, prior, param_or_closure, setting_closure=None):
        raise NotImplementedError("priors are outside the Volt hot path (only TrainBasicModel registers one)")

    def initialize(self, **kwargs):
        for name, val in kwargs.items():
            mod, leaf = self, name
            if "." in name:
                head, leaf = name.rsplit(".", 1)
                mod = self.get_submodule(head)
            p = getattr(mod, leaf)
            if not torch.is_tensor(val):
                val = torch.as_tensor(val)
            p.data = val.to(p).expand_as(p).clone()
        return self


# ------------------------------------------------------------------------------------------------ lazy covariance
class LazyKernelTensor:
    """Stand-in for gpytorch.lazy.LazyEvaluatedKernelTensor: remembers (kernel, x1, x2) and evaluates on demand.
    `fused()` exposes the generator spec (kind, x, gen) so the MLL never has to materialise K."""

    def __init__(self, x1, x2, kernel, last_dim_is_batch=False, **params):
        self.x1, self.x2, self.kernel = x1, x2, kernel
        self.last_dim_is_batch, self.params = last_dim_is_batch, params

    def evaluate(self):
        return self.kernel.forward(self.x1, self.x2, diag=False, last_dim_is_batch=self.last_dim_is_batch, **self.params)

    to_dense = evaluate

    def detach(self):
        return LazyKernelTensor(self.x1.detach(), self.x2.detach(), self.kernel, self.last_dim_is_batch, **self.params)

    def fused(self):
        spec = getattr(self.kernel, "fused_spec", None)
        return None if spec is None or self.last_dim_is_batch else spec(self.x1, self.x2)

    def add_jitter(self, j=1e-3):
        d = self.evaluate()
        return d + j * torch.eye(d.shape[-1], dtype=d.dtype, device=d.device)

    @property
    def shape(self):
        return self.evaluate().shape

    def size(self, *a):
        return self.evaluate().size(*a)

    def __getitem__(self, idx):
        return self.evaluate()[idx]


def _dense(c):
    return c.evaluate() if hasattr(c, "evaluate") else c


class Kernel(Module):
    """gpytorch.kernels.Kernel protocol (no lengthscale machinery: the Volt kernels set has_lengthscale = False)."""
    has_lengthscale = False

    def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), active_dims=None, lengthscale_prior=None,
                 lengthscale_constraint=None, eps=1e-6, **kwargs):
        super().__init__()
        self._batch_shape = torch.Size(batch_shape)

    @property
    def batch_shape(self):
        return self._batch_shape

    def __call__(self, x1, x2=None, diag=False, last_dim_is_batch=False, **params):
        x1_, x2_ = x1, x2
        if x1_.ndimension() == 1:
            x1_ = x1_.unsqueeze(1)
        if x2_ is not None:
            if x2_.ndimension() == 1:
                x2_ = x2_.unsqueeze(1)
            if not x1_.size(-1) == x2_.size(-1):
                raise RuntimeError("x1_ and x2_ must have the same number of dimensions!")
        if x2_ is None:
            x2_ = x1_
        if diag:
            res = self.forward(x1_, x2_, diag=True, last_dim_is_batch=last_dim_is_batch, **params)
            if res.dim() >= 2 and res.shape[-1] == res.shape[-2] == x1_.shape[-2] and x1_.shape[-2] > 1:
                res = torch.diagonal(res, dim1=-2, dim2=-1)
            return res
        return LazyKernelTensor(x1_, x2_, self, last_dim_is_batch, **params)


class Mean(Module):
    def __call__(self, x):
        if x.ndimension() == 1:
            x = x.unsqueeze(1)
        return self.forward(x)


class ConstantMean(Mean):
    def __init__(self, prior=None, batch_shape=torch.Size(), **kwargs):
        super().__init__()
        self.batch_shape = batch_shape
        self.register_parameter(name="constant", param=nn.Parameter(torch.zeros(*batch_shape, 1)))

    def forward(self, input):
        if input.shape[:-2] == self.batch_shape:
            return self.constant.expand(input.shape[:-1])
        return self.constant.expand(torch.broadcast_shapes(input.shape[:-1], self.constant.shape))


class LinearMean(Mean):
    def __init__(self, input_size, batch_shape=torch.Size(), bias=True):
        super().__init__()
        self.register_parameter(name="weights", param=nn.Parameter(torch.randn(*batch_shape, input_size, 1)))
        if bias:
            self.register_parameter(name="bias", param=nn.Parameter(torch.randn(*batch_shape, 1)))
        else:
            self.bias = None

    def forward(self, x):
        res = x.matmul(self.weights).squeeze(-1)
        if self.bias is not None:
            res = res + self.bias
        return res


# ------------------------------------------------------------------------------------------------ cholesky helper
def psd_safe_cholesky(A, upper=False, out=None, jitter=None, max_tries=3):
    """gpytorch.utils.cholesky.psd_safe_cholesky on the GPU (volt_potrf): jitter only on failure, only on the
    failing batch members, escalating jitter * 10^i."""
    L, _, _ = ops.potrf(A, jitter=jitter, max_tries=max_tries, check=True)
    return L.transpose(-1, -2) if upper else L


# ------------------------------------------------------------------------------------------------ distributions
class MultivariateNormal:
    """gpytorch.distributions.MultivariateNormal (mean + lazy or dense covariance [+ likelihood noise])."""

    def __init__(self, mean, covariance_matrix, validate_args=False, _noise=None):
        self.loc = mean
        self._covar = covariance_matrix
        self._noise = _noise

    @property
    def mean(self):
        return self.loc

    @property
    def lazy_covariance_matrix(self):
        return self._covar

    @property
    def covariance_matrix(self):
        K = _dense(self._covar)
        if self._noise is not None:
            n = K.shape[-1]
            K = K + self._noise.to(K).unsqueeze(-1) * torch.eye(n, dtype=K.dtype, device=K.device)
        return K

    @property
    def event_shape(self):
        return self.loc.shape[-1:]

    @property
    def variance(self):
        return torch.diagonal(self.covariance_matrix, dim1=-2, dim2=-1)

    def log_prob(self, value):
        """-1/2 (r^T A^-1 r + logdet A + T log 2pi), Cholesky branch, fused on the GPU."""
        diff = value - self.loc
        T = diff.shape[-1]
        noise = self._noise if self._noise is not None else torch.zeros(1, dtype=diff.dtype, device=diff.device)
        spec = self._covar.fused() if isinstance(self._covar, LazyKernelTensor) else None
        if spec is not None:
            kind, x, gen = spec
        else:
            kind, x, gen = "dense", None, _dense(self._covar)
        return ops.exact_mll(kind, x, gen, diff, noise) * T

    def rsample(self, sample_shape=torch.Size(), base_samples=None):
        covar = self.covariance_matrix
        num_samples = sample_shape.numel() or 1
        H = covar.shape[-1]
        if base_samples is None:
            base_samples = torch.randn(*covar.shape[:-2], H, num_samples, dtype=self.loc.dtype, device=self.loc.device)
        res = ops.mvn_sample(self.loc.reshape(-1, H), covar.reshape(-1, H, H), base_samples.reshape(-1, H, num_samples))
        res = res.reshape(*covar.shape[:-2], num_samples, H)
        if covar.dim() > 2:
            res = res.movedim(-2, 0)
        res = res.to(self.loc.device)
        return res.reshape(sample_shape + self.loc.shape)

    def sample(self, sample_shape=torch.Size(), base_samples=None):
        with torch.no_grad():
            return self.rsample(sample_shape, base_samples)


# ------------------------------------------------------------------------------------------------ likelihood
class _HomoskedasticNoise(Module):
    def __init__(self, batch_shape=torch.Size()):
        super().__init__()
        self.register_parameter(name="raw_noise", param=nn.Parameter(torch.zeros(*batch_shape, 1)))
        self.register_constraint("raw_noise", GreaterThan(1e-4))

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        if not torch.is_tensor(value):
            value = torch.as_tensor(value).to(self.raw_noise)
        self.initialize(raw_noise=self.raw_noise_constraint.inverse_transform(value))


class GaussianLikelihood(Module):
    """gpytorch.likelihoods.GaussianLikelihood: noise = softplus(raw_noise) + 1e-4, added to the diagonal."""

    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size(), **kwargs):
        super().__init__()
        self.noise_covar = _HomoskedasticNoise(batch_shape)

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value

    @property
    def raw_noise(self):
        return self.noise_covar.raw_noise

    @raw_noise.setter
    def raw_noise(self, value):
        self.noise_covar.initialize(raw_noise=value)

    def __call__(self, dist, *a, **k):
        return MultivariateNormal(dist.mean, dist.lazy_covariance_matrix, _noise=self.noise)


# ------------------------------------------------------------------------------------------------ ExactGP
class ExactGP(Module):
    """gpytorch.models.ExactGP: train mode returns the prior at the training inputs; eval mode returns the exact
    posterior at the test inputs (DefaultPredictionStrategy, Cholesky branch)."""

    def __init__(self, train_inputs, train_targets, likelihood):
        if train_inputs is not None and torch.is_tensor(train_inputs):
            train_inputs = (train_inputs,)
        super().__init__()
        if train_inputs is not None:
            self.train_inputs = tuple(t.unsqueeze(-1) if t.ndimension() == 1 else t for t in train_inputs)
            self.train_targets = train_targets
        else:
            self.train_inputs, self.train_targets = None, None
        self.likelihood = likelihood

    def _posterior(self, test_x):
        """Generic exact prediction: dense kernel blocks (GPU), volt_potrf, volt_potrs; subclasses may fuse."""
        train_x = self.train_inputs[0]
        n = train_x.shape[-2]
        full = self.forward(torch.cat([train_x, test_x], dim=-2))
        full_mean, full_cov = full.mean, full.covariance_matrix
        noise = self.likelihood.noise
        A = full_cov[..., :n, :n] + noise.unsqueeze(-1) * torch.eye(n, dtype=full_cov.dtype, device=full_cov.device)
        L = psd_safe_cholesky(A)
        K_ts = full_cov[..., :n, n:]
        rhs = torch.cat([K_ts, (self.train_targets - full_mean[..., :n]).unsqueeze(-1)], dim=-1)
        W = ops.potrs(L, rhs, forward_only=True)
        Wk, v = W[..., :-1], W[..., -1:]
        pred_mean = full_mean[..., n:] + (Wk.transpose(-1, -2) @ v).squeeze(-1)
        pred_cov = full_cov[..., n:, n:] - Wk.transpose(-1, -2) @ Wk
        return MultivariateNormal(pred_mean, pred_cov)

    def __call__(self, *args, **kwargs):
        inputs = [a.unsqueeze(-1) if a.ndimension() == 1 else a for a in args]
        if self.training:
            if not all(torch.equal(a, b) for a, b in zip(self.train_inputs, inputs)):
                raise RuntimeError("You must train on the training inputs!")
            return self.forward(*inputs, **kwargs)
        return self._posterior(inputs[0])


class ExactMarginalLogLikelihood(Module):
    """gpytorch.mlls.ExactMarginalLogLikelihood: likelihood(f).log_prob(y) / T."""

    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood, self.model = likelihood, model

    def forward(self, function_dist, target, *params):
        output = self.likelihood(function_dist, *params)
        res = output.log_prob(target)
        return res.div(function_dist.event_shape.numel())


# no-op stand-ins for gpytorch.settings context managers used around the hot path
class _Ctx:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class settings:
    max_cholesky_size = _Ctx      # the CUDA path is always the exact Cholesky branch
    cholesky_jitter = _Ctx
    fast_pred_var = _Ctx
    debug = _Ctx


__all__ = ["Interval", "GreaterThan", "Positive", "Module", "Kernel", "Mean", "ConstantMean", "LinearMean",
           "LazyKernelTensor", "MultivariateNormal", "GaussianLikelihood", "ExactGP", "ExactMarginalLogLikelihood",
           "psd_safe_cholesky", "NotPSDError", "NumericalWarning", "settings", "warnings"]
