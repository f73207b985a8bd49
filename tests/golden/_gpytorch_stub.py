"""Minimal stand-in for the slice of GPyTorch / BoTorch the Volt hot path touches.

TEST INFRASTRUCTURE ONLY.  GPyTorch and BoTorch are not installable in the build
container (no network), so the reference's own files (`/root/reference/voltron/...`)
cannot be imported as shipped.  `install()` registers fake `gpytorch` / `botorch`
modules in `sys.modules` so that those files load *unchanged* through importlib and
can be executed to produce golden vectors (see make_golden.py).

Everything here restates GPyTorch 1.6-1.8 behaviour FROM MEMORY;
the GPyTorch slice of the parity chain is therefore "unpinned" -- the goldens pin the
reference's own arithmetic (covariance build, EWMA, rollout algebra, model glue,
training-loop flag logic), not GPyTorch's.

Pure torch, CPU, float32/float64.  Never imported by the product package.
"""
import math
import sys
import types
import warnings

import torch
from torch import nn
from torch.nn.functional import softplus


# --------------------------------------------------------------------------- constraints
class Interval(nn.Module):
    def __init__(self, lower_bound, upper_bound, transform=None, inv_transform=None, initial_value=None):
        super().__init__()
        self.lower_bound = torch.as_tensor(float(lower_bound))
        self.upper_bound = torch.as_tensor(float(upper_bound))

    def transform(self, t):
        return self.lower_bound + (self.upper_bound - self.lower_bound) * torch.sigmoid(t)

    def inverse_transform(self, t):
        u = (t - self.lower_bound) / (self.upper_bound - self.lower_bound)
        return torch.log(u) - torch.log1p(-u)


class GreaterThan(Interval):
    def __init__(self, lower_bound, **kw):
        super().__init__(lower_bound, math.inf)

    def transform(self, t):
        return softplus(t) + self.lower_bound

    def inverse_transform(self, t):
        u = t - self.lower_bound
        return u + torch.log(-torch.expm1(-u))


class Positive(GreaterThan):
    def __init__(self, **kw):
        super().__init__(0.0)


# --------------------------------------------------------------------------- module base
class Module(nn.Module):
    def register_constraint(self, param_name, constraint, replace=True):
        self.add_module(param_name + "_constraint", constraint)

    def register_prior(self, name, prior, param_or_closure, setting_closure=None):
        self.__dict__.setdefault("_stub_priors", []).append((name, prior, param_or_closure))

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


# --------------------------------------------------------------------------- lazy kernel tensor
class LazyEvaluatedKernelTensor:
    def __init__(self, x1, x2, kernel, last_dim_is_batch=False, **params):
        self.x1, self.x2, self.kernel = x1, x2, kernel
        self.last_dim_is_batch, self.params = last_dim_is_batch, params

    def evaluate(self):
        return self.kernel.forward(self.x1, self.x2, diag=False,
                                   last_dim_is_batch=self.last_dim_is_batch, **self.params)

    def detach(self):
        return LazyEvaluatedKernelTensor(self.x1.detach(), self.x2.detach(), self.kernel,
                                         self.last_dim_is_batch, **self.params)

    def add_jitter(self, j=1e-3):
        d = self.evaluate()
        return d + j * torch.eye(d.shape[-1], dtype=d.dtype)

    @property
    def shape(self):
        return self.evaluate().shape

    def __getitem__(self, idx):
        return self.evaluate()[idx]


def _dense(c):
    return c.evaluate() if hasattr(c, "evaluate") else c


class Kernel(Module):
    has_lengthscale = False

    def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), active_dims=None,
                 lengthscale_prior=None, lengthscale_constraint=None, eps=1e-6, **kwargs):
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
        return LazyEvaluatedKernelTensor(x1_, x2_, self, last_dim_is_batch, **params)


class _Dummy(Module):
    def __init__(self, *a, **k):
        super().__init__()


# --------------------------------------------------------------------------- means
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


# --------------------------------------------------------------------------- cholesky helper
class NotPSDError(RuntimeError):
    pass


class NanError(RuntimeError):
    pass


class NumericalWarning(RuntimeWarning):
    pass


def psd_safe_cholesky(A, upper=False, out=None, jitter=None, max_tries=3):
    L, info = torch.linalg.cholesky_ex(A)
    if not torch.any(info):
        return L.transpose(-1, -2) if upper else L
    isnan = torch.isnan(A)
    if isnan.any():
        raise NanError("cholesky_cpu: matrix contains NaNs")
    if jitter is None:
        jitter = 1e-6 if A.dtype == torch.float32 else 1e-8
    Aprime = A.clone()
    jitter_prev = 0
    for i in range(max_tries):
        jitter_new = jitter * (10 ** i)
        diag_add = ((info > 0) * (jitter_new - jitter_prev)).unsqueeze(-1).expand(*Aprime.shape[:-1])
        Aprime.diagonal(dim1=-1, dim2=-2).add_(diag_add)
        jitter_prev = jitter_new
        warnings.warn(f"A not p.d., added jitter of {jitter_new:.1e} to the diagonal", NumericalWarning)
        L, info = torch.linalg.cholesky_ex(Aprime)
        if not torch.any(info):
            return L.transpose(-1, -2) if upper else L
    raise NotPSDError(f"Matrix not positive definite after repeatedly adding jitter up to {jitter_new:.1e}.")


# --------------------------------------------------------------------------- distributions
class MultivariateNormal:
    def __init__(self, mean, covariance_matrix, validate_args=False):
        self.loc = mean
        self._covar = covariance_matrix

    @property
    def mean(self):
        return self.loc

    @property
    def lazy_covariance_matrix(self):
        return self._covar

    @property
    def covariance_matrix(self):
        return _dense(self._covar)

    @property
    def event_shape(self):
        return self.loc.shape[-1:]

    @property
    def variance(self):
        return torch.diagonal(self.covariance_matrix, dim1=-2, dim2=-1)

    def log_prob(self, value):
        diff = value - self.loc
        covar = self.covariance_matrix
        L = psd_safe_cholesky(covar)
        z = torch.linalg.solve_triangular(L, diff.unsqueeze(-1), upper=False).squeeze(-1)
        inv_quad = (z * z).sum(-1)
        logdet = 2.0 * torch.diagonal(L, dim1=-2, dim2=-1).log().sum(-1)
        return -0.5 * sum([inv_quad, logdet, diff.size(-1) * math.log(2 * math.pi)])

    def rsample(self, sample_shape=torch.Size(), base_samples=None):
        covar = self.covariance_matrix
        num_samples = sample_shape.numel() or 1
        root = psd_safe_cholesky(covar)
        if base_samples is None:
            base_samples = torch.randn(*covar.shape[:-2], root.size(-1), num_samples,
                                       dtype=self.loc.dtype, device=self.loc.device)
        samples = root.matmul(base_samples)
        samples = samples.permute(-1, *range(samples.dim() - 1)).contiguous()
        res = samples + self.loc.unsqueeze(0)
        return res.view(sample_shape + self.loc.shape)

    def sample(self, sample_shape=torch.Size(), base_samples=None):
        with torch.no_grad():
            return self.rsample(sample_shape, base_samples)


# --------------------------------------------------------------------------- likelihood
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
        covar = dist.covariance_matrix
        n = covar.shape[-1]
        return MultivariateNormal(dist.mean, covar + self.noise.unsqueeze(-1) * torch.eye(n, dtype=covar.dtype))


# --------------------------------------------------------------------------- ExactGP
class ExactGP(Module):
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

    def __call__(self, *args, **kwargs):
        inputs = [a.unsqueeze(-1) if a.ndimension() == 1 else a for a in args]
        if self.training:
            if not all(torch.equal(a, b) for a, b in zip(self.train_inputs, inputs)):
                raise RuntimeError("You must train on the training inputs!")
            return self.forward(*inputs, **kwargs)
        # eval: exact posterior at the test inputs (DefaultPredictionStrategy, Cholesky branch)
        train_x = self.train_inputs[0]
        n = train_x.shape[-2]
        full_x = torch.cat([train_x, inputs[0]], dim=-2)
        full = self.forward(full_x, **kwargs)
        full_mean, full_cov = full.mean, full.covariance_matrix
        K_tt = full_cov[..., :n, :n]
        noise = self.likelihood.noise
        A = K_tt + noise.unsqueeze(-1) * torch.eye(n, dtype=K_tt.dtype)
        L = psd_safe_cholesky(A)
        resid = (self.train_targets - full_mean[..., :n]).unsqueeze(-1)
        mean_cache = torch.cholesky_solve(resid, L).squeeze(-1)
        K_st = full_cov[..., n:, :n]
        pred_mean = full_mean[..., n:] + (K_st @ mean_cache.unsqueeze(-1)).squeeze(-1)
        pred_cov = full_cov[..., n:, n:] - K_st @ torch.cholesky_solve(K_st.transpose(-1, -2), L)
        return MultivariateNormal(pred_mean, pred_cov)


class ExactMarginalLogLikelihood(Module):
    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood, self.model = likelihood, model

    def forward(self, function_dist, target, *params):
        output = self.likelihood(function_dist, *params)
        res = output.log_prob(target)
        num_data = function_dist.event_shape.numel()
        return res.div(num_data)


# --------------------------------------------------------------------------- settings (no-op context managers)
class _Ctx:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def install():
    """Register fake gpytorch / botorch modules. Idempotent."""
    if "gpytorch" in sys.modules and getattr(sys.modules["gpytorch"], "_IS_VOLT_STUB", False):
        return sys.modules["gpytorch"]

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    g = mod("gpytorch", _IS_VOLT_STUB=True, Module=Module)
    g.constraints = mod("gpytorch.constraints", Interval=Interval, GreaterThan=GreaterThan, Positive=Positive)
    g.kernels = mod("gpytorch.kernels", Kernel=Kernel, ScaleKernel=_Dummy, RBFKernel=_Dummy,
                    MaternKernel=_Dummy, MultitaskKernel=_Dummy, SpectralMixtureKernel=_Dummy)
    g.means = mod("gpytorch.means", Mean=Mean, ConstantMean=ConstantMean, LinearMean=LinearMean,
                  MultitaskMean=_Dummy)
    g.distributions = mod("gpytorch.distributions", MultivariateNormal=MultivariateNormal,
                          MultitaskMultivariateNormal=_Dummy)
    g.likelihoods = mod("gpytorch.likelihoods", GaussianLikelihood=GaussianLikelihood,
                        MultitaskGaussianLikelihood=_Dummy, Likelihood=_Dummy,
                        _OneDimensionalLikelihood=_Dummy)
    g.models = mod("gpytorch.models", ExactGP=ExactGP, ApproximateGP=_Dummy)
    g.mlls = mod("gpytorch.mlls", ExactMarginalLogLikelihood=ExactMarginalLogLikelihood, VariationalELBO=_Dummy)
    g.priors = mod("gpytorch.priors", NormalPrior=_Dummy)
    g.lazy = mod("gpytorch.lazy")
    g.settings = mod("gpytorch.settings", max_cholesky_size=_Ctx, num_gauss_hermite_locs=_Ctx,
                     fast_pred_var=_Ctx, debug=_Ctx, cholesky_jitter=_Ctx)
    g.utils = mod("gpytorch.utils")
    g.utils.cholesky = mod("gpytorch.utils.cholesky", psd_safe_cholesky=psd_safe_cholesky)
    g.utils.errors = mod("gpytorch.utils.errors", NotPSDError=NotPSDError, NanError=NanError)
    g.utils.warnings = mod("gpytorch.utils.warnings", NumericalWarning=NumericalWarning)
    b = mod("botorch")
    b.models = mod("botorch.models", KroneckerMultiTaskGP=_Dummy)
    return g


def load_reference(root="/root/reference"):
    """Load the hot-path reference files unchanged, under the stub. Returns a namespace of modules."""
    import importlib.util
    import os

    install()

    def pkg(name):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        return m

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec.loader.exec_module(m)
        return m

    v = pkg("voltron")
    k = pkg("voltron.kernels")
    bm = load("voltron.kernels.BMKernel", "voltron/kernels/BMKernel.py")
    vk = load("voltron.kernels.VolKernel", "voltron/kernels/VolKernel.py")
    k.BMKernel, k.VolatilityKernel, k.FBMKernel, k.CumTrapz = bm.BMKernel, vk.VolatilityKernel, _Dummy, vk.CumTrapz
    me = pkg("voltron.means")
    ew = load("voltron.means.EWMA", "voltron/means/EWMA.py")
    ll = load("voltron.means.loglinear_mean", "voltron/means/loglinear_mean.py")
    for n_ in ("EWMAMean", "DEWMAMean", "TEWMAMean", "MeanRevertingEMAMean", "EWMA"):
        setattr(me, n_, getattr(ew, n_))
    me.LogLinearMean = ll.LogLinearMean
    lk = pkg("voltron.likelihoods")
    lk.VolatilityGaussianLikelihood = _Dummy
    mo = pkg("voltron.models")
    bmgp = load("voltron.models.BMGP", "voltron/models/BMGP.py")
    mo.BMGP, mo.MultitaskBMGP = bmgp.BMGP, bmgp.MultitaskBMGP
    vg = load("voltron.models.VoltronGP", "voltron/models/VoltronGP.py")
    vm = load("voltron.models.VoltMagpie", "voltron/models/VoltMagpie.py")
    mo.VoltronGP, mo.VoltMagpie = vg.VoltronGP, vm.VoltMagpie
    mo.SingleTaskVariationalGP = mo.MaternGP = mo.SMGP = _Dummy
    tu = load("voltron.train_utils", "voltron/train_utils.py")
    ru = load("voltron.rollout_utils", "voltron/rollout_utils.py")
    v.kernels, v.means, v.models, v.train_utils, v.rollout_utils = k, me, mo, tu, ru
    return types.SimpleNamespace(kernels=k, means=me, models=mo, train_utils=tu, rollout_utils=ru,
                                 VolKernel=vk, BMKernel=bm, EWMA=ew)
