"""volt_b200 -- B200-native implementation of the Volt GP inference hot path (covariance build -> Cholesky ->
exact MLL + gradients -> Monte-Carlo rollout) behind the reference's voltron.* API.

Layout: csrc/ (CUDA kernels + C ABI, include/volt_b200.h), _lib.py (ctypes), ops.py (tensor-level ops + autograd),
gp.py (GPyTorch protocol slice), kernels.py / means.py / models.py / train_utils.py / rollout_utils.py (the voltron
mirror), batched.py (many independent series, multi-GPU sharding)."""
from . import batched, gp, gpcv, ops, option_utils  # noqa: F401
from .kernels import BMKernel, CumTrapz, VolatilityKernel  # noqa: F401
from .means import DEWMAMean, EWMAMean, LogLinearMean, MeanRevertingEMAMean, TEWMAMean  # noqa: F401
from .models import BMGP, VoltMagpie, VoltronGP  # noqa: F401
from .rollout_utils import GeneratePrediction, Rollouts  # noqa: F401
from .train_utils import LearnGPCV, TrainDataModel, TrainVolModel, TrainVoltMagpieModel  # noqa: F401

__version__ = "0.1.0"
