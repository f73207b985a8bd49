"""Drop-in alias: `import voltron` resolves to the B200-native implementation (volt_b200) with the reference's module
layout (voltron/__init__.py:1-12 of g-benton/Volt).  Only the hot-path surface is provided."""
__version__ = "alpha-b200"
from volt_b200 import kernels, means, models, option_utils, rollout_utils, train_utils  # noqa: F401
from volt_b200.kernels import BMKernel, VolatilityKernel  # noqa: F401
from volt_b200.models import BMGP, VoltMagpie, VoltronGP  # noqa: F401
from volt_b200.rollout_utils import GeneratePrediction, Rollouts  # noqa: F401
from volt_b200.train_utils import LearnGPCV  # noqa: F401
import sys as _sys

for _n in ("kernels", "means", "models", "option_utils", "rollout_utils", "train_utils"):
    _sys.modules[f"voltron.{_n}"] = getattr(_sys.modules["volt_b200"], _n)
