"""tinyda_b200: a B200-native batched MCMC engine behind the tinyDA Python surface.

Same public names as the reference package (tinyDA/__init__.py star-exports): ``sample``,
``Posterior``, ``GaussianLogLike``, ``AdaptiveGaussianLogLike``, the proposals, ``Link``,
``to_inference_data`` / ``get_samples``.  The arithmetic runs in hand-written sm_100a CUDA
kernels reached through the C ABI in include/tinyda_b200.h; there is no CPU path.
"""
__version__ = "0.1.0"

from .models import LinearModel, Rosenbrock, Poisson1D
from .distributions import (GaussianLogLike, AdaptiveGaussianLogLike, AdaptiveLogLike,
                            DefaultGaussianLogLike, DiagonalGaussianLogLike,
                            IsotropicGaussianLogLike)
from .posterior import Posterior
from .link import Link, LinkSequence
from .proposal import (Proposal, IndependenceSampler, GaussianRandomWalk, CrankNicolson, OperatorWeightedCrankNicolson,
                       AdaptiveMetropolis, MALA, DREAMZ, DREAM, SingleDreamZ, MultipleTry)
from .lowering import lower_problem
from .sampler import sample
from .utils import get_MAP, get_ML, LinkEvaluator
from .diagnostics import to_inference_data, get_samples, to_xarray, ess_bulk, rhat
