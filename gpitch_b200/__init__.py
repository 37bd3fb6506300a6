"""gpitch_b200 -- B200-native variational-GP inner loop of gpitch, behind gpitch's own kernel / model / likelihood /
windowing API (reference: gpitch/__init__.py re-exports).  See DESIGN.md and include/gpitch_b200.h."""
from .methods import (logistic, ilogistic, softplus, isoftplus, gaussfun, logistic_tf, softplus_tf, gaussfun_tf,  # noqa: F401
                      midi2freq, freq2midi, find_ideal_f0)
from .window_overlap import segmented  # noqa: F401
from .init_models import init_liv, init_iv  # noqa: F401
from . import window_overlap, methods, param, kernels, matern12_spectral_mixture, init_kernels, likelihoods, train  # noqa: F401
from . import sgpr_ss, pdgp, batched, synthetic, init_models, kernelfit, driver  # noqa: F401
from .matern12_spectral_mixture import Matern12sm, MercerMatern12sm  # noqa: F401
from .kernels import Matern32, Add  # noqa: F401
from .sgpr_ss import SGPRSS  # noqa: F401
from .pdgp import Pdgp  # noqa: F401
from .likelihoods import MpdLik, ModLik  # noqa: F401
from .train import AdamOptimizer  # noqa: F401
from .audio import Audio  # noqa: F401
from . import transcription  # noqa: F401
from .transcription import AMT, SoSp  # noqa: F401
