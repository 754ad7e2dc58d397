"""nqcuda -- host-side mirror of NeuralQuantum.jl's machine / sampler / operator / algorithm /
parallel interfaces over libnqcuda (hand-written CUDA for sm_100a).  Importing this package loads
the shared library and fails loudly if it is missing: there is no CPU fallback.
"""
from . import _lib
from ._lib import NQError, PosDefException, NotConvergedError, EXPORTS, LIB_PATH
from .core import Context, HomogeneousSpin, HomogeneousFock, Hilbert, unique_id
from .operators import (LocalOperator as _LocalOperator, KLocalOperatorRow, Liouvillian, liouvillian, sigmax, sigmay,
                        sigmaz, sigmam, sigmap, destroy, create, number, DeviceOperator)
from .machines import RBM, RBMSplit, NDM, NDMSymm, symmetry_maps, af_softplus, af_logcosh, init_random_pars_, ket, densitymatrix
from .samplers import (MetropolisSampler, MetropolisSamplerCache, LocalRule, ExchangeRule, NagyRule, OperatorRule, ExactSampler,
                       ExactSamplerCache)
from .algorithms import (SR, Descent, Nesterov, update_, local_scalar, local_grad, stat_analysis, Measurement, sr_cholesky,
                         sr_cg, sr_minres, sr_qlp, sr_shift, sr_multiplicative, sr_none)
from .iterative import BatchedSampler, BatchedObsDMSampler
from .parallel import shard_chains, init_comm, world_from_env
from . import models


def LocalOperator(hilb):
    """LocalOperator(hilb): the zero operator on `hilb` (KLocalZero.jl)."""
    return _LocalOperator(hilb)
