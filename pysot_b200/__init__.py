"""pysot_b200 -- pySOT's RBF-surrogate hot path (fit / predict / predict_deriv and the candidate-search
acquisition) on NVIDIA B200, behind pySOT's own Python API.

    from pysot_b200 import GpuRBFInterpolant, CubicKernel, LinearTail, install
    rbf = GpuRBFInterpolant(dim=d, lb=lb, ub=ub, kernel=CubicKernel(), tail=LinearTail(d))
    install()          # SRBFStrategy / DYCORSStrategy / SOPStrategy now score candidates on the GPU

The numerical work is in ``libb200rbf.so`` (hand-written sm_100a CUDA, C ABI in ``include/b200rbf.h``), reached
through ctypes; there is no CPU fallback.
"""
from .interface import (  # noqa: F401
    HAVE_PYSOT, ConstantTail, CubicKernel, Kernel, LinearKernel, LinearTail, Surrogate, Tail, TPSKernel, identity,
    median_capping,
)
from .pareto import nd_sorting  # noqa: F401
from .rbf import GpuRBFInterpolant  # noqa: F401
from .candidates import (  # noqa: F401
    candidate_dycors, candidate_srbf, candidate_uniform, generate_dycors, generate_on_device, generate_srbf,
    generate_uniform, install, merit_select_indices, set_candidate_generator, uninstall, weighted_distance_merit,
)

__version__ = "0.1.0"
