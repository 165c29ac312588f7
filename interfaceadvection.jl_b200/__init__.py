"""interfaceadvection.jl_b200 -- B200-native (sm_100a) VOF + CMOM advection path of InterfaceAdvection.jl.

The product is `libifadv_b200.so` (C ABI in include/ifadv.h, hand-written CUDA in csrc/).  This package is the
host-side mirror of the reference's operator interface for that path (same function names minus Julia's `!`,
same argument order and meaning, same error behaviour) over torch CUDA tensors, used by the parity tests and
bench.py.  PyTorch only supplies device memory and streams.  There is no CPU fallback.
"""
from ._lib import Context, IfadvError, Report, LIMITERS, NORMAL_SCHEMES, LIB_PATH, IFADV_NO_RHOUF  # noqa: F401
from .api import (  # noqa: F401
    BC, BCf, MPCFL, MPFMomStep, Flow, TwoPhaseSimulation, advect, advectVOF, advectVOFrhouu, advectfq, applyVOF, cVOF,
    from_numpy, jl_empty, mom_advect_step, u2rhou_advectfq, jl_zeros, rhou2u, sim_step, sim_time, sum_inside, to_numpy, u2rhou, context_for,
    viscSurfTenrhou, updateU, updateL, mom_step_forcing,
    LevelSet, computeL, redistaningStage, redistaning, metrics, enstrophy,
    Poisson, update, psolver, myproject, project_with,
    MultiLevelPoisson, Vcycle, smooth, residual, solver,
)
from . import vtkio  # noqa: F401,E402  (load! / save of VTK restart files, ext/IntfAdvReadVTKExt.jl)
