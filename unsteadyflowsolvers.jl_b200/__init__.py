"""unsflow-b200: B200-native (sm_100a, FP64) discrete-vortex time-step hot path of
UnsteadyFlowSolvers.jl behind the reference's own API names.

Host side (this package) mirrors the reference's Julia types and entry points; all compute
goes through the C ABI of libunsflow_cuda.so (include/unsflow.h).  There is no CPU fallback.
"""
from . import _cabi
from ._cabi import UnsflowError, device_count, lib
from .calcs import ind_vel, mutual_ind, update_indbound, wakeroll
from .delvort import delNone, delSpalart, delVortDef
from .kinem import (ConstDef, CosDef, EldRampReturnDef, EldRampReturntstartDef, EldUpDef, EldUpIntDef, EldUpInttstartDef, EldUptstartDef, FileDef,
                    KinemDef, KinemPar, LinearDef, MotionDef, SinDef, StepGustDef)
from .postprocess import calc_delcp, calc_edgeVel, writeStamp, meshgrid
from .solvers import LVE, ensemble_shard, kinematics_table, lautat, ldvm, ldvm2DOF, ldvm2DOF_batch, ldvm_batch, ldvmLin, lve_table
from .typedefs import KinemPar2DOF, TwoDFlowField, TwoDOFPar, TwoDSurf, TwoDVort, VortList
from .utils import find_tstep, simpleTrapz

__all__ = [n for n in dir() if not n.startswith("_")]
