"""ctypes binding of include/unsflow.h -- the same symbols julia/UnsflowCUDA.jl `ccall`s.

The product has NO CPU fallback: if libunsflow_cuda.so is missing this module raises at
first use, and every compute entry returns an error status when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libunsflow_cuda.so")

c_double_p = C.POINTER(C.c_double)


class Vorts(C.Structure):
    """struct unsflow_vorts (include/unsflow.h) -- SoA view of Vector{TwoDVort}."""
    _fields_ = [("n", C.c_int64)] + [(k, c_double_p) for k in ("x", "z", "s", "vc", "vx", "vz")]


class Surf(C.Structure):
    """struct unsflow_surf (include/unsflow.h) -- TwoDSurf, typedefs.jl:38-176."""
    _fields_ = [
        ("ndiv", C.c_int32), ("naterm", C.c_int32),
        ("c", C.c_double), ("uref", C.c_double), ("pvt", C.c_double),
        ("lespcrit", C.c_double),
        ("levflag", C.c_int32), ("_pad", C.c_int32),
        ("kinem", C.c_double * 6),
        ("a0", C.c_double), ("a0dot", C.c_double), ("a0prev", C.c_double),
    ] + [(k, c_double_p) for k in (
        "theta", "x", "cam", "cam_slope", "bnd_x", "bnd_z", "uind", "wind", "downwash",
        "aterm", "adot", "aprev")] + [("bv", Vorts)]


class Field(C.Structure):
    """struct unsflow_field (include/unsflow.h) -- TwoDFlowField, typedefs.jl:20-36."""
    _fields_ = [("u", C.c_double), ("w", C.c_double), ("tev", Vorts), ("lev", Vorts), ("extv", Vorts)]


class Opts(C.Structure):
    """struct unsflow_opts (include/unsflow.h)."""
    _fields_ = [
        ("driver", C.c_int32), ("solve_mode", C.c_int32), ("delvort_kind", C.c_int32), ("device", C.c_int32),
        ("del_limit", C.c_int64), ("del_dist", C.c_double), ("del_tol", C.c_double),
        ("dt", C.c_double), ("max_steps", C.c_int64),
    ]


DRIVER_LDVM, DRIVER_LDVM2DOF, DRIVER_LDVMLIN, DRIVER_LAUTAT = 0, 1, 2, 3
SOLVE_EXACT, SOLVE_NLSOLVE = 0, 1
DEL_NONE, DEL_SPALART = 0, 1
KERNEL_AUTO, KERNEL_DIRECT, KERNEL_SYMMETRIC = 0, 1, 2

_lib = None


class UnsflowError(RuntimeError):
    pass


def lib():
    """Load libunsflow_cuda.so (built in-tree by __graft_entry__.build()).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # fresh checkout: compile the extension in-tree once (nvcc cross-compiles without a GPU);
        # this is a build step, not a fallback -- if it fails the error below is raised
        import subprocess
        try:
            subprocess.run(["make", "-s", "-j8", "-C", os.path.join(_HERE, "csrc")], check=True, capture_output=True, timeout=900)
        except Exception:
            pass
    if not os.path.exists(LIB_PATH):
        raise UnsflowError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C unsteadyflowsolvers.jl_b200/csrc). There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    L.unsflow_last_error.restype = C.c_char_p
    d, i64, vp = c_double_p, C.c_int64, C.c_void_p
    sigs = {
        "unsflow_version": [],
        "unsflow_device_count": [C.POINTER(C.c_int)],
        "unsflow_set_device": [C.c_int],
        "unsflow_device_info": [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(i64), C.POINTER(i64)],
        "unsflow_ind_vel": [i64, d, d, d, d, i64, d, d, d, d],
        "unsflow_mutual_ind": [i64, d, d, d, d, d, d],
        "unsflow_wakeroll": [i64, d, d, d, d, d, d, i64, d, d, d, d, C.c_double, C.c_double, C.c_double],
        "unsflow_update_indbound": [i64, d, d, d, d, i64, d, d, d, d],
        "unsflow_release_scratch": [],
        "unsflow_set_precision": [C.c_int],
        "unsflow_nccl_unique_id": [vp],
        "unsflow_partition": [i64, C.c_int, C.c_int, C.c_int, C.POINTER(i64), C.POINTER(i64)],
        "unsflow_wake_create": [C.POINTER(vp), i64, i64, C.c_int, C.c_int, vp, C.c_int],
        "unsflow_wake_upload": [vp, d, d, d, d, d, d, d, d],
        "unsflow_wake_download": [vp, d, d, d, d],
        "unsflow_wake_step": [vp, i64, C.c_double, C.c_double, C.c_double],
        "unsflow_wake_sync": [vp],
        "unsflow_wake_timing": [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(i64)],
        "unsflow_wake_destroy": [vp],
        "unsflow_ldvm_create": [C.POINTER(vp), C.POINTER(Surf), C.POINTER(Field), C.POINTER(Opts)],
        "unsflow_ldvm_multi_create": [C.POINTER(vp), C.c_int32, C.POINTER(Surf), C.POINTER(Field), C.POINTER(Opts)],
        "unsflow_ldvm_multi_get_state": [vp, C.c_int32, C.POINTER(Surf), C.POINTER(Field)],
        "unsflow_lve_create": [C.POINTER(vp), C.POINTER(Surf), C.POINTER(Field), C.POINTER(Opts), C.c_double, C.c_double, C.c_double,
                               C.c_double, C.c_double, C.c_double],
        "unsflow_lve_get_circ": [vp, d, d],
        "unsflow_ldvm_set_kinematics": [vp, i64, d],
        "unsflow_ldvm_set_parallel": [vp, C.c_int, C.c_int, vp],
        "unsflow_ldvm_parallel_info": [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)],
        "unsflow_ldvm_set_2dof": [vp, d, d],
        "unsflow_ldvm_get_2dof": [vp, d],
        "unsflow_ldvm_run": [vp, i64, d],
        "unsflow_ldvm_counts": [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)],
        "unsflow_ldvm_get_state": [vp, C.POINTER(Surf), C.POINTER(Field)],
        "unsflow_ldvm_timing": [vp, C.POINTER(C.c_float), C.POINTER(i64)],
        "unsflow_ldvm_destroy": [vp],
        "unsflow_ldvm_batch_run": [C.POINTER(Surf), C.POINTER(Opts), i64, i64, d, d, d, C.POINTER(i64), C.POINTER(i64)],
        "unsflow_ldvm2dof_batch_run": [C.POINTER(Surf), C.POINTER(Opts), i64, i64, d, d, d, d, d, d, C.POINTER(i64), C.POINTER(i64)],
        "unsflow_ldvm_batch_timing": [C.POINTER(C.c_float)],
        "unsflow_bench_sweep": [i64, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)],
        "unsflow_fp64_peak": [d, d],
        "unsflow_rsqrt_probe": [i64, C.c_double, C.c_double, d, d, d],
        "unsflow_rsqrt_probe_order": [C.c_int, i64, C.c_double, C.c_double, d],
    }
    for name, args in sigs.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = L
    return L


SYMBOLS_EXPECTED = None  # filled by tests from include/unsflow.h


def check(rc):
    if rc != 0:
        msg = lib().unsflow_last_error()
        raise UnsflowError(msg.decode() if msg else f"libunsflow_cuda error {rc}")


def dptr(a):
    """double* of a C-contiguous float64 numpy array (None -> NULL)."""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "need contiguous float64"
    return a.ctypes.data_as(c_double_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def device_count():
    n = C.c_int(0)
    check(lib().unsflow_device_count(C.byref(n)))
    return n.value
