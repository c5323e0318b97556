"""Build + ctypes binding of libstochy_b200.so (the C ABI in include/stochy_b200.h).

The library is built IN-TREE with nvcc for sm_100a and loaded with ctypes.  There is no CPU
fallback anywhere in this package: if the shared library is missing or no CUDA device is
present, the calls fail loudly.
"""
import ctypes as C
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
_CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.environ.get("STOCHY_B200_LIB") or os.path.join(_PKG, "libstochy_b200.so")   # same override as the Julia shim
SOURCES = ["stochy_b200.cu"]
HEADERS = ["sb_device.cuh", "sb_kernels.cuh", "sb_small.cuh", "sb_resident.cuh"]

OK, ENAN, EZEROMASS, EINVAL, EUNSUPPORTED, ECUDA, EPEER, ENOMEM = range(8)
FP32, FP64 = 0, 1
SYSTEMATIC, MULTINOMIAL, LITERAL = 0, 1, 2
W_LINEAR_F64, W_LOG_F64, W_LOG_F32 = 0, 1, 2
ERP_FLIP, ERP_RANDOMINTEGER, ERP_NORMAL, ERP_BETA, ERP_GAMMA, ERP_UNIFORM = range(6)
TILE = 2048

EXPORTS = [
    "sb_version", "sb_options_default", "sb_create", "sb_destroy", "sb_last_error", "sb_get_stats", "sb_set_profile",
    "sb_model_discrete_hmm", "sb_model_linear_gaussian", "sb_model_bayesnet", "sb_model_ops",
    "sb_smc_sweep", "sb_pmcmc_run", "sb_pmcmc_chains", "sb_mh_run", "sb_enum_hmm",
    "sb_resample_systematic", "sb_resample_multinomial", "sb_resample_literal", "sb_logsumexp", "sb_gather",
    "sb_dev_resample_systematic", "sb_dev_gather", "sb_sync", "sb_timer_start", "sb_timer_stop",
    "sb_erp_sample", "sb_erp_logpdf", "sb_erp_sample_dirichlet", "sb_erp_logpdf_dirichlet",
    "sb_erp_sample_discrete", "sb_erp_logpdf_discrete",
    "sb_get_history", "sb_get_retained", "sb_get_particles",
    "sb_dist_init", "sb_dist_reserve_history", "sb_dist_ipc_export", "sb_dist_ipc_import",
]


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("precision", C.c_int32), ("resample_mode", C.c_int32),
                ("retained_pick", C.c_int32), ("store_logw", C.c_int32), ("use_graph", C.c_int32),
                ("seed", C.c_uint64)]


class OpStep(C.Structure):
    _fields_ = [("draw", C.c_int32), ("score", C.c_int32), ("dp", C.c_double * 4), ("sp", C.c_double * 4),
                ("obs_off", C.c_int32), ("obs_n", C.c_int32), ("par_off", C.c_int32), ("par_n", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("device_ms", C.c_double), ("step_kernel_ms", C.c_double),
                ("particle_steps", C.c_int64), ("logZ_first", C.c_double), ("logZ_last", C.c_double),
                ("accepts", C.c_int64)]


def nvcc_command(out=LIB_PATH):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = [os.path.join(_CSRC, s) for s in SOURCES if os.path.exists(os.path.join(_CSRC, s))]
    return [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
            "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-I", os.path.join(_ROOT, "include"),
            "-o", out] + srcs + ["-ldl"]


def build(force=False, verbose=False):
    """Compile the CUDA extension in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    deps = [os.path.join(_CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(_ROOT, "include", "stochy_b200.h")]
    deps = [d for d in deps if os.path.exists(d)]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(d) <= os.path.getmtime(LIB_PATH) for d in deps):
        return LIB_PATH
    cmd = nvcc_command()
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    """Load libstochy_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libstochy_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(this package has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
    sig = {
        "sb_version": (i32, []),
        "sb_options_default": (None, [C.POINTER(Options)]),
        "sb_create": (i32, [C.POINTER(Options), C.POINTER(vp)]),
        "sb_destroy": (None, [vp]),
        "sb_last_error": (C.c_char_p, [vp]),
        "sb_get_stats": (i32, [vp, C.POINTER(Stats)]),
        "sb_set_profile": (i32, [vp, i32]),
        "sb_model_discrete_hmm": (i32, [vp, i32, i32, i32, vp, vp, vp, vp]),
        "sb_model_linear_gaussian": (i32, [vp, i32, dbl, dbl, dbl, dbl, dbl, dbl, vp]),
        "sb_model_bayesnet": (i32, [vp, i32, i32, vp, vp, vp, vp, i32, u32]),
        "sb_model_ops": (i32, [vp, i32, vp, vp, i64]),
        "sb_smc_sweep": (i32, [vp, i64, u32, i32, C.POINTER(dbl)]),
        "sb_pmcmc_run": (i32, [vp, i64, i64, vp, vp]),
        "sb_pmcmc_chains": (i32, [vp, i64, i64, i64, i64, vp, vp]),
        "sb_mh_run": (i32, [vp, i64, i64, i64, vp]),
        "sb_enum_hmm": (i32, [vp, vp, vp, C.POINTER(dbl)]),
        "sb_resample_systematic": (i32, [vp, vp, i32, i64, dbl, i64, vp]),
        "sb_resample_multinomial": (i32, [vp, vp, i32, i64, vp, i64, vp]),
        "sb_resample_literal": (i32, [vp, vp, i64, vp, i64, vp]),
        "sb_logsumexp": (i32, [vp, vp, i32, i64, C.POINTER(dbl)]),
        "sb_gather": (i32, [vp, vp, i32, i64, vp, i64, vp]),
        "sb_dev_resample_systematic": (i32, [vp, vp, i32, i64, dbl, i64, vp]),
        "sb_dev_gather": (i32, [vp, vp, i32, i64, vp, i64, vp]),
        "sb_sync": (i32, [vp]),
        "sb_timer_start": (i32, [vp]),
        "sb_timer_stop": (i32, [vp, C.POINTER(dbl)]),
        "sb_erp_sample": (i32, [vp, i32, vp, i64, u64, vp]),
        "sb_erp_logpdf": (i32, [vp, i32, vp, vp, i64, vp]),
        "sb_erp_sample_dirichlet": (i32, [vp, vp, i32, i64, u64, vp]),
        "sb_erp_logpdf_dirichlet": (i32, [vp, vp, i32, vp, i64, vp]),
        "sb_erp_sample_discrete": (i32, [vp, vp, i32, i64, u64, vp]),
        "sb_erp_logpdf_discrete": (i32, [vp, vp, i32, vp, i64, vp]),
        "sb_get_history": (i32, [vp, i32, i64, vp, vp, vp]),
        "sb_get_retained": (i32, [vp, vp]),
        "sb_get_particles": (i32, [vp, i64, vp]),
        "sb_dist_init": (i32, [vp, i32, i32]),
        "sb_dist_reserve_history": (i32, [vp, i32]),
        "sb_dist_ipc_export": (i32, [vp, i64, vp]),
        "sb_dist_ipc_import": (i32, [vp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L
