"""ctypes binding of the C ABI declared in include/pvder_b200.h (libpvder_b200.so).

This is the only bridge between the Python host code and the CUDA kernels; signatures carry
plain pointers/sizes only.  There is deliberately no fallback: if the shared library is missing
the loader raises with the build command.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("PVDER_B200_LIB") or os.path.join(CSRC, "libpvder_b200.so")   # env override: kernel-variant sweeps
SOURCES = ["pvder_kernels.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-diag-suppress", "177", "-shared", "-Xcompiler", "-fPIC"]

ABI_VERSION = 3
MAX_STATES = 23
OBS_DIM = 11
N_ACTIONS = 5
FINE_LEVELS = 3
SI_K, SI_STEPS, SI_EPISODE, SI_STATUS, SI_DONE, SI_HIST, SI_WINDUP, SI_EXACT, SI_REDO_LIST, SI_REDO_CTRL, SI_FIELDS = 0, 1, 2, 3, 4, 5, 10, 11, 12, 13, 14
GOALS = {"voltage_regulation": 0, "Q_regulation": 1, "power_regulation": 2}
REWARD_TERMS = {"voltage_error": 0, "Q_error": 1, "power_error": 2, "Vdc_error": 3}
EVENT_MODES = {"none": 0, "philox": 1, "table": 2}
STATUS_OK, STATUS_BAD_ACTION, STATUS_NONFINITE, STATUS_UNBALANCED = 0, 1, 2, 3
THREE_PHASE_MODES = {"general": 0, "balanced": 1, "auto": 2, "split": 3}


def sd_fields(ns):
    return ns + 6


def sd_index(ns):
    return dict(Q_ref=ns, Vdc_ref=ns + 1, Vgrid=ns + 2, Sinsol=ns + 3, ep_return=ns + 4, last_reward=ns + 5)


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "Rf", "Rt", "Xt", "inv_Lf", "inv_wb", "Kp_GCC", "Ki_GCC", "Kp_DC", "Ki_DC", "Kp_Q", "Ki_Q", "wp",
        "Kp_PLL", "Ki_PLL", "inv_C", "w0", "dw", "vgs", "np_iph100", "np_irs", "kappa", "pv_scale",
        "Vrms_ref", "iref_limit", "m_limit10", "p_target", "q_target", "Lf", "Rf_Rt")]


class EnvConfigC(C.Structure):
    _fields_ = [
        ("par", Params),
        ("phases", C.c_int32), ("n_sub_per_step", C.c_int32), ("base_level", C.c_int32), ("done_substep", C.c_int32),
        ("discrete_reward", C.c_int32), ("goal", C.c_int32), ("auto_reset", C.c_int32), ("event_mode", C.c_int32),
        ("ev_start_k", C.c_int32), ("ev_step_k", C.c_int32), ("ev_count", C.c_int32),
        ("ev_voltage_enable", C.c_int32), ("ev_insol_enable", C.c_int32), ("balanced3", C.c_int32),
        ("refine_input_level", C.c_int32), ("refine_on_action", C.c_int32), ("startup_substeps", C.c_int32),
        ("startup_level", C.c_int32), ("reward_terms", C.c_int32 * 4),
        ("ev_v_min", C.c_double), ("ev_v_max", C.c_double), ("ev_s_min", C.c_double), ("ev_s_max", C.c_double),
        ("delQ_pu", C.c_double), ("delVdc_pu", C.c_double), ("max_sim_time", C.c_double),
        ("substeps_per_sec", C.c_double), ("seed", C.c_uint64), ("Q_ref0", C.c_double), ("Vdc_ref0", C.c_double),
        ("y0", C.c_double * MAX_STATES),
        ("vg_ratio_b", C.c_double), ("vg_ratio_c", C.c_double),
    ]


_vp, _i64, _i32, _u64, _dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_double
_cfgp = C.POINTER(EnvConfigC)

# name -> (restype, argtypes); every symbol include/pvder_b200.h declares
SIGNATURES = {
    "pvder_abi_version": (C.c_int, []),
    "pvder_error_string": (C.c_char_p, [C.c_int]),
    "pvder_sd_fields": (C.c_size_t, [C.c_int]),
    "pvder_si_fields": (C.c_size_t, []),
    "pvder_config_size": (C.c_size_t, []),
    "pvder_steady_state": (C.c_int, [C.POINTER(Params), C.c_int, _dbl, _dbl, _dbl, _dbl, _dbl, _vp, _vp, _vp]),
    "pvder_reset": (C.c_int, [_cfgp, _vp, _vp, _i64, _vp, _i32, _vp, _vp, _i64, _i64, _vp]),
    "pvder_step": (C.c_int, [_cfgp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    "pvder_step_record": (C.c_int, [_cfgp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _i64, _i64, _vp]),
    "pvder_generate_events": (C.c_int, [_cfgp, _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "pvder_sample_actions": (C.c_int, [_u64, _i64, _vp, _i64, _i64, _vp]),
    "pvder_qnet_policy": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_float, _u64, _i64, _vp, _vp, _vp, _i64, _i64, _vp]),
    "pvder_qnet_collect": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_float, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64,
                                      _vp, _vp, _vp, _i32, _i64, _i64, _vp]),
    "pvder_stats_reduce": (C.c_int, [_vp, _vp, _i64, C.c_int, _i64, _vp, _vp]),
    "pvder_fp64_peak": (C.c_int, [C.c_int, C.POINTER(_dbl), C.POINTER(_dbl)]),
    "pvder_env_create": (C.c_int, [_cfgp, _i64, _i64, C.POINTER(_vp)]),
    "pvder_env_destroy": (C.c_int, [_vp]),
    "pvder_env_reconfigure": (C.c_int, [_vp, _cfgp]),
    "pvder_env_set_event_tables": (C.c_int, [_vp, _vp, _vp]),
    "pvder_env_reset_host": (C.c_int, [_vp, _vp, _vp]),
    "pvder_env_step_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "pvder_env_step_host_compact": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "pvder_env_state_host": (C.c_int, [_vp, _vp, _vp]),
    "pvder_env_set_refs_host": (C.c_int, [_vp, _vp]),
    "pvder_env_device_ptrs": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64)]),
    "pvder_env_kernel_ms": (C.c_int, [_vp, C.POINTER(_dbl), C.POINTER(_i64)]),
    "pvder_env_pipeline_info": (C.c_int, [_vp, C.POINTER(C.c_int32), C.POINTER(_dbl)]),
    "pvder_plan_chunks": (C.c_int, [_i64, _dbl, C.POINTER(_i64)]),
    "pvder_host_alloc": (_vp, [C.c_size_t]),
    "pvder_host_free": (None, [_vp]),
}

_lib = None


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into csrc/libpvder_b200.so (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(_HERE), "include", "pvder_b200.h"))
    if not force and os.path.exists(LIB_PATH):
        if os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(d) for d in deps):
            return LIB_PATH
    tmp = f"{LIB_PATH}.{os.getpid()}.tmp"      # built under a private name, then renamed: never a half-written library
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(res.stderr)
    return LIB_PATH


def load():
    """Load libpvder_b200.so; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"CUDA extension {LIB_PATH} is missing. Build it with `python -c \"import __graft_entry__ as g; "
            "g.build()\"` (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.pvder_abi_version() != ABI_VERSION:
        raise RuntimeError("libpvder_b200.so ABI version mismatch; rebuild")
    if lib.pvder_config_size() != C.sizeof(EnvConfigC):
        raise RuntimeError("pvder_env_config layout differs between libpvder_b200.so and _cabi.EnvConfigC; rebuild")
    _lib = lib
    return lib


class PVDERError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = load().pvder_error_string(rc)
        raise PVDERError(f"libpvder_b200 call failed ({rc}): {msg.decode() if msg else '?'}")
