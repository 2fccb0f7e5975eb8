"""ctypes binding of ``libdeluca_b200.so`` (the C ABI declared in ``include/deluca_b200.h``).

There is deliberately no fallback: if the shared library is missing, or the device is not a
Blackwell (sm_100) part, every entry point raises.  Nothing here imports ``oracle/``.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdeluca_b200.so")

c_f32p = ctypes.c_void_p  # device pointers travel as integers
_lock = threading.Lock()
_lib = None


class DlcError(RuntimeError):
    """Raised when an ABI call returns non-zero (message from ``dlc_last_error``)."""


class DlcLdsParams(ctypes.Structure):
    _fields_ = [
        ("N", ctypes.c_int64),
        ("env_offset", ctypes.c_int64),
        ("T", ctypes.c_int32),
        ("d", ctypes.c_int32),
        ("m", ctypes.c_int32),
        ("H", ctypes.c_int32),
        ("HH", ctypes.c_int32),
        ("policy", ctypes.c_int32),
        ("t0", ctypes.c_int32),
        ("decay", ctypes.c_int32),
        ("noise_uses_prev_action", ctypes.c_int32),
        ("shared_dynamics", ctypes.c_int32),
        ("traj_stride", ctypes.c_int32),
        ("lr_scale", ctypes.c_float),
        ("w_std", ctypes.c_float),
        ("bpc_delta", ctypes.c_float),
        ("kernel_variant", ctypes.c_int32),
        ("mode", ctypes.c_int32),
        ("seed", ctypes.c_uint64),
        ("env_t0", ctypes.c_int32),
        ("pad_", ctypes.c_int32),
    ]


class DlcLungParams(ctypes.Structure):
    _fields_ = [
        ("N", ctypes.c_int64), ("T", ctypes.c_int32), ("sim", ctypes.c_int32), ("nk", ctypes.c_int32),
        ("kf_per_env", ctypes.c_int32), ("kx", ctypes.c_float * 16), ("period", ctypes.c_float),
        ("xp_in", ctypes.c_float), ("pid_decay", ctypes.c_float), ("default_dt", ctypes.c_float),
        ("run_dt", ctypes.c_float), ("env_flow", ctypes.c_float),
        ("leak", ctypes.c_int32), ("peep_valve", ctypes.c_float), ("PC", ctypes.c_float), ("P0", ctypes.c_float),
        ("R", ctypes.c_float), ("dt", ctypes.c_float), ("min_volume", ctypes.c_float), ("r0", ctypes.c_float),
        ("r0_sq", ctypes.c_float), ("leak_coef", ctypes.c_float),
        ("delay", ctypes.c_int32), ("d_R", ctypes.c_float), ("d_C", ctypes.c_float), ("inertia", ctypes.c_float),
        ("control_gain", ctypes.c_float), ("mode", ctypes.c_int32), ("kernel_variant", ctypes.c_int32),
    ]


class DlcSharedGpcParams(ctypes.Structure):
    _fields_ = [
        ("N", ctypes.c_int64), ("T", ctypes.c_int32), ("d", ctypes.c_int32), ("m", ctypes.c_int32),
        ("H", ctypes.c_int32), ("HH", ctypes.c_int32), ("t0", ctypes.c_int32), ("decay", ctypes.c_int32),
        ("lr_scale", ctypes.c_float), ("reserved", ctypes.c_int32),
    ]


class DlcOfParams(ctypes.Structure):
    _fields_ = [
        ("N", ctypes.c_int64), ("env_offset", ctypes.c_int64), ("T", ctypes.c_int32), ("d", ctypes.c_int32),
        ("n", ctypes.c_int32), ("p", ctypes.c_int32), ("agent", ctypes.c_int32), ("m", ctypes.c_int32),
        ("h", ctypes.c_int32), ("F", ctypes.c_int32), ("L", ctypes.c_int32), ("t0", ctypes.c_int32),
        ("decay", ctypes.c_int32), ("shared_dynamics", ctypes.c_int32), ("mode", ctypes.c_int32),
        ("lr", ctypes.c_float), ("R_M", ctypes.c_float), ("w_std", ctypes.c_float), ("lanes_per_env", ctypes.c_int32),
        ("seed", ctypes.c_uint64), ("flags", ctypes.c_int32), ("env_t0", ctypes.c_int32),
    ]


OF_AGENT = {"lqg": 0, "drc": 1, "fgrc": 2}

LUNG_BUFFER_FIELDS = [
    "K", "kf", "u_in_in", "u_out_in", "steps_io", "time_io", "volume_io", "pressure_io", "target_io",
    "pipe_pressure_io", "in_history_io", "out_history_io", "obs_pressure_io", "obs_time_io",
    "pid_p_io", "pid_i_io", "pid_d_io", "pid_time_io", "pid_dt_io", "pid_steps_io",
    "exp_time_io", "exp_dt_io", "exp_steps_io",
    "timestamp_out", "u_in_out", "u_out_out", "pressure_out", "flow_out", "target_out", "status_out",
]


class DlcLungBuffers(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in LUNG_BUFFER_FIELDS]


class DlcEnvSpec(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("p", ctypes.c_float * 8)]


class DlcPlanParams(ctypes.Structure):
    _fields_ = [
        ("N", ctypes.c_int64), ("H", ctypes.c_int32), ("T", ctypes.c_int32),
        ("env_true", DlcEnvSpec), ("env_sim", DlcEnvSpec),
        ("q", ctypes.c_float * 8), ("goal", ctypes.c_float * 8), ("r", ctypes.c_float * 4),
        ("goal_u", ctypes.c_float * 4),
        ("algo", ctypes.c_int32), ("backtracking", ctypes.c_int32), ("fix_backtracking", ctypes.c_int32),
        ("n_alpha", ctypes.c_int32), ("w_is", ctypes.c_int32), ("alpha0", ctypes.c_float), ("lr", ctypes.c_float),
    ]


class DlcEnvAgentParams(ctypes.Structure):
    _fields_ = [
        ("N", ctypes.c_int64), ("T", ctypes.c_int32), ("agent", ctypes.c_int32), ("env", DlcEnvSpec),
        ("gains_per_env", ctypes.c_int32), ("state_stride", ctypes.c_int32), ("want_grad", ctypes.c_int32),
        ("pid_decay", ctypes.c_float),
        ("q", ctypes.c_float * 8), ("goal", ctypes.c_float * 8), ("r", ctypes.c_float * 4),
        ("goal_u", ctypes.c_float * 4),
    ]


POLICY = {"lqr": 0, "gpc": 1, "bpc": 2, "open": 3}

# every symbol include/deluca_b200.h declares: name -> (restype, argtypes)
_P = ctypes.c_void_p
SYMBOLS = {
    "dlc_abi_version": (ctypes.c_int32, []),
    "dlc_last_error": (ctypes.c_char_p, []),
    "dlc_check_device": (ctypes.c_int32, []),
    "dlc_gaussian_disturbance_f32": (ctypes.c_int32, [_P, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                                      ctypes.c_uint64, ctypes.c_int64, ctypes.c_int32,
                                                      ctypes.c_float, _P]),
    "dlc_sizeof": (ctypes.c_size_t, [ctypes.c_int32]),
    "dlc_lung_pid_rollout_f32": (ctypes.c_int32, [ctypes.POINTER(DlcLungParams), ctypes.POINTER(DlcLungBuffers), _P]),
    "dlc_lung_pid_grad_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(DlcLungParams)]),
    "dlc_lung_pid_grad_f32": (ctypes.c_int32, [ctypes.POINTER(DlcLungParams), _P, _P, _P, ctypes.c_int32,
                                               ctypes.c_float, ctypes.c_float, _P, _P, _P, _P, ctypes.c_size_t, _P]),
    "dlc_lds_lqr_grad_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32]),
    "dlc_lds_lqr_grad_f32": (ctypes.c_int32, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                              ctypes.c_int32, _P, _P, _P, _P, _P, ctypes.c_uint64, ctypes.c_int64,
                                              ctypes.c_int32, ctypes.c_float, _P, _P, _P, _P, _P, ctypes.c_size_t,
                                              _P]),
    "dlc_pid_f32": (ctypes.c_int32, [ctypes.c_int64, ctypes.c_int32, ctypes.c_float] + [_P] * 9),
    "dlc_env_rollout_f32": (ctypes.c_int32, [ctypes.POINTER(DlcPlanParams), ctypes.c_int32] + [_P] * 10),
    "dlc_linearize_f32": (ctypes.c_int32, [ctypes.POINTER(DlcPlanParams), ctypes.c_int32] + [_P] * 10),
    "dlc_riccati_backward_f32": (ctypes.c_int32, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
                                 + [_P] * 10),
    "dlc_lqg_backward_f32": (ctypes.c_int32, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
                             + [_P] * 9 + [ctypes.c_float, ctypes.c_float, ctypes.c_int32] + [_P] * 5),
    "dlc_igpc_rollout_f32": (ctypes.c_int32, [ctypes.POINTER(DlcPlanParams)] + [_P] * 10),
    "dlc_plan_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(DlcPlanParams)]),
    "dlc_plan_f32": (ctypes.c_int32, [ctypes.POINTER(DlcPlanParams)] + [_P] * 11 + [ctypes.c_size_t, _P]),
    "dlc_lds_shared_packed_floats": (ctypes.c_size_t, [ctypes.c_int32, ctypes.c_int32]),
    "dlc_lds_shared_pack_f32": (ctypes.c_int32, [_P, _P, _P, ctypes.c_int32, ctypes.c_int32, _P, _P]),
    "dlc_lds_shared_lqr_rollout_f32": (ctypes.c_int32, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                                        ctypes.c_int32] + [_P] * 7),
    "dlc_lds_shared_gpc_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(DlcSharedGpcParams)]),
    "dlc_lds_shared_gpc_rollout_f32": (ctypes.c_int32, [ctypes.POINTER(DlcSharedGpcParams)] + [_P] * 12
                                       + [ctypes.c_size_t, _P]),
    "dlc_lds_of_smem_bytes": (ctypes.c_size_t, [ctypes.POINTER(DlcOfParams)]),
    "dlc_lds_of_rollout_f32": (ctypes.c_int32, [ctypes.POINTER(DlcOfParams)] + [_P] * 18),
    "dlc_lds_open_rollout_f32": (ctypes.c_int32, [ctypes.c_int64] + [ctypes.c_int32] * 5 + [_P] * 6
                                 + [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_float,
                                    ctypes.c_int32, _P, _P, _P]),
    "dlc_sinus_disturbance_f32": (ctypes.c_int32, [_P, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, _P,
                                                   ctypes.c_float, ctypes.c_int32, _P]),
    "dlc_gae_f32": (ctypes.c_int32, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, _P, _P, _P, _P,
                                     ctypes.c_float, ctypes.c_float, ctypes.c_int32, _P, _P, _P]),
    "dlc_env_agent_rollout_f32": (ctypes.c_int32, [ctypes.POINTER(DlcEnvAgentParams)] + [_P] * 11),
    "dlc_env_collect_f32": (ctypes.c_int32, [ctypes.POINTER(DlcEnvSpec), ctypes.c_int64, ctypes.c_int32,
                                             ctypes.c_uint64, ctypes.c_int64, ctypes.c_float, _P, _P, _P, _P, _P]),
    "dlc_microbench_f32": (ctypes.c_int32, [ctypes.c_int32, _P, ctypes.c_int32, ctypes.c_int32, _P]),
    "dlc_copy_rows_async": (ctypes.c_int32, [_P, ctypes.c_size_t, _P, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
                                             ctypes.c_int32, _P]),
    "dlc_lds_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(DlcLdsParams)]),
    "dlc_lds_rollout_f32": (ctypes.c_int32, [ctypes.POINTER(DlcLdsParams)] + [_P] * 18
                            + [_P, ctypes.c_size_t, _P]),
    "dlc_lds_rollout_cost_f32": (ctypes.c_int32, [ctypes.POINTER(DlcLdsParams)] + [_P] * 16
                                 + [_P, ctypes.c_size_t, _P]),
}


def lib():
    """Load the shared library once.  Raises if it has not been built (``make -C deluca_b200/csrc``
    or ``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} not found: build it with `make -C deluca_b200/csrc` "
                    "(there is no CPU fallback)")
            L = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            for which, struct in enumerate((DlcLdsParams, DlcLungParams, DlcLungBuffers, DlcPlanParams,
                                            DlcSharedGpcParams, DlcOfParams)):
                if L.dlc_sizeof(which) != ctypes.sizeof(struct):
                    raise ImportError(f"{struct.__name__}: ctypes layout ({ctypes.sizeof(struct)} B) does not match "
                                      f"the library ({L.dlc_sizeof(which)} B); rebuild libdeluca_b200.so")
            _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().dlc_last_error().decode("utf-8", "replace")
        raise DlcError(f"{what} failed with code {rc}: {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def current_stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
