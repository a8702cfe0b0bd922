"""ctypes binding of liblmc_b200.so (include/lmc_b200.h).  Fails loudly when the CUDA library is missing: there
is no CPU fallback anywhere in this package."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LMC_LIB_PATH") or os.path.join(_HERE, "liblmc_b200.so")  # env: kernel experiments only

ABI_VERSION = 3
OK, ERR_BADARG, ERR_UNSUPPORTED, ERR_LAUNCH, ERR_WORKSPACE = 0, -1, -2, -3, -4
TARGET_DIAG_GAUSSIAN, TARGET_FUNNEL = 0, 1
RNG_TAPE, RNG_PHILOX = 0, 1
KIND_NUTS, KIND_HMC = 0, 1
ADAPT_LOG_STEP, ADAPT_LOG_BAR, ADAPT_HBAR, ADAPT_COUNT, ADAPT_MU = 0, 1, 2, 3, 4
ADAPT_W_FG, ADAPT_W_BG, ADAPT_NSAMPLES, ADAPT_WINDOW, ADAPT_STRIDE = 5, 6, 7, 8, 10
NSTATS = 13
(STAT_DEPTH, STAT_TREE_SIZE, STAT_ACCEPT, STAT_ENERGY, STAT_ENERGY_ERROR, STAT_MAX_ENERGY_ERROR, STAT_MODEL_LOGP,
 STAT_DIVERGING, STAT_TUNE, STAT_STEP_SIZE, STAT_STEP_SIZE_BAR, STAT_N_UNIFORMS,
 STAT_REACHED_MAX_TREEDEPTH) = range(13)
STATUS_BAD_INITIAL_ENERGY, STATUS_TAPE_EXHAUSTED = 1, 2

_ERR_NAMES = {ERR_BADARG: "LMC_ERR_BADARG", ERR_UNSUPPORTED: "LMC_ERR_UNSUPPORTED", ERR_LAUNCH: "LMC_ERR_LAUNCH",
              ERR_WORKSPACE: "LMC_ERR_WORKSPACE"}


class Target(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("tau", C.c_void_p), ("v_scale", C.c_double)]


class Rng(C.Structure):
    _fields_ = [("mode", C.c_int32), ("reserved", C.c_int32), ("normals", C.c_void_p), ("uniforms", C.c_void_p),
                ("u_stride", C.c_int64), ("seeds", C.c_void_p)]


class SamplerArgs(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_chains", C.c_int32), ("ndim", C.c_int32), ("reserved0", C.c_int32),
        ("ld", C.c_int64), ("target", Target),
        ("q", C.c_void_p), ("var", C.c_void_p),
        ("adapt_mass", C.c_int32), ("adapt_step_size", C.c_int32),
        ("mean_fg", C.c_void_p), ("rawvar_fg", C.c_void_p), ("mean_bg", C.c_void_p), ("rawvar_bg", C.c_void_p),
        ("adapt", C.c_void_p), ("window_multiplier", C.c_double),
        ("target_accept", C.c_double), ("gamma", C.c_double), ("k", C.c_double), ("t0", C.c_double),
        ("iter0", C.c_int64), ("n_tune", C.c_int64), ("n_trans", C.c_int32), ("reserved1", C.c_int32),
        ("Emax", C.c_double), ("max_treedepth", C.c_int32), ("early_max_treedepth", C.c_int32),
        ("path_length", C.c_double), ("max_steps", C.c_int32), ("reserved2", C.c_int32),
        ("rng", Rng),
        ("trace", C.c_void_p), ("trace_chain_stride", C.c_int64), ("trace_draw_stride", C.c_int64),
        ("stats", C.c_void_p), ("status", C.c_void_p), ("step_size_override", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("stream", C.c_void_p),
        ("tune_group", C.c_int32), ("tune_smem_vecs", C.c_int32), ("tune_max_slots", C.c_int32),
        ("tune_chunk", C.c_int32),
        ("trace_skip", C.c_int32), ("progress_block", C.c_int32), ("progress", C.c_void_p),
    ]


class CallbackArgs(C.Structure):
    _fields_ = [("base", SamplerArgs), ("q_eval", C.c_void_p), ("g_eval", C.c_void_p), ("logp_eval", C.c_void_p),
                ("machine", C.c_void_p), ("machine_bytes", C.c_int64), ("n_running", C.c_void_p)]


class DenseArgs(C.Structure):
    _fields_ = [("base", SamplerArgs), ("q_eval", C.c_void_p), ("g_eval", C.c_void_p), ("logp_eval", C.c_void_p),
                ("x_eval", C.c_void_p), ("v_eval", C.c_void_p), ("n_eval", C.c_void_p), ("p0_eval", C.c_void_p),
                ("need", C.c_void_p), ("machine", C.c_void_p), ("machine_bytes", C.c_int64), ("n_running", C.c_void_p)]


NEED_GRAD, NEED_VEL, NEED_MOM, NEED_UPDATE, NEED_HOLD = 1, 2, 4, 8, 16

# every symbol include/lmc_b200.h declares: (restype, argtypes)
_P, _I32, _I64 = C.c_void_p, C.c_int32, C.c_int64
SYMBOLS = {
    "lmc_abi_version": (C.c_int, []),
    "lmc_workspace_bytes": (_I64, [_I32, _I32, _I32, _I32, _I32]),
    "lmc_nuts_sample": (C.c_int, [C.POINTER(SamplerArgs)]),
    "lmc_hmc_sample": (C.c_int, [C.POINTER(SamplerArgs)]),
    "lmc_callback_state_bytes": (_I64, [_I32, _I32, _I32, _I32]),
    "lmc_callback_begin": (C.c_int, [_I32, C.POINTER(CallbackArgs)]),
    "lmc_callback_advance": (C.c_int, [_I32, C.POINTER(CallbackArgs)]),
    "lmc_callback_loop_create": (C.c_int, [_P, _P, _P, _I64, C.POINTER(_P)]),
    "lmc_callback_loop_launch": (C.c_int, [_P, _P]),
    "lmc_callback_loop_destroy": (C.c_int, [_P]),
    "lmc_user_kernel_build": (C.c_int, [C.c_char_p, C.c_char_p, _I32, _I32, _I32, _I32, C.POINTER(C.c_char_p), _I32,
                                        C.c_char_p, C.POINTER(_P)]),
    "lmc_user_kernel_log": (C.c_char_p, []),
    "lmc_user_sample": (C.c_int, [_P, C.POINTER(SamplerArgs), _P]),
    "lmc_user_kernel_destroy": (C.c_int, [_P]),
    "lmc_dense_state_bytes": (_I64, [_I32, _I32, _I32, _I32]),
    "lmc_dense_begin": (C.c_int, [_I32, C.POINTER(DenseArgs)]),
    "lmc_dense_advance": (C.c_int, [_I32, C.POINTER(DenseArgs)]),
    "lmc_dense_matvec": (C.c_int, [_P, _I32, _P, _I64, _I64, _I32, _I64, _P, _P, _I32, _P]),
    "lmc_dense_cov_update": (C.c_int, [_P, _I32, _I32, _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "lmc_chain_moments": (C.c_int, [_P, _I32, _I32, _I32, _I64, _I64, _I32, _P, _P, _P]),
    "lmc_compute_state": (C.c_int, [C.POINTER(Target), _I32, _I32, _I64, _P, _P, _P, _I64, _P, _P, _P, _P, _P]),
    "lmc_leapfrog_step": (C.c_int, [C.POINTER(Target), _I32, _I32, _I64, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P,
                                    _P, _P, _P]),
    "lmc_leapfrog_half1": (C.c_int, [_I32, _I32, _I64, _P, _P, _P, _P, _P, _P, _I64, _P]),
    "lmc_leapfrog_half2": (C.c_int, [_I32, _I32, _I64, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P]),
    "lmc_rng_fill": (C.c_int, [_P, _I32, _I32, _I64, _I32, _I64, _P, _P, _P]),
    "lmc_memcpy2d_d2h": (C.c_int, [_P, _I64, _P, _I64, _I64, _I64, _P]),
    "lmc_last_error": (C.c_char_p, []),
}

_lib = None


class LmcError(RuntimeError):
    pass


def load():
    """Load liblmc_b200.so (built in-tree by __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LmcError(
            "littlemcmc_b200: %s not found. Build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    if lib.lmc_abi_version() != ABI_VERSION:
        raise LmcError("ABI mismatch: library %d, binding %d" % (lib.lmc_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc, what):
    if rc != OK:
        msg = load().lmc_last_error().decode() if rc == ERR_LAUNCH else ""
        raise LmcError("%s failed: %s %s" % (what, _ERR_NAMES.get(rc, rc), msg))
