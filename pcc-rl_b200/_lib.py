"""ctypes binding of include/pcc_b200.h (libpcc_b200.so).  No CPU fallback: if the library or a
CUDA device is missing, loading / pcc_create fails loudly."""
import ctypes as C
import os

from . import build as _build

PCC_ABI_VERSION = 2
PCC_OK, PCC_EINVAL, PCC_ECUDA, PCC_EOVERFLOW, PCC_ENODEV = 0, -1, -2, -3, -4
PCC_RNG_MT19937, PCC_RNG_PHILOX = 0, 1
PCC_MAX_FEATURES = 12
PCC_INFO_WIDTH = 12


class PccConsts(C.Structure):
    _fields_ = [("max_rate", C.c_double), ("min_rate", C.c_double), ("delta_scale", C.c_double),
                ("reward_scale", C.c_double), ("max_steps", C.c_int32), ("bytes_per_packet", C.c_int32)]


class PccConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("n_envs", C.c_int64),
                ("history_len", C.c_int32), ("n_features", C.c_int32),
                ("feature_ids", C.c_int32 * PCC_MAX_FEATURES), ("rng_kind", C.c_int32),
                ("reserved0", C.c_int32), ("ring_capacity", C.c_int64), ("consts", PccConsts)]


class PccPolicy(C.Structure):
    _fields_ = [("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
                ("w3", C.c_void_p), ("b3", C.c_void_p), ("n_in", C.c_int32), ("h1", C.c_int32),
                ("h2", C.c_int32), ("stochastic", C.c_int32), ("log_std", C.c_double), ("noise_seed", C.c_uint64),
                ("vw1", C.c_void_p), ("vb1", C.c_void_p), ("vw2", C.c_void_p), ("vb2", C.c_void_p),
                ("vw3", C.c_void_p), ("vb3", C.c_void_p)]


class PccVariant(C.Structure):
    _fields_ = [("use_cwnd", C.c_int32), ("use_latency_noise", C.c_int32), ("max_latency_noise", C.c_double),
                ("initial_cwnd", C.c_int32), ("min_cwnd", C.c_int32), ("max_cwnd", C.c_int32), ("reserved0", C.c_int32)]


class PccFlowsConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("n_flows", C.c_int64),
                ("history_len", C.c_int32), ("n_features", C.c_int32),
                ("feature_ids", C.c_int32 * PCC_MAX_FEATURES), ("delta_scale", C.c_double),
                ("min_rate", C.c_double), ("max_rate", C.c_double), ("rate_style", C.c_int32),
                ("reserved0", C.c_int32)]


class PccMiBatch(C.Structure):
    _fields_ = [("n_records", C.c_int64), ("flow", C.c_void_p), ("bytes_sent", C.c_void_p),
                ("bytes_acked", C.c_void_p), ("bytes_lost", C.c_void_p), ("packet_size", C.c_void_p),
                ("send_start", C.c_void_p), ("send_end", C.c_void_p), ("recv_start", C.c_void_p),
                ("recv_end", C.c_void_p), ("rtt_offsets", C.c_void_p), ("rtt_samples", C.c_void_p)]


PCC_RATE_CLIENT, PCC_RATE_SHIM = 0, 1
PCC_FLOW_RESET_NEW, PCC_FLOW_RESET_CLIENT, PCC_FLOW_RESET_SHIM = 0, 1, 2
PCC_N_METRICS = 12


class PccError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "libpcc_b200 error %d: %s" % (code, msg))
        self.code = code


# every symbol include/pcc_b200.h declares
EXPORTS = ["pcc_default_consts", "pcc_default_config", "pcc_ring_capacity_for", "pcc_workspace_bytes",
           "pcc_create", "pcc_destroy", "pcc_attach", "pcc_seed", "pcc_get_mt_state", "pcc_set_mt_state",
           "pcc_reset", "pcc_step", "pcc_step_host", "pcc_step_host_submit", "pcc_step_host_wait", "pcc_rollout", "pcc_check", "pcc_get_column", "pcc_launch_count",
           "pcc_last_error", "pcc_abi_version", "pcc_multi_workspace_bytes", "pcc_multi_create", "pcc_multi_destroy",
           "pcc_multi_seed", "pcc_multi_reset", "pcc_multi_step", "pcc_multi_check", "pcc_multi_launch_count",
           "pcc_default_variant", "pcc_multi_set_variant", "pcc_multi_step_cwnd",
           "pcc_flows_default_config", "pcc_flows_workspace_bytes", "pcc_flows_create", "pcc_flows_attach",
           "pcc_flows_destroy", "pcc_flows_give_samples", "pcc_flows_reset", "pcc_flows_get_obs", "pcc_flows_set_rates",
           "pcc_flows_get_rates", "pcc_flows_act", "pcc_flows_get_column", "pcc_flows_check", "pcc_flows_launch_count"]

_lib = None


def load(rebuild_if_stale=True):
    """Loads libpcc_b200.so (building it with nvcc first if the sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PCC_B200_LIB") or _build.LIB   # override: experiments with alternative builds
    if path == _build.LIB and rebuild_if_stale and not _build.up_to_date():
        try:
            _build.build()
        except Exception as ex:
            if not os.path.exists(path):
                raise
            import warnings
            warnings.warn("libpcc_b200.so is older than its sources and rebuilding failed (%s): loading the STALE "
                          "library" % str(ex).splitlines()[0])
    if not os.path.exists(path):
        raise RuntimeError("libpcc_b200.so is missing (run `python __graft_entry__.py build`); "
                           "there is no CPU fallback")
    L = C.CDLL(path)
    vp, u8p, dp = C.c_void_p, C.c_void_p, C.c_void_p
    L.pcc_default_consts.argtypes = [C.POINTER(PccConsts)]
    L.pcc_default_consts.restype = None
    L.pcc_default_config.argtypes = [C.POINTER(PccConfig)]
    L.pcc_default_config.restype = None
    L.pcc_ring_capacity_for.argtypes = [C.c_double] * 4
    L.pcc_ring_capacity_for.restype = C.c_int64
    L.pcc_workspace_bytes.argtypes = [C.POINTER(PccConfig), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.pcc_create.argtypes = [C.POINTER(vp), C.POINTER(PccConfig), vp, vp]
    L.pcc_attach.argtypes = [C.POINTER(vp), C.POINTER(PccConfig), vp, vp]
    L.pcc_destroy.argtypes = [vp]
    L.pcc_destroy.restype = None
    L.pcc_seed.argtypes = [vp, vp, u8p, vp]
    L.pcc_get_mt_state.argtypes = [vp, C.c_int64, C.POINTER(C.c_uint32)]
    L.pcc_set_mt_state.argtypes = [vp, C.c_int64, C.POINTER(C.c_uint32)]
    L.pcc_reset.argtypes = [vp, u8p, dp, dp, vp, dp, dp, dp, vp]
    L.pcc_step.argtypes = [vp, dp, dp, dp, u8p, vp, dp, vp]
    L.pcc_step_host.argtypes = [vp, dp, dp, dp, u8p, vp, vp]
    L.pcc_step_host_submit.argtypes = [vp, dp, dp, dp, u8p, vp, dp, vp, C.POINTER(C.c_int64)]
    L.pcc_step_host_wait.argtypes = [vp, C.c_int64]
    L.pcc_rollout.argtypes = [vp, C.c_int32, dp, C.POINTER(PccPolicy), dp, C.c_int32, dp, dp, dp, u8p, vp, dp, vp]
    L.pcc_check.argtypes = [vp, vp]
    L.pcc_multi_workspace_bytes.argtypes = [C.POINTER(PccConfig), C.c_int32, C.POINTER(C.c_uint64)]
    L.pcc_multi_create.argtypes = [C.POINTER(vp), C.POINTER(PccConfig), C.c_int32, vp]
    L.pcc_multi_destroy.argtypes = [vp]
    L.pcc_multi_destroy.restype = None
    L.pcc_multi_seed.argtypes = [vp, vp, vp]
    L.pcc_multi_reset.argtypes = [vp, u8p, dp, dp, vp, dp, dp, dp, vp]
    L.pcc_multi_step.argtypes = [vp, dp, dp, dp, u8p, vp, vp]
    L.pcc_multi_check.argtypes = [vp, vp]
    L.pcc_multi_launch_count.argtypes = [vp]
    L.pcc_multi_launch_count.restype = C.c_int64
    L.pcc_default_variant.argtypes = [C.POINTER(PccVariant)]
    L.pcc_default_variant.restype = None
    L.pcc_multi_set_variant.argtypes = [vp, C.POINTER(PccVariant)]
    L.pcc_multi_step_cwnd.argtypes = [vp, dp, dp, dp, dp, u8p, vp, vp, vp]
    L.pcc_flows_default_config.argtypes = [C.POINTER(PccFlowsConfig)]
    L.pcc_flows_default_config.restype = None
    L.pcc_flows_workspace_bytes.argtypes = [C.POINTER(PccFlowsConfig), C.POINTER(C.c_uint64)]
    L.pcc_flows_create.argtypes = [C.POINTER(vp), C.POINTER(PccFlowsConfig), vp]
    L.pcc_flows_attach.argtypes = [C.POINTER(vp), C.POINTER(PccFlowsConfig), vp]
    L.pcc_flows_destroy.argtypes = [vp]
    L.pcc_flows_destroy.restype = None
    L.pcc_flows_give_samples.argtypes = [vp, C.POINTER(PccMiBatch), C.c_int32, dp, dp, vp]
    L.pcc_flows_reset.argtypes = [vp, u8p, C.c_int32, vp]
    L.pcc_flows_get_obs.argtypes = [vp, dp, vp]
    L.pcc_flows_set_rates.argtypes = [vp, u8p, dp, C.c_double, vp]
    L.pcc_flows_get_rates.argtypes = [vp, dp, u8p, dp, vp]
    L.pcc_flows_act.argtypes = [vp, C.POINTER(PccPolicy), dp, vp]
    L.pcc_flows_get_column.argtypes = [vp, C.c_char_p, dp, vp]
    L.pcc_flows_check.argtypes = [vp, vp]
    L.pcc_flows_launch_count.argtypes = [vp]
    L.pcc_flows_launch_count.restype = C.c_int64
    L.pcc_get_column.argtypes = [vp, C.c_char_p, dp, vp]
    L.pcc_launch_count.argtypes = [vp]
    L.pcc_launch_count.restype = C.c_int64
    L.pcc_last_error.restype = C.c_char_p
    L.pcc_abi_version.restype = C.c_int
    if L.pcc_abi_version() != PCC_ABI_VERSION:
        raise RuntimeError("libpcc_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise PccError(rc, load().pcc_last_error().decode("utf-8", "replace"))
