"""ctypes binding of libphe_b200.so (the C ABI in include/phe_b200.h).

There is no fallback: if the library is missing or CUDA is unusable, every call raises.
ctypes releases the GIL for the duration of each foreign call, so callers on many Python
threads (env/workflow.py:848-878 runs game loops as threads) overlap their launches.
"""
import ctypes as C
import os

from . import build as _build

_lib = None

PHE_OK = 0
LAYOUT_IDS = 0
LAYOUT_MASKS = 1
NUM_RANKS = 7462
NUM_HANDS_7 = 133784560

# every symbol include/phe_b200.h declares
SYMBOLS = [
    "phe_strerror", "phe_last_cuda_error", "phe_version", "phe_init",
    "phe_eval7", "phe_eval7_host", "phe_enum7", "phe_enum7_host", "phe_enum7_part", "phe_enum7_allreduce_workspace", "phe_enum7_allreduce", "phe_enum7_allreduce_nvls", "phe_enum7_allreduce_status",
    "phe_equity_mc", "phe_equity_mc_host", "phe_equity_plan", "phe_equity_plan_host", "phe_plan_num_deals",
    "phe_showdown", "phe_showdown_host", "phe_deal", "phe_deal_host", "phe_deal_samples", "phe_deal_samples_host", "phe_settle", "phe_settle_host", "phe_launch_count",
    "phe_nn_add_layernorm", "phe_nn_embed_obs", "phe_nn_attention", "phe_nn_policy_value_heads", "phe_obs_num_fields", "phe_obs_cut_table_bytes", "phe_encode_obs",
]


class PheError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PHE_B200_LIB") or _build.LIB        # the override selects an experiment build (tools/lab/mc_variants.py)
    if not os.path.exists(path):
        # built artefacts are not in git; build in-tree when a compiler is present, else fail loudly
        try:
            _build.build_library()
        except Exception as e:  # noqa: BLE001
            raise PheError(f"{path} is missing and could not be built ({e}); run __graft_entry__.build()") from e
    L = C.CDLL(path)
    vp, i64, u64, u32, i32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32, C.c_int
    L.phe_strerror.argtypes = [i32]; L.phe_strerror.restype = C.c_char_p
    L.phe_last_cuda_error.argtypes = []; L.phe_last_cuda_error.restype = C.c_char_p
    L.phe_version.restype = i32
    L.phe_init.argtypes = [i32]; L.phe_init.restype = i32
    L.phe_eval7.argtypes = [vp, i32, i64, vp, vp]; L.phe_eval7.restype = i32
    L.phe_eval7_host.argtypes = [vp, i32, i64, vp]; L.phe_eval7_host.restype = i32
    L.phe_enum7.argtypes = [vp, vp, u64, u64, vp]; L.phe_enum7.restype = i32
    L.phe_enum7_host.argtypes = [vp, vp, u64, u64]; L.phe_enum7_host.restype = i32
    L.phe_enum7_part.argtypes = [vp, vp, u64, u64, u32, u32, vp]; L.phe_enum7_part.restype = i32
    L.phe_enum7_allreduce_workspace.argtypes = []; L.phe_enum7_allreduce_workspace.restype = u64
    L.phe_enum7_allreduce.argtypes = [vp, u64, u64, i32, i32, vp, vp]; L.phe_enum7_allreduce.restype = i32
    L.phe_enum7_allreduce_nvls.argtypes = [vp, u64, u64, i32, i32, vp, vp, vp]; L.phe_enum7_allreduce_nvls.restype = i32
    L.phe_enum7_allreduce_status.argtypes = [vp, vp, vp, vp]; L.phe_enum7_allreduce_status.restype = i32
    L.phe_equity_mc.argtypes = [vp, i64, i32, u64, u64, u64, u32, vp, vp]; L.phe_equity_mc.restype = i32
    L.phe_equity_mc_host.argtypes = [vp, i64, i32, u64, u64, u64, u32, vp]; L.phe_equity_mc_host.restype = i32
    L.phe_equity_plan.argtypes = [vp, i32, i64, vp, i32, u64, u32, vp, vp]; L.phe_equity_plan.restype = i32
    L.phe_equity_plan_host.argtypes = [vp, i32, i64, vp, i32, u64, u32, vp]; L.phe_equity_plan_host.restype = i32
    L.phe_plan_num_deals.argtypes = [i32, vp, i32]; L.phe_plan_num_deals.restype = u64
    L.phe_showdown.argtypes = [vp, vp, vp, i64, i32, vp, vp, vp]; L.phe_showdown.restype = i32
    L.phe_showdown_host.argtypes = [vp, vp, vp, i64, i32, vp, vp]; L.phe_showdown_host.restype = i32
    L.phe_deal.argtypes = [vp, i64, i32, u64, u64, u32, vp, vp]; L.phe_deal.restype = i32
    L.phe_deal_host.argtypes = [vp, i64, i32, u64, u64, u32, vp]; L.phe_deal_host.restype = i32
    L.phe_deal_samples.argtypes = [vp, i64, i32, u64, u64, u32, vp, vp]; L.phe_deal_samples.restype = i32
    L.phe_deal_samples_host.argtypes = [vp, i64, i32, u64, u64, u32, vp]; L.phe_deal_samples_host.restype = i32
    L.phe_settle.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, vp, vp, vp, vp, vp]; L.phe_settle.restype = i32
    L.phe_settle_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, vp, vp, vp, vp]; L.phe_settle_host.restype = i32
    L.phe_launch_count.restype = u64
    L.phe_nn_add_layernorm.argtypes = [vp, vp, vp, vp, vp, i64, i32, C.c_float, i32, vp]; L.phe_nn_add_layernorm.restype = i32
    L.phe_obs_num_fields.argtypes = [i32, i32, i32]; L.phe_obs_num_fields.restype = i32
    L.phe_obs_cut_table_bytes.argtypes = []; L.phe_obs_cut_table_bytes.restype = u64
    L.phe_encode_obs.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, vp, vp]; L.phe_encode_obs.restype = i32
    L.phe_nn_attention.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32, i32, i64, i64, i64, i64, i32, vp]; L.phe_nn_attention.restype = i32
    L.phe_nn_policy_value_heads.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, i32, i64, i32, vp]; L.phe_nn_policy_value_heads.restype = i32
    L.phe_nn_embed_obs.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, i32, vp]; L.phe_nn_embed_obs.restype = i32
    _lib = L
    return L


def check(rc):
    if rc != PHE_OK:
        L = lib()
        msg = L.phe_strerror(rc).decode()
        if rc == -2:
            msg += ": " + L.phe_last_cuda_error().decode()
        raise PheError(f"phe_b200 status {rc}: {msg}")


def launch_count():
    return int(lib().phe_launch_count())
