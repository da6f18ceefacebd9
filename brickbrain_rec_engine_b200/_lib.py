"""ctypes binding of libbrickscore.so (the C ABI declared in include/brickscore.h).

The library is built in-tree by ``brickbrain_rec_engine_b200.build``.  There is
no Python / CPU fallback: if the shared object is missing, ``load()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BRICKSCORE_LIB") or os.path.join(HERE, "libbrickscore.so")   # (override: A/B runs of two builds)

BB_OK = 0
BB_ERR_INVALID = -22
BB_ERR_NOMEM = -12
BB_ERR_CUDA = -5
BB_ERR_STATE = -1
BB_ERR_UNSUPPORTED = -95

BB_METRIC_COSINE = 0
BB_METRIC_IP = 1
BB_CFG_KEEP_FP32 = 1
BB_CFG_BUILD_BF16 = 2
BB_CFG_BUILD_SPLIT = 4
BB_PREC_FP32 = 0
BB_PREC_BF16 = 1
BB_PREC_BF16_RESCORE = 2
BB_Q_SCORES_AS_L2 = 1
BB_Q_SHARE_THRESHOLDS = 2
BB_Q_EMB_PRENORMALIZED = 4
BB_C_ALWAYS_FALSE = 1
BB_IPC_HANDLE_BYTES = 64

PRECISIONS = {"fp32": BB_PREC_FP32, "bf16": BB_PREC_BF16, "bf16_rescore": BB_PREC_BF16_RESCORE}
METRICS = {"cosine": BB_METRIC_COSINE, "ip": BB_METRIC_IP}

INT32_MAX = 2**31 - 1
INT32_MIN = -2**31


class bb_config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("n_items", C.c_int64),
                ("id_offset", C.c_int64), ("dim", C.c_int32), ("cf_dim", C.c_int32), ("k_max", C.c_int32),
                ("max_batch", C.c_int32), ("theme_bits", C.c_int32), ("metric", C.c_int32), ("flags", C.c_uint32)]


class bb_constraint(C.Structure):
    _fields_ = [("parts_min", C.c_int32), ("parts_max", C.c_int32), ("year_min", C.c_int16),
                ("year_max", C.c_int16), ("theme_set", C.c_int32), ("allow_bitmap", C.c_int32),
                ("excl_list", C.c_int32 * 3), ("flags", C.c_uint32)]


#: numpy view of ``bb_constraint`` (same layout, 36 bytes)
CONSTRAINT_DTYPE = np.dtype({
    "names": ["parts_min", "parts_max", "year_min", "year_max", "theme_set", "allow_bitmap", "excl_list", "flags"],
    "formats": ["<i4", "<i4", "<i2", "<i2", "<i4", "<i4", ("<i4", (3,)), "<u4"],
    "offsets": [0, 4, 8, 10, 12, 16, 20, 32],
    "itemsize": 36,
})
assert C.sizeof(bb_constraint) == CONSTRAINT_DTYPE.itemsize == 36


def no_constraint() -> np.ndarray:
    """A descriptor that accepts every row (no ``num_parts > 0`` either)."""
    c = np.zeros((), dtype=CONSTRAINT_DTYPE)
    c["parts_min"], c["parts_max"] = INT32_MIN, INT32_MAX
    c["year_min"], c["year_max"] = -32768, 32767
    c["theme_set"] = c["allow_bitmap"] = -1
    c["excl_list"] = -1
    return c


class bb_query_batch(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_queries", C.c_int32), ("q_emb", C.c_void_p),
                ("q_item", C.c_void_p), ("q_cf", C.c_void_p), ("w_content", C.c_void_p), ("w_cf", C.c_void_p),
                ("w_content_all", C.c_float), ("w_cf_all", C.c_float), ("constraints", C.c_void_p),
                ("n_constraints", C.c_int32), ("constraint_index", C.c_void_p), ("self_id", C.c_void_p), ("precision", C.c_int32),
                ("flags", C.c_uint32), ("allow_bitmaps", C.c_void_p), ("n_allow", C.c_int32),
                ("n_excl_per_query", C.c_int32), ("excl_lists", C.c_void_p)]


class bb_info(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("sm_count", C.c_int32), ("k_pad", C.c_int32),
                ("cta_group", C.c_int32), ("launches", C.c_int64), ("last_units", C.c_int64),
                ("last_candidates", C.c_int64), ("last_appends", C.c_int64), ("last_slow_chunks", C.c_int64),
                ("last_compactions", C.c_int64), ("mma_wait_epilogue_cycles", C.c_int64),
                ("mma_wait_tma_cycles", C.c_int64), ("mma_total_cycles", C.c_int64), ("bytes_fp32", C.c_double), ("bytes_bf16", C.c_double),
                ("scan_ms", C.c_double), ("scan_launches", C.c_int64), ("epilogue_cycles", C.c_int64 * 7), ("round_cycles", C.c_int64 * 16), ("round_wait_cycles", C.c_int64 * 16),
                ("last_wasted_chunks", C.c_int64), ("last_flagged_rows", C.c_int64)]


#: every symbol include/brickscore.h declares: name -> (restype, argtypes)
_P = C.c_void_p
PROTOTYPES = {
    "bb_version": (C.c_char_p, []),
    "bb_last_error": (C.c_char_p, [_P]),
    "bb_create": (C.c_int, [C.POINTER(_P), C.POINTER(bb_config)]),
    "bb_destroy": (None, [_P]),
    "bb_get_info": (C.c_int, [_P, C.POINTER(bb_info)]),
    "bb_set_profiling": (C.c_int, [_P, C.c_int]),
    "bb_load_catalog": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, _P]),
    "bb_fetch_embeddings": (C.c_int, [_P, C.c_int32, _P, _P, _P]),
    "bb_set_exclusion_lists": (C.c_int, [_P, C.c_int64, _P, _P]),
    "bb_set_theme_sets": (C.c_int, [_P, C.c_int32, _P]),
    "bb_set_allow_bitmaps": (C.c_int, [_P, C.c_int32, _P]),
    "bb_search": (C.c_int, [_P, C.POINTER(bb_query_batch), C.c_int32, _P, _P, _P]),
    "bb_apply_constraints": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P]),
    "bb_hybrid_combine": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, C.c_int32, _P, _P, _P, C.c_int32, C.c_int32,
                                    _P, _P, _P]),
    "bb_merge_topk": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P, _P, C.c_int32, _P, _P, _P]),
    "bb_item_item_topk": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P, _P, C.c_int32, _P, _P, _P]),
    "bb_debug_scores": (C.c_int, [_P, C.POINTER(bb_query_batch), _P, _P]),
    "bb_csr_spmm_f64": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, _P, _P, _P, _P, C.c_int32, _P, _P]),
    "bb_comm_export": (C.c_int, [_P, _P]),
    "bb_comm_import": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "bb_comm_link_local": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "bb_search_sharded": (C.c_int, [_P, C.POINTER(bb_query_batch), C.c_int32, _P, _P, _P]),
    "bb_sharded_scatter": (C.c_int, [_P, C.POINTER(bb_query_batch), C.c_int32, _P]),
    "bb_sharded_collect": (C.c_int, [_P, _P, _P, _P]),
    "bb_sharded_scatter_after": (C.c_int, [_P, C.POINTER(bb_query_batch), C.c_int32, _P, _P, _P]),
}

_lib = None


class BrickscoreError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"brickscore error {code}: {message}")
        self.code = code


def load() -> C.CDLL:
    """Loads libbrickscore.so and installs the prototypes.  Raises if the CUDA
    extension has not been built -- the product path never falls back to CPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m brickbrain_rec_engine_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI and the header diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
