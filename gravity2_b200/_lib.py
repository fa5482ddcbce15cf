"""ctypes binding of libgravity_b200.so (C ABI: include/gravity2_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C gravity2_b200/csrc``.
There is no CPU fallback: a missing library or a machine without a B200 raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GV2_LIB", os.path.join(_HERE, "libgravity_b200.so"))

OK = 0
ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED, ERR_ZERODIV = -1, -2, -3, -4, -5
SCHEMES = {"P": 0, "G": 1, "PG": 2, "R": 3, "RG": 4, "PR": 5, "L": 6, "PL": 7}
OUT_SIMILARITY, OUT_DISTANCE = 0, 1
TILE = 128
GOM_ALL_COLUMNS = 1
TIME_PAIRSUMS, TIME_TAIL, TIME_GOM, TIME_LINKAGE, TIME_BASE = 0, 1, 2, 3, 4
PATH_NAMES = {0: "none", 1: "fp32", 2: "mma", 3: "sparse"}

c_i64 = C.c_int64
c_dp = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_vp = C.c_void_p


class GjComponent(C.Structure):
    """gv2_gj_component"""
    _fields_ = [("min_sums", c_vp), ("rowsums", c_vp), ("nanflags", c_vp), ("ksplit", C.c_int32),
                ("replicate_stride_sums", c_i64), ("replicate_stride_rows", c_i64),
                ("replicate_stride_nan", c_i64)]


# every symbol include/gravity2_b200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "gv2_version": (C.c_int, []),
    "gv2_device_count": (C.c_int, []),
    "gv2_create": (C.c_int, [C.c_int, C.POINTER(c_vp)]),
    "gv2_create_on_stream": (C.c_int, [C.c_int, c_vp, C.POINTER(c_vp)]),
    "gv2_destroy": (None, [c_vp]),
    "gv2_last_error": (C.c_char_p, [c_vp]),
    "gv2_synchronize": (C.c_int, [c_vp]),
    "gv2_launch_count": (C.c_uint64, [c_vp]),
    "gv2_timing_enable": (C.c_int, [c_vp, C.c_int]),
    "gv2_timing_read": (C.c_int, [c_vp, C.POINTER(C.c_double), C.POINTER(c_i64)]),
    "gv2_timing_read_kind": (C.c_int, [c_vp, C.c_int, C.POINTER(C.c_double), C.POINTER(c_i64)]),
    "gv2_job_replicate_path": (C.c_int, [c_vp]),
    "gv2_job_copresent_cells": (C.c_uint64, [c_vp]),
    "gv2_similarity_host": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, C.c_int,
                                      C.c_double, C.c_double, C.c_int, C.c_int, c_vp]),
    "gv2_similarity_nbhd_host": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, C.c_double, C.c_double, c_vp, c_vp]),
    "gv2_linkage_host": (C.c_int, [c_vp, c_vp, c_i64, C.c_char_p, c_vp]),
    "gv2_newick_from_linkage": (C.c_int, [c_vp, c_i64, C.POINTER(C.c_char_p), c_vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "gv2_job_replicates_newick": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, C.c_char_p, C.POINTER(C.c_char_p), c_vp, c_vp]),
    "gv2_gom_table_host": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64,
                                     C.c_int, c_vp]),
    "gv2_shared_counts_host": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "gv2_bootstrap_host": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, C.c_int,
                                     C.c_double, C.c_int, c_vp]),
    "gv2_philox_indices_host": (C.c_int, [c_vp, C.c_uint64, c_i64, c_i64, c_vp]),
    "gv2_dev_philox_weights": (C.c_int, [c_vp, C.c_uint64, c_i64, c_i64, c_i64, c_vp]),
    "gv2_job_create": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, c_i64, c_i64, c_vp, c_i64, c_i64, C.c_int, C.c_double,
                                 C.c_int, C.POINTER(c_vp)]),
    "gv2_job_create_coo": (C.c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, C.c_int,
                                     C.c_double, C.c_int, C.POINTER(c_vp)]),
    "gv2_job_destroy": (None, [c_vp]),
    "gv2_job_base": (C.c_int, [c_vp, c_i64, c_i64, C.c_int, c_vp]),
    "gv2_job_replicates": (C.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "gv2_job_replicates_stream": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "gv2_job_replicates_condensed": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "gv2_job_replicates_device": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "gv2_dev_alloc": (C.c_int, [c_vp, C.c_size_t, C.POINTER(c_vp)]),
    "gv2_dev_free": (C.c_int, [c_vp, c_vp]),
    "gv2_host_alloc": (C.c_int, [c_vp, C.c_size_t, C.POINTER(c_vp)]),
    "gv2_host_free": (C.c_int, [c_vp, c_vp]),
    "gv2_memcpy_h2d": (C.c_int, [c_vp, c_vp, c_vp, C.c_size_t]),
    "gv2_memcpy_d2h": (C.c_int, [c_vp, c_vp, c_vp, C.c_size_t]),
    "gv2_mem_info": (C.c_int, [c_vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "gv2_job_gom_table": (C.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "gv2_job_device_bytes": (C.c_size_t, [c_vp]),
    "gv2_job_gom_stats": (C.c_int, [c_vp, C.POINTER(c_i64)]),
    "gv2_dev_shared_counts_packed": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp]),
    "gv2_dev_unpack_tiles_i32": (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "gv2_dev_popc_peak": (C.c_int, [c_vp, C.c_int, C.POINTER(C.c_double)]),
    "gv2_multi_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(c_vp)]),
    "gv2_multi_destroy": (None, [c_vp]),
    "gv2_multi_last_error": (C.c_char_p, [c_vp]),
    "gv2_multi_device_count": (C.c_int, [c_vp]),
    "gv2_multi_uses_nccl": (C.c_int, [c_vp]),
    "gv2_multi_context": (c_vp, [c_vp, C.c_int]),
    "gv2_multi_job_create": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, C.c_int, C.c_double, C.c_int, C.POINTER(c_vp)]),
    "gv2_multi_job_destroy": (None, [c_vp]),
    "gv2_multi_job_base_host": (C.c_int, [c_vp, c_vp]),
    "gv2_multi_job_replicates_newick": (C.c_int, [c_vp, c_vp, c_i64, C.c_char_p, C.POINTER(C.c_char_p), c_i64, c_vp, c_vp]),
    "gv2_multi_job_replicates_host": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "gv2_dev_rowmax_block": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "gv2_dev_gather_cells": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp]),
    "gv2_num_tiles": (c_i64, [c_i64]),
    "gv2_dev_prepare_table": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, C.c_double, C.c_int, c_vp, c_i64, c_i64,
                                        c_vp, c_vp]),
    "gv2_dev_weights_to_float": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp]),
    "gv2_dev_weighted_rowsums": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "gv2_dev_gj_min_sums": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, C.c_int, c_vp]),
    "gv2_dev_gj_epilogue": (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, C.POINTER(GjComponent),
                                      C.POINTER(GjComponent), c_vp, C.c_int, C.c_double, C.c_int, C.c_int, c_vp,
                                      c_i64]),
    "gv2_dev_gj_fused": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, C.POINTER(GjComponent), C.POINTER(GjComponent),
                                   C.c_int, C.c_double, C.c_int, c_vp, c_i64]),
    "gv2_dev_unpack_tiles": (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "gv2_dev_pack_presence": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp]),
    "gv2_dev_shared_counts": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp]),
}

_lib = None


REPLICATE_CB = C.CFUNCTYPE(C.c_int, c_i64, c_vp, c_vp)      # gv2_replicate_cb
NEWICK_CB = C.CFUNCTYPE(C.c_int, c_i64, c_vp, C.c_size_t, c_vp)   # gv2_newick_cb


class GravityNativeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libgravity_b200 error {code}: {msg}")
        self.code = code


def load() -> C.CDLL:
    """Load the shared library and bind every prototype.  Loading touches no CUDA API."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C gravity2_b200/csrc`).  gravity2_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(ctx, rc: int) -> None:
    if rc != OK:
        msg = load().gv2_last_error(ctx)
        raise GravityNativeError(rc, msg.decode() if msg else "?")
