"""ctypes binding of libsvgt.so (the C ABI declared in include/svgt.h).

The library is built in-tree by `svtyper_b200.build.build_native()` (nvcc, sm_100a).
There is no CPU fallback: if the library is missing, or no CUDA device is visible, the
compute entry points raise.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVGT_LIB") or os.path.join(HERE, "libsvgt.so")   # SVGT_LIB: A/B builds of the same ABI

ABI_VERSION = 3
OK, ERR_ARG, ERR_CUDA, ERR_LOG_TABLE, ERR_LIB_INDEX, ERR_NO_DEVICE, ERR_RANGE = 0, -1, -2, -3, -4, -5, -6
ERR_NAMES = {ERR_ARG: "SVGT_ERR_ARG", ERR_CUDA: "SVGT_ERR_CUDA", ERR_LOG_TABLE: "SVGT_ERR_LOG_TABLE",
             ERR_LIB_INDEX: "SVGT_ERR_LIB_INDEX", ERR_NO_DEVICE: "SVGT_ERR_NO_DEVICE",
             ERR_RANGE: "SVGT_ERR_RANGE"}
VAR_DIRECT, VAR_BULK = 0, 1
VARIANTS = (0, 1)            # wide-row cross-check kernels; the default path is the compact one

# every symbol include/svgt.h declares (tests check the library exports all of them)
SYMBOLS = ("svgt_abi_version", "svgt_last_error", "svgt_device_count", "svgt_score_batch",
           "svgt_launches_per_batch", "svgt_set_variant", "svgt_ctx_create", "svgt_ctx_destroy",
           "svgt_ctx_score_host", "svgt_ctx_last_traffic", "svgt_ctx_last_kernel_ms",
           "svgt_score_compact", "svgt_ctx_score_host_compact", "svgt_shared_alloc", "svgt_shared_open",
           "svgt_shared_close", "svgt_shared_free", "svgt_wait_flags", "svgt_memcpy_d2h", "svgt_peer_copy", "svgt_set_flag",
           "svgt_plan_count", "svgt_plan_fill", "svgt_ctx_last_pieces")
PLAN_CHUNK_BYTES = 784
LAYOUT_SITE_ORDER = 1


class SvgtBatch(ctypes.Structure):
    """struct svgt_batch (include/svgt.h)."""
    _fields_ = [
        ("sites", ctypes.c_void_p), ("n_sites", ctypes.c_int64),
        ("frags", ctypes.c_void_p), ("n_frag", ctypes.c_int64),
        ("splits", ctypes.c_void_p), ("n_split", ctypes.c_int64),
        ("order", ctypes.c_void_p),
        ("lib_f64", ctypes.c_void_p), ("lib_i32", ctypes.c_void_p), ("n_lib", ctypes.c_int32),
        ("hist", ctypes.c_void_p), ("n_hist", ctypes.c_int64),
        ("pm", ctypes.c_void_p),
        ("logt", ctypes.c_void_p), ("n_log", ctypes.c_int64),
        ("consts", ctypes.c_void_p),
        ("min_aligned", ctypes.c_int32), ("split_slop", ctypes.c_int32),
        ("assoc_mode", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("split_weight", ctypes.c_double), ("disc_weight", ctypes.c_double),
    ]


class SvgtSegPlan(ctypes.Structure):
    """struct svgt_segplan (include/svgt.h): the piece plan of a compact batch."""
    _fields_ = [
        ("entries", ctypes.c_void_p), ("n_entries", ctypes.c_int64),
        ("pieces", ctypes.c_void_p), ("n_pieces", ctypes.c_int64),
        ("heavy", ctypes.c_void_p), ("n_heavy", ctypes.c_int64),
        ("scratch", ctypes.c_void_p), ("scratch_chunks", ctypes.c_int64),
    ]


class SvgtCBatch(ctypes.Structure):
    """struct svgt_cbatch (include/svgt.h): the compact schema."""
    _fields_ = [
        ("sites", ctypes.c_void_p), ("n_sites", ctypes.c_int64),
        ("rows", ctypes.c_void_p), ("n_rows", ctypes.c_int64),
        ("order", ctypes.c_void_p),
        ("lib_f64", ctypes.c_void_p), ("lib_i32", ctypes.c_void_p), ("n_lib", ctypes.c_int32),
        ("hist_max", ctypes.c_uint32),
        ("hist", ctypes.c_void_p), ("n_hist", ctypes.c_int64),
        ("pm", ctypes.c_void_p),
        ("logt", ctypes.c_void_p), ("n_log", ctypes.c_int64),
        ("consts", ctypes.c_void_p),
        ("min_aligned", ctypes.c_int32), ("split_slop", ctypes.c_int32),
        ("assoc_mode", ctypes.c_int32), ("unit_mode", ctypes.c_int32),
        ("split_weight", ctypes.c_double), ("disc_weight", ctypes.c_double),
        ("out_final", ctypes.c_void_p), ("done_flag", ctypes.c_void_p),
        ("done_value", ctypes.c_int32), ("flags", ctypes.c_int32),
        ("rows_min_aligned", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("plan", ctypes.POINTER(SvgtSegPlan)),
    ]


class SvgtError(RuntimeError):
    def __init__(self, code, message):
        self.code = code
        RuntimeError.__init__(self, "%s (%d): %s" % (ERR_NAMES.get(code, "SVGT_ERR"), code, message))


_lib = None


def lib():
    """Load libsvgt.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libsvgt.so is not built: run `python -c 'import __graft_entry__ as g; "
                              "g.build()'` (needs nvcc); there is no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        L.svgt_abi_version.restype = ctypes.c_int
        L.svgt_last_error.restype = ctypes.c_char_p
        L.svgt_device_count.restype = ctypes.c_int
        L.svgt_set_variant.restype = ctypes.c_int
        L.svgt_set_variant.argtypes = [ctypes.c_int]
        L.svgt_launches_per_batch.restype = ctypes.c_int
        L.svgt_launches_per_batch.argtypes = [ctypes.POINTER(SvgtBatch)]
        L.svgt_score_batch.restype = ctypes.c_int
        L.svgt_score_batch.argtypes = [ctypes.POINTER(SvgtBatch), ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p]
        L.svgt_ctx_create.restype = ctypes.c_int
        L.svgt_ctx_create.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        L.svgt_ctx_destroy.restype = ctypes.c_int
        L.svgt_ctx_destroy.argtypes = [ctypes.c_void_p]
        L.svgt_ctx_score_host.restype = ctypes.c_int
        L.svgt_ctx_score_host.argtypes = [ctypes.c_void_p, ctypes.POINTER(SvgtBatch), ctypes.c_void_p]
        L.svgt_ctx_last_traffic.restype = ctypes.c_int
        L.svgt_ctx_last_traffic.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64),
                                            ctypes.POINTER(ctypes.c_int64)]
        L.svgt_ctx_last_kernel_ms.restype = ctypes.c_int
        L.svgt_ctx_last_kernel_ms.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]
        L.svgt_score_compact.restype = ctypes.c_int
        L.svgt_score_compact.argtypes = [ctypes.POINTER(SvgtCBatch), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.svgt_ctx_score_host_compact.restype = ctypes.c_int
        L.svgt_ctx_score_host_compact.argtypes = [ctypes.c_void_p, ctypes.POINTER(SvgtCBatch), ctypes.c_void_p]
        L.svgt_shared_alloc.restype = ctypes.c_int
        L.svgt_shared_alloc.argtypes = [ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p]
        L.svgt_shared_open.restype = ctypes.c_int
        L.svgt_shared_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
        L.svgt_shared_close.restype = ctypes.c_int
        L.svgt_shared_close.argtypes = [ctypes.c_void_p]
        L.svgt_shared_free.restype = ctypes.c_int
        L.svgt_shared_free.argtypes = [ctypes.c_void_p]
        L.svgt_peer_copy.restype = ctypes.c_int
        L.svgt_peer_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        L.svgt_set_flag.restype = ctypes.c_int
        L.svgt_set_flag.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
        L.svgt_memcpy_d2h.restype = ctypes.c_int
        L.svgt_memcpy_d2h.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        L.svgt_wait_flags.restype = ctypes.c_int
        L.svgt_wait_flags.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
        L.svgt_plan_count.restype = ctypes.c_int
        L.svgt_plan_count.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                      ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)] + [ctypes.POINTER(ctypes.c_int64)] * 4
        L.svgt_plan_fill.restype = ctypes.c_int
        L.svgt_plan_fill.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.svgt_ctx_last_pieces.restype = ctypes.c_int
        L.svgt_ctx_last_pieces.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]
        if L.svgt_abi_version() != ABI_VERSION:
            raise ImportError("libsvgt.so ABI %d != expected %d" % (L.svgt_abi_version(), ABI_VERSION))
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise SvgtError(rc, lib().svgt_last_error().decode("utf-8", "replace"))


def set_variant(v):
    rc = lib().svgt_set_variant(int(v))
    if rc < 0:
        check(rc)
    return rc


def plan_pieces(sites, min_aligned=20, split_slop=3, resident_warps=0, force_chunks=0):
    """svgt_plan_count + svgt_plan_fill on host site rows ([n][12] int32): None when no site is longer than a piece,
    else dict(max_chunks, entries, pieces [n][4], heavy [n][4], scratch_chunks) -- numpy int32 arrays.
    `resident_warps` <= 0: those of the current CUDA device (148 x 20 without one)."""
    import numpy as np
    L = lib()
    sites = np.ascontiguousarray(sites, dtype=np.int32)
    n = int(sites.shape[0])
    mc = ctypes.c_int32()
    ne, npc, nh, sc = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    check(L.svgt_plan_count(sites.ctypes.data, n, int(min_aligned), int(split_slop), int(resident_warps),
                            int(force_chunks), ctypes.byref(mc), ctypes.byref(ne), ctypes.byref(npc), ctypes.byref(nh),
                            ctypes.byref(sc)))
    if nh.value == 0:
        return None
    entries = np.empty(ne.value, dtype=np.int32)
    pieces = np.empty((npc.value, 4), dtype=np.int32)
    heavy = np.empty((nh.value, 4), dtype=np.int32)
    check(L.svgt_plan_fill(sites.ctypes.data, n, int(min_aligned), int(split_slop), mc.value, entries.ctypes.data,
                           pieces.ctypes.data, heavy.ctypes.data))
    return {"max_chunks": mc.value, "entries": entries, "pieces": pieces, "heavy": heavy,
            "scratch_chunks": sc.value}
