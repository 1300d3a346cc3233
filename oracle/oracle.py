"""ctypes front-end for the CPU parity oracle (oracle/svgt_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
CPU-baseline legs.  The product package (svtyper_b200) never imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsvgt_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "svgt_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = ctypes.CDLL(LIB_PATH)
        L.svgt_oracle_log_choose.restype = ctypes.c_double
        L.svgt_oracle_log_choose.argtypes = [ctypes.c_int64, ctypes.c_int64]
        L.svgt_oracle_bayes_gt.restype = None
        L.svgt_oracle_bayes_gt.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                           ctypes.POINTER(ctypes.c_double)]
        L.svgt_oracle_prob_mapq.restype = ctypes.c_double
        L.svgt_oracle_prob_mapq.argtypes = [ctypes.c_int]
        L.svgt_oracle_max_threads.restype = ctypes.c_int
        L.svgt_oracle_score.restype = ctypes.c_int
        L.svgt_oracle_score.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_int]
        _lib = L
    return _lib


def log_choose(n, k):
    return lib().svgt_oracle_log_choose(int(n), int(k))


def bayes_gt(ref, alt, is_dup):
    out = (ctypes.c_double * 3)()
    lib().svgt_oracle_bayes_gt(int(ref), int(alt), int(bool(is_dup)), out)
    return tuple(out)


def prob_mapq(q):
    return lib().svgt_oracle_prob_mapq(int(q))


def max_threads():
    return lib().svgt_oracle_max_threads()


def score(batch, min_aligned=20, split_slop=3, split_weight=1.0, disc_weight=1.0,
          assoc_mode=0, n_threads=1):
    """Score an svtyper_b200.evidence.EvidenceBatch on the CPU; returns OUT_DTYPE rows."""
    from svtyper_b200.evidence import OUT_DTYPE
    out = np.zeros(batch.n_sites, dtype=OUT_DTYPE)
    lt = batch.libs
    arrs = [np.ascontiguousarray(a) for a in
            (batch.sites, batch.frags, batch.splits, lt.lib_f64, lt.lib_i32, lt.hist)]
    rc = lib().svgt_oracle_score(
        arrs[0].ctypes.data, batch.n_sites, arrs[1].ctypes.data, arrs[2].ctypes.data,
        arrs[3].ctypes.data, arrs[4].ctypes.data, lt.n_lib, arrs[5].ctypes.data,
        int(min_aligned), int(split_slop), float(split_weight), float(disc_weight),
        int(assoc_mode), out.ctypes.data, int(n_threads))
    if rc != 0:
        raise RuntimeError("svgt_oracle_score failed: %d" % rc)
    return out
