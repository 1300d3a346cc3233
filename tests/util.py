"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

from svtyper_b200 import evidence as ev

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INT_FIELDS = ("GT", "GQ", "DP", "RO", "AO", "QR", "QA", "RS", "AS", "ASC", "RP", "AP")
GL_TOL = 1e-6          # BASELINE.json north_star: GL within 1e-6; integer fields bit-exact


def fixture_library_table():
    with open(os.path.join(REPO, "tests", "data", "NA12878.bam.json")) as f:
        lib = json.load(f)["NA12878"]["libraryArray"][0]
    hist = {int(k): int(v) for k, v in lib["histogram"].items()}
    return ev.LibraryTable([(float(lib["mean"]), float(lib["sd"]), hist)])


def batch_from_npz(z):
    libs = fixture_library_table()
    assert np.array_equal(libs.hist, z["hist"]) and np.array_equal(libs.lib_f64, z["lib_f64"])
    return ev.EvidenceBatch(z["sites"], z["frags"], z["splits"], libs)


def assert_rows_match(got, exp, exact_gl=False, where=""):
    """Integer FORMAT fields bit-exact; GL within GL_TOL (or bit-exact); SQ within 1e-9 relative."""
    assert got.shape == exp.shape, (got.shape, exp.shape)
    for k in INT_FIELDS:
        bad = np.nonzero(got[k] != exp[k])[0]
        assert bad.size == 0, "%s field %s differs at sites %s: got %s want %s" % (
            where, k, bad[:8], got[k][bad[:8]], exp[k][bad[:8]])
    if exact_gl:
        assert np.array_equal(got["GL"], exp["GL"]), where + " GL not bit-exact"
    else:
        assert np.allclose(got["GL"], exp["GL"], rtol=0, atol=GL_TOL), where + " GL beyond 1e-6"
    called = exp["GT"] >= 0
    assert np.allclose(got["SQ"][called], exp["SQ"][called], rtol=1e-9, atol=1e-9), where + " SQ"
