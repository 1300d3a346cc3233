"""CPU: the C oracle restatement against the reference's own golden vectors.

tests/golden/fixture_evidence.npz was produced by tests/golden/make_golden.py from the
REFERENCE ITSELF on the reference's own fixture (tests/data, reference
tests/test_singlesample.py:20-44): per-breakpoint tally counts, bayes_gt likelihoods and
FORMAT integers for 211 breakpoints, plus the classic entry point's debug dump.
"""
import math

import numpy as np
import pytest

from svtyper_b200 import evidence as ev
from util import INT_FIELDS, assert_rows_match


def test_fixture_shape(fixture_npz):
    z = fixture_npz
    assert z["sites"].shape == (211, ev.SITE_WORDS)
    assert z["expected_sso"].shape == (211,)
    sv = list(z["svtypes"])
    assert (sv.count("DEL"), sv.count("DUP"), sv.count("INV"), sv.count("BND")) == (204, 1, 5, 1)


def test_oracle_matches_reference_sso(oracle, fixture_batch, fixture_npz):
    got = oracle.score(fixture_batch, assoc_mode=ev.ASSOC_SSO)
    exp = fixture_npz["expected_sso"]
    for k in INT_FIELDS:
        assert np.array_equal(got[k], exp[k]), k
    assert np.array_equal(got["GL"], exp["GL"])          # same libm, same order: bit-exact
    called = exp["GT"] >= 0
    assert np.allclose(got["SQ"][called], exp["SQ"][called], rtol=1e-12)
    # the golden VCF's genotype mix (84 0/0, 93 0/1, 35 1/1 over 212 records,
    # one BND pair sharing a breakpoint)
    assert np.bincount(exp["GT"], minlength=3).tolist() == [84, 92, 35]


def test_oracle_matches_reference_classic(oracle, fixture_batch, fixture_npz):
    got = oracle.score(fixture_batch, assoc_mode=ev.ASSOC_CLASSIC)
    cl = fixture_npz["expected_classic"]
    has = cl["has_gl"] == 1
    assert np.allclose(got["GL"][has], cl["GL"][has], rtol=0, atol=1e-9)   # debug dump prints repr
    sso = fixture_npz["expected_sso"]
    for k in INT_FIELDS:                     # both entry points give one golden VCF
        assert np.array_equal(got[k], sso[k]), k


# Known-answer tests computed from reference statistics.py
KATS = [
    ((0, 59, False), (-176.99999999999997, -17.760769744174887, -2.6996919430798316)),
    ((45, 33, False), (-76.982039933686, -1.442826565645733, -24.472484092357483)),
    ((126, 0, False), (-0.054748483526229144, -37.929779453661624, -126.0)),
    ((0, 0, False), (0.0, 0.0, 0.0)),
    ((10, 10, True), (-14.777049502900372, -2.692201622316625, -1.26552658662931)),
    ((5000, 1500, False), (-2979.1528406405264, -433.67525358631656, -3545.616517611452)),
]


@pytest.mark.parametrize("args,want", KATS)
def test_bayes_gt_kats(oracle, args, want):
    assert oracle.bayes_gt(*args) == want


def test_log_choose_kats(oracle):
    assert oracle.log_choose(78, 33) == 22.037513096144796
    assert oracle.log_choose(4000, 2000) == 1202.2208655783427


def test_prob_mapq_and_luts(oracle):
    pm, logt, consts = ev.build_luts(4096)
    for q in range(256):
        assert pm[q] == oracle.prob_mapq(q) == 1 - 10 ** (-q / 10.0)
    assert pm[0] == 0.0 and pm[255] == 1.0
    assert logt[1000] == math.log(1000, 10) == 2.9999999999999996
    assert consts[ev.C_NONDUP_ALT] == -2.9999999999999996
    assert 10.0 ** consts[ev.C_POW10_MIN_X] > 0.0
    assert 10.0 ** math.nextafter(consts[ev.C_POW10_MIN_X], -math.inf) == 0.0


def test_order_sensitivity_hazards(oracle):
    """sequential fp64 sums truncate differently from tree sums."""
    from svtyper_b200 import synth
    hz = synth.hazard_batch()
    out = oracle.score(hz)
    pm = ev.build_luts(16)[0]
    # 30 fragments of MAPQ 10 whose first read covers the breakpoint -> RS = int(seq. sum) = 26
    s = 0.0
    for _ in range(30):
        s += pm[10]
    assert out["RS"][0] == int(s) == 26 and int(30 * pm[10]) == 27
    acc = 0.0
    for _ in range(30):
        acc += pm[10] * pm[10] / 2       # pair straddles A only: (1 + 0) * p / 2
    assert out["RP"][0] == int(acc) == 12
    assert out["GT"][7] == ev.GT_BLANK                 # two MAPQ-0 fragments: all sums 0.0
