"""CPU: the compact (16-byte row) schema carries exactly the evidence of the wide one.

The oracle reads wide rows; the kernels read compact rows.  These tests pin the encoding:
wide -> compact -> wide must score identically under the oracle on every config shape, on the
reference's own fixture, and on the escape cases (EXTRA rows with both slots, reads longer than
the 14-bit length field, split pieces longer than 16 bits at every chunk position).
"""
import numpy as np
import pytest

from svtyper_b200 import compact as cp, evidence as ev, synth
from util import assert_rows_match


def roundtrip(b):
    cb = cp.compact_from_wide(b)
    return cb, cp.wide_from_compact(cb)


@pytest.mark.parametrize("config,n", [("del10k", 3000), ("mixed100k", 4000), ("del1m4lib", 3000), ("stress1m", 1500)])
def test_roundtrip_scores_identically(oracle, config, n):
    b = synth.generate(config, n_sites=n)
    cb, w = roundtrip(b)
    assert cb.rows.shape[1] == 4 and cb.sites.shape[1] == 12
    for assoc in (ev.ASSOC_SSO, ev.ASSOC_CLASSIC):
        exp = oracle.score(b, assoc_mode=assoc)
        got = oracle.score(w, assoc_mode=assoc)
        assert got.tobytes() == exp.tobytes(), config
    # rows: every wide row is one compact row, except EXTRA rows (their verdict rides on the MULTI row)
    skip = (b.sites[:, 9] & ev.SITE_SKIP) != 0
    assert cb.n_split == int(b.sites[~skip, 15].sum())
    assert cb.n_frag == int(b.sites[~skip, 12].sum()) - int(((b.frags[:, 7] & ev.F_EXTRA) != 0).sum())
    assert cb.algorithmic_bytes() == cb.n_sites * 128 + 16 * cb.n_rows
    # second conversion of the decoded batch reproduces the compact rows bit for bit
    cb2 = cp.compact_from_wide(w)
    assert np.array_equal(cb2.rows, cb.rows) and np.array_equal(cb2.sites, cb.sites)


def test_fixture_roundtrip(oracle, fixture_batch, fixture_npz):
    cb, w = roundtrip(fixture_batch)
    got = oracle.score(w)
    exp = fixture_npz["expected_sso"]
    for k in ("GT", "GQ", "DP", "RO", "AO", "QR", "QA", "RS", "AS", "ASC", "RP", "AP"):
        assert np.array_equal(got[k], exp[k]), k
    assert got.tobytes() == oracle.score(fixture_batch).tobytes()


def test_slice_and_order(oracle):
    b = synth.generate("mixed100k", n_sites=2000, seed=4)
    cb = cp.compact_from_wide(b)
    assert cb.order is not None and sorted(cb.order.tolist()) == list(range(2000))
    part = cb.slice_sites(500, 900)
    exp = oracle.score(b.slice_sites(500, 900))
    assert oracle.score(cp.wide_from_compact(part)).tobytes() == exp.tobytes()
    assert cb.slice_sites(7, 7).n_sites == 0


def _one_site(frags, splits, svtype=ev.SV_DEL, posA=100_000, length=3000, tids=(3, 3)):
    libs = synth.make_libraries(1)
    o1, o2 = 0, 1
    site = [posA + o1, posA + length + o2, 0, 0, 0, 0, tids[0], tids[1], length if svtype == ev.SV_DEL else 0,
            svtype | (o1 << 2) | (o2 << 3), 0, 0, len(frags), 0, 0, len(splits)]
    return ev.EvidenceBatch(np.array([site], np.int32), np.array(frags, np.int64).astype(np.int32).reshape(-1, 8),
                            np.array(splits, np.int64).astype(np.int32).reshape(-1, 8), libs)


def test_escapes_long_reads_and_two_slot_extra(oracle):
    pA, L = 100_000, 3000
    q = 60 | (60 << 8)
    P = ev.F_HAS_A | ev.F_HAS_B | ev.F_PAIRED | ev.F_REV_B
    frags = [
        # long single-block read A (20 kb span) covering breakend A: must hit through an escape row
        (pA - 50, pA - 50 + 20_000, pA + 200, pA + 301, 3, 3, q, P),
        # EXTRA row holding an interval for both slots, then the MULTI_A | MULTI_B main row
        (pA - 40, pA + 30, pA + L - 30, pA + L + 40, 3, 3, 0, ev.F_EXTRA | ev.F_HAS_A | ev.F_HAS_B),
        (pA - 40, pA + 500, pA + L - 30, pA + L + 400, 3, 3, q, P | ev.F_MULTI_A | ev.F_MULTI_B),
        # long read B
        (pA - 300, pA - 199, pA - 20, pA + 70_000, 3, 3, q, P),
        # ordinary pair + lone read
        (pA - 250, pA - 149, pA + 60, pA + 161, 3, 3, q, P),
        (pA - 30, pA + 71, 0, 0, 3, 0, 60, ev.F_HAS_A),
    ]
    b = _one_site(frags, [])
    cb, w = roundtrip(b)
    assert cb.n_frag == 5              # the EXTRA row and the over-long spans travel as hit bits
    w3 = cb.rows[:, 3].astype(np.int64) & 0xFFFFFFFF
    assert int(((w3 & cp.CF_MULTI_A) != 0).sum()) == 2 and int(((w3 & cp.CF_MULTI_B) != 0).sum()) == 2
    exp = oracle.score(b)
    assert exp["RS"][0] >= 4
    assert oracle.score(w).tobytes() == exp.tobytes()


@pytest.mark.parametrize("n_before", [0, 5, 30, 31, 62, 63])
def test_wide_split_pieces(oracle, n_before):
    pA, L = 100_000, 3000
    m = 60 | (60 << 8)
    ok = (3, pA - 80, pA, 3, pA + L + 1, pA + L + 90, m | (ev.S_FIRST << 16), 0)
    wide = (3, pA - 70_000, pA, 3, pA + L + 1, pA + L + 100_000, m | (ev.S_FIRST << 16), 0)
    splits = [ok] * n_before + [wide, ok, wide]
    b = _one_site([], splits)
    cb, w = roundtrip(b)
    sp = cb.rows[int(cb.sites[0, 10]):]
    pos = np.nonzero(sp[:, 3].astype(np.int64) & cp.CSP_WIDE)[0]
    assert pos.size == 2 and not (pos % 32 == 31).any()
    assert ((sp[pos + 1, 3] & cp.CSP_XEND) != 0).all()
    exp = oracle.score(b)
    assert exp["AS"][0] == n_before + 2     # (n + 3) * (1 - 1e-6), truncated
    assert oracle.score(w).tobytes() == exp.tobytes()


def test_min_aligned_is_part_of_the_encoding(oracle):
    b = synth.generate("mixed100k", n_sites=1500, seed=12)
    for m in (5, 20, 35):
        cb = cp.compact_from_wide(b, min_aligned=m)
        assert cb.min_aligned == m
        got = oracle.score(cp.wide_from_compact(cb), min_aligned=m)
        assert got.tobytes() == oracle.score(b, min_aligned=m).tobytes(), m


def test_suggested_unit_mode_follows_the_tail():
    heavy = cp.compact_from_wide(synth.generate("stress1m", n_sites=20000, seed=2))
    even = cp.compact_from_wide(synth.generate("del1m4lib", n_sites=20000, seed=2))
    assert heavy.suggest_unit_mode() == 3                      # a 5000-row site dwarfs a warp's share of 20k sites
    assert even.suggest_unit_mode(resident_warps=8) == 1       # few warps: every warp has plenty of even units
    assert cp.CompactBatch(even.sites, even.rows, even.libs).suggest_unit_mode() == 1     # no launch order


def test_library_index_limit():
    b = synth.generate("del10k", n_sites=50, seed=1)
    bad = ev.EvidenceBatch(b.sites.copy(), b.frags.copy(), b.splits.copy(), b.libs)
    bad.frags[:, 6] |= 600 << 16
    with pytest.raises(ValueError):
        cp.compact_from_wide(bad)
