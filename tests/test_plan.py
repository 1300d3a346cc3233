"""CPU: the piece planner of libsvgt.so (svgt_plan_count / svgt_plan_fill, include/svgt.h) -- host code, no GPU.

A plan cuts sites that are too long for one warp into pieces of whole 32-row chunks; the tally kernel scores the
pieces, svgt_replay_pieces_kernel sums them per site in row order (reference singlesample.py:364-378: the fp64 sums
run over the fragments one after another).  Here: the plan tiles every cut site exactly, names every other site with
rows exactly once, is sorted heaviest first, and its scratch ranges are disjoint.
"""
import numpy as np
import pytest

from svtyper_b200 import compact as cp, native, synth


def chunks(n):
    return (n + 31) // 32


def check_plan(sites, pl, L):
    n = sites.shape[0]
    nf, ns = sites[:, 10].astype(np.int64), sites[:, 11].astype(np.int64)
    skip = (sites[:, 7] & 16) != 0
    nf = np.where(skip, 0, nf)
    ns = np.where(skip, 0, ns)
    tot = chunks(nf) + chunks(ns)
    heavy_sites = np.nonzero(tot > L)[0]
    assert pl["max_chunks"] == L
    assert np.array_equal(pl["heavy"][:, 0], heavy_sites)
    ent = pl["entries"]
    whole = ent[ent >= 0]
    assert np.array_equal(np.sort(whole), np.nonzero((tot > 0) & (tot <= L))[0])          # each light site once, empty ones left out
    pieces = pl["pieces"]
    assert np.array_equal(np.sort(~ent[ent < 0]), np.arange(pieces.shape[0]))             # each piece once
    # heaviest first
    w = np.where(ent >= 0, tot[np.maximum(ent, 0)], chunks(pieces[np.maximum(~ent, 0), 2].astype(np.int64) & 0x7fffffff))
    assert (np.diff(w) <= 0).all() and w.max() <= L and w.min() >= 1
    # pieces tile their site: fragment chunks then split chunks, scratch chunks in row order
    at = 0
    k = 0
    for h in pl["heavy"]:
        site, q0, cf, cs = (int(x) for x in h)
        assert q0 == at and cf == chunks(nf[site]) and cs == chunks(ns[site])
        for part, rows, nch in ((0, nf[site], cf), (1, ns[site], cs)):
            r = 0
            while r < rows:
                p = pieces[k]
                cnt = int(p[2]) & 0x7fffffff
                assert int(p[0]) == site and int(p[1]) == r and r % 32 == 0
                assert (int(p[2]) < 0) == bool(part)
                assert cnt == min(rows - r, 32 * L)
                assert int(p[3]) == at + (cf if part else 0) + r // 32
                r += cnt
                k += 1
        at += cf + cs
    assert k == pieces.shape[0] and at == pl["scratch_chunks"]


@pytest.mark.parametrize("config,n", [("stress1m", 3000), ("del10k", 2000), ("mixed100k", 2000)])
@pytest.mark.parametrize("L", [1, 3, 8, 33])
def test_forced_piece_length(config, n, L):
    cb = cp.compact_from_wide(synth.generate(config, n_sites=n, seed=7))
    pl = native.plan_pieces(cb.sites, force_chunks=L)
    tot = chunks(np.where(cb.sites[:, 7] & 16, 0, cb.sites[:, 10])) + chunks(np.where(cb.sites[:, 7] & 16, 0, cb.sites[:, 11]))
    if tot.max() <= L:
        assert pl is None
    else:
        check_plan(cb.sites, pl, L)


def test_policy_scales_with_the_batch():
    """No plan for an evenly sized large batch (the benchmark shape keeps its kernel); pieces for a heavy-tailed or
    small one; the piece length grows with the batch's chunks per resident warp and never drops below 4."""
    big = cp.compact_from_wide(synth.generate("del1m4lib", n_sites=60_000, seed=1))
    tot = int((chunks(big.sites[:, 10]) + chunks(big.sites[:, 11])).sum())
    assert native.plan_pieces(big.sites, resident_warps=8) is None          # few warps: every site is short by comparison
    pl = native.plan_pieces(big.sites, resident_warps=2960)
    assert pl is not None and pl["max_chunks"] == max(4, -(-tot // (14 * 2960)))
    check_plan(big.sites, pl, pl["max_chunks"])
    small = cp.compact_from_wide(synth.generate("del10k", n_sites=300, seed=2))
    pl = native.plan_pieces(small.sites, resident_warps=2960)
    assert pl is not None and pl["max_chunks"] == 4
    check_plan(small.sites, pl, 4)


def test_refused_sites_get_no_pieces():
    """SKIP sites, sites outside the coordinate range and malformed counts have no rows for the kernel: the plan
    leaves them out (they are still called -- blank or flagged -- by the call kernel)."""
    cb = cp.compact_from_wide(synth.generate("del10k", n_sites=200, seed=4))
    s = cb.sites.copy()
    s[3, 7] |= 16
    s[5, 0] = (1 << 30) + 1
    s[9, 10] = -4
    pl = native.plan_pieces(s, force_chunks=1)
    for bad in (3, 5, 9):
        assert bad not in pl["heavy"][:, 0] and bad not in pl["entries"]
    assert native.plan_pieces(s[:0], force_chunks=1) is None
