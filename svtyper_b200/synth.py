"""Synthetic VCF + read-evidence generators for the BASELINE.json configs.

The generators emit batches directly in the device layout of `evidence.py` (there
is no BAM behind them), with fragments already in their final (sorted-qname) order.
All randomness is a seeded `numpy.random.Philox` stream, so every rank / test can
regenerate the same shard.

  config 2  "del10k"    10k DEL sites, one library (the fixture's real insert histogram),
                        <=1000 reads/site
  config 3  "mixed100k" 100k sites, 70% DEL / 10% DUP / 10% INV / 10% BND
  config 4  "del1m4lib" 1M DEL sites, 4 libraries with distinct histograms
  config 5  "stress1m"  1M sites, reads/site up to max_reads=10000 incl. empty and
                        over-threshold (SKIP) sites

`n_sites` can be overridden (tests use a few hundred sites of each shape).
"""
from __future__ import annotations

import json
import math
import os

import numpy as np

from . import evidence as ev

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "data")

CONFIGS = {
    "del10k": dict(n_sites=10_000, mix=(1.0, 0.0, 0.0, 0.0), n_lib=1, reads_mu=math.log(220),
                   reads_sigma=0.5, reads_min=4, max_reads=1000, seed_off=2),
    "mixed100k": dict(n_sites=100_000, mix=(0.7, 0.1, 0.1, 0.1), n_lib=1, reads_mu=math.log(220),
                      reads_sigma=0.5, reads_min=4, max_reads=1000, seed_off=3),
    "del1m4lib": dict(n_sites=1_000_000, mix=(1.0, 0.0, 0.0, 0.0), n_lib=4, reads_mu=math.log(220),
                      reads_sigma=0.5, reads_min=4, max_reads=1000, seed_off=4),
    "stress1m": dict(n_sites=1_000_000, mix=(0.7, 0.1, 0.1, 0.1), n_lib=4, reads_mu=math.log(200),
                     reads_sigma=1.5, reads_min=0, max_reads=10_000, seed_off=5),
}
BASE_SEED = 20261017
READ_LEN = 101


def fixture_library():
    """(mean, sd, hist) of the reference fixture library (tests/data/NA12878.bam.json)."""
    with open(os.path.join(_DATA, "NA12878.bam.json")) as f:
        lib = json.load(f)["NA12878"]["libraryArray"][0]
    return float(lib["mean"]), float(lib["sd"]), {int(k): int(v) for k, v in lib["histogram"].items()}


def gaussian_library(mean, sd, n=1_000_000):
    """Discretised Gaussian insert-size histogram with the stated mean/sd recorded verbatim."""
    lo, hi = max(1, int(mean - 4 * sd)), int(mean + 4 * sd)
    xs = np.arange(lo, hi + 1)
    w = np.exp(-0.5 * ((xs - mean) / sd) ** 2)
    c = np.floor(w / w.sum() * n).astype(np.int64)
    return float(mean), float(sd), {int(x): int(v) for x, v in zip(xs, c) if v > 0}


def make_libraries(n_lib):
    libs = [fixture_library()]
    # (550, 150) gives mean + 3*sd == 1000.0: an integral float key
    for mean, sd in ((250, 40), (400, 100), (550.0, 150.0)):
        if len(libs) < n_lib:
            libs.append(gaussian_library(mean, sd))
    return ev.LibraryTable(libs[:n_lib])


def _mapq_sampler(rng, n):
    """Fixture-like MAPQ mix: 94% 60, 1.7% 40, 0.3% 0, rest spread over 1..59."""
    u = rng.random(n)
    q = np.full(n, 60, dtype=np.int64)
    q[u < 0.06] = rng.integers(1, 60, size=int((u < 0.06).sum()))
    q[u < 0.02] = 40
    q[u < 0.003] = 0
    return q


def generate(config="del10k", n_sites=None, seed=None, rank=0, bucket=True, libs=None) -> ev.EvidenceBatch:
    cfg = CONFIGS[config]
    n = int(cfg["n_sites"] if n_sites is None else n_sites)
    seed = (BASE_SEED + cfg["seed_off"]) if seed is None else seed
    rng = np.random.Generator(np.random.Philox(key=[seed, rank]))
    libs = make_libraries(cfg["n_lib"]) if libs is None else libs
    n_lib = libs.n_lib
    flank = libs.lib_f64[:, 0]

    # ---------------- sites ----------------
    svtype = rng.choice(4, size=n, p=np.array(cfg["mix"]) / sum(cfg["mix"])).astype(np.int64)
    tidA = rng.integers(0, 24, size=n)
    tidB = tidA.copy()
    inter = (svtype == ev.SV_BND) & (rng.random(n) < 0.5)
    tidB[inter] = (tidA[inter] + 1 + rng.integers(0, 23, size=int(inter.sum()))) % 24
    posA = rng.integers(20_000, 150_000_000, size=n)
    length = np.exp(rng.uniform(math.log(50), math.log(50_000), size=n)).astype(np.int64)
    posB = posA + length
    posB[inter] = rng.integers(20_000, 150_000_000, size=int(inter.sum()))
    o1 = np.zeros(n, dtype=np.int64)
    o2 = np.zeros(n, dtype=np.int64)
    o2[svtype == ev.SV_DEL] = 1
    o1[svtype == ev.SV_DUP] = 1
    bnd = svtype == ev.SV_BND
    o1[bnd] = rng.integers(0, 2, size=int(bnd.sum()))
    o2[bnd] = rng.integers(0, 2, size=int(bnd.sum()))
    var_length = np.where(svtype == ev.SV_DEL, posB - posA, 0)
    posA = posA + o1
    posB = posB + o2
    ci = np.zeros((n, 4), dtype=np.int64)
    wide = rng.random(n) >= 0.7
    lohi = np.sort(rng.integers(-10, 11, size=(n, 2, 2)), axis=2)
    ci[wide, 0:2] = lohi[wide, 0]
    ci[wide, 2:4] = lohi[wide, 1]

    # reads per site -> fragments per site
    reads = np.rint(np.exp(rng.normal(cfg["reads_mu"], cfg["reads_sigma"], size=n))).astype(np.int64)
    reads = np.clip(reads, cfg["reads_min"], None)
    skip = reads > cfg["max_reads"]            # too-many-reads rows (singlesample.py:172)
    # hazard stratum: homogeneous low-MAPQ sites at counts that make the
    # fp64 sums land exactly on / next to integers
    hazard = (rng.random(n) < 0.02) & ~skip
    hz_count = rng.choice([30, 100, 700, 1000], size=n)
    hz_mapq = rng.choice([10, 20, 30], size=n)
    nfrag = (reads + 1) // 2
    nfrag[hazard] = np.minimum(hz_count[hazard], max(cfg["max_reads"], 1000))
    nfrag[skip] = 0
    N = int(nfrag.sum())
    site_of = np.repeat(np.arange(n), nfrag)

    # ---------------- fragments ----------------
    lib_p = np.array(([0.4, 0.3, 0.2, 0.1] + [0.05] * n_lib)[:n_lib])
    lib = rng.choice(n_lib, size=N, p=lib_p / lib_p.sum())
    # insert size ~ library histogram (inverse CDF)
    ins = np.zeros(N, dtype=np.int64)
    for l in range(n_lib):
        off, hlen = libs.lib_i32[l, 0], libs.lib_i32[l, 1]
        cdf = np.cumsum(libs.hist[off:off + hlen].astype(np.float64))
        sel = np.nonzero(lib == l)[0]
        ins[sel] = np.searchsorted(cdf, rng.random(sel.size) * cdf[-1], side="right")
    ins = np.maximum(ins, 2 * READ_LEN // 2)
    m = 20
    cat = rng.choice(6, size=N, p=[0.28, 0.27, 0.15, 0.12, 0.10, 0.08])
    # 0 ref-span at A, 1 ref-span at B, 2 alt-span, 3 near-miss geometry, 4 far away, 5 single primary
    # per-site true genotype shapes the evidence mix: hom-ref sites lose their alt pairs,
    # hom-alt sites lose most reference-spanning pairs
    truth = rng.choice(3, size=n, p=[0.4, 0.35, 0.25])
    tr = truth[site_of]
    r_ = rng.random(N)
    cat = np.where((tr == 0) & ((cat == 2) | (cat == 3)) & (r_ < 0.97), 0, cat)
    cat = np.where((tr == 2) & ((cat == 0) | (cat == 1)) & (r_ < 0.92), 2, cat)
    sA, sB = posA[site_of], posB[site_of]
    so1, so2 = o1[site_of], o2[site_of]
    cA0, cA1, cB0, cB1 = (ci[site_of, k] for k in range(4))
    fl = flank[lib]
    u = (rng.random(N) * (ins - 2 * m - 2)).astype(np.int64) + m + 1     # distance of a_start left of the point

    a_start = np.zeros(N, dtype=np.int64)
    b_end = np.zeros(N, dtype=np.int64)
    tA = tidA[site_of].copy()
    tB = tidA[site_of].copy()
    revA = np.zeros(N, dtype=np.int64)
    revB = np.ones(N, dtype=np.int64)

    k0 = cat == 0
    a_start[k0] = sA[k0] - u[k0]
    b_end[k0] = a_start[k0] + ins[k0]
    k1 = cat == 1
    a_start[k1] = sB[k1] - u[k1]
    b_end[k1] = a_start[k1] + ins[k1]
    tA[k1] = tidB[site_of][k1]
    tB[k1] = tidB[site_of][k1]
    # alt-span: each mate placed on its own side of the junction, oriented like the breakend
    k2 = (cat == 2) | (cat == 3)
    du = (rng.random(N) * np.minimum(fl, ins)).astype(np.int64)
    dv = np.maximum(ins - du, 0)
    near = cat == 3                      # push one side just past its window edge
    du = np.where(near & (rng.random(N) < 0.5), (fl + rng.integers(0, 3, size=N)).astype(np.int64), du)
    dv = np.where(near & (du <= fl), -rng.integers(1, 3, size=N), dv)
    i0 = np.where(so1 == 0, sA + cA1 - du, sA + cA0 + du)
    i1 = np.where(so2 == 0, sB + cB1 - dv, sB + cB0 + dv)
    a_start[k2] = (i0 - m)[k2]
    b_end[k2] = (i1 + m + 1)[k2]
    revA[k2] = so1[k2]
    revB[k2] = so2[k2]
    tB[k2] = tidB[site_of][k2]
    inv_flip = k2 & (svtype[site_of] == ev.SV_INV) & (rng.random(N) < 0.5)   # reciprocal INV pairs
    revA[inv_flip] = 1 - revA[inv_flip]
    revB[inv_flip] = 1 - revB[inv_flip]
    k4 = cat == 4
    a_start[k4] = sA[k4] + rng.integers(-3000, 3000, size=int(k4.sum()))
    b_end[k4] = a_start[k4] + ins[k4]
    revA[k4] = rng.integers(0, 2, size=int(k4.sum()))
    k5 = cat == 5
    a_start[k5] = np.where(rng.random(int(k5.sum())) < 0.5, sA[k5], sB[k5]) - rng.integers(0, 140, size=int(k5.sum()))
    b_end[k5] = a_start[k5] + READ_LEN
    tA[k5] = np.where(rng.random(int(k5.sum())) < 0.9, tA[k5], (tA[k5] + 1) % 24)
    a_start = np.maximum(a_start, 0)
    b_end = np.maximum(b_end, a_start + 1)
    rl_a = READ_LEN - (rng.random(N) < 0.1) * rng.integers(1, 60, size=N)     # some clipped reads
    rl_b = READ_LEN - (rng.random(N) < 0.1) * rng.integers(1, 60, size=N)
    a_end = a_start + rl_a
    b_start = np.maximum(b_end - rl_b, 0)

    mqA = _mapq_sampler(rng, N)
    mqB = _mapq_sampler(rng, N)
    hz = hazard[site_of]
    mqA[hz] = hz_mapq[site_of][hz]
    mqB[hz] = hz_mapq[site_of][hz]

    flags = np.full(N, ev.F_HAS_A | ev.F_HAS_B | ev.F_PAIRED, dtype=np.int64)
    flags |= revA * ev.F_REV_A | revB * ev.F_REV_B
    flags[k5] = ev.F_HAS_A | (revA[k5] * ev.F_REV_A)
    # rare structural rows: continuation rows (num_primary > 2) and gapped reads (EXTRA + MULTI)
    first_of_site = np.zeros(N, dtype=bool)
    first_of_site[np.cumsum(nfrag)[:-1][nfrag[1:] > 0]] = True if N else False
    if N:
        first_of_site[0] = True
    prev_paired = np.concatenate(([False], cat[:-1] != 5))   # CONT only extends a 2-primary fragment
    cont = k5 & (rng.random(N) < 0.05) & ~first_of_site & prev_paired
    flags[cont] |= ev.F_CONT
    # a fragment with a third primary has num_primary != 2, so its first two reads no longer
    # form a readA/readB pair (parsers.py:827): drop PAIRED on the row being continued
    flags[np.nonzero(cont)[0] - 1] &= ~ev.F_PAIRED
    frags = np.zeros((N, ev.FRAG_WORDS), dtype=np.int64)
    frags[:, 0], frags[:, 1], frags[:, 2], frags[:, 3] = a_start, a_end, b_start, b_end
    frags[:, 4], frags[:, 5] = tA, tB
    frags[k5, 2] = frags[k5, 3] = frags[k5, 5] = 0
    frags[:, 6] = mqA | (np.where(k5, 0, mqB) << 8) | (lib << 16)
    # gapped reads: turn a main row into (EXTRA row carrying the two pieces' first interval,
    # main row flagged MULTI_A) by rewriting the row *before* it when that row is far-away filler
    gap = np.zeros(N, dtype=bool)
    if N > 2:
        cand = np.nonzero((cat[1:] != 5) & (cat[:-1] == 4) & (site_of[1:] == site_of[:-1])
                          & (rng.random(N - 1) < 0.03))[0] + 1
        cand = cand[np.concatenate(([True], np.diff(cand) > 1))] if cand.size else cand
        gap[cand] = True
        e = cand - 1
        cut = frags[cand, 0] + rl_a[cand] // 2
        frags[e, :] = 0
        frags[e, 0], frags[e, 1], frags[e, 4] = frags[cand, 0], cut, frags[cand, 4]
        frags[e, 6] = lib[cand] << 16
        flags[e] = ev.F_EXTRA | ev.F_HAS_A
        flags[cand] |= ev.F_MULTI_A
    frags[:, 7] = flags

    # ---------------- splits ----------------
    nsplit = rng.binomial(nfrag, 0.16)
    S = int(nsplit.sum())
    ssite = np.repeat(np.arange(n), nsplit)
    soft = rng.random(S) < 0.55
    side_b = rng.random(S) < 0.5                      # which breakend the primary piece sits on
    p_pos = np.where(side_b, posB[ssite], posA[ssite])
    p_tid = np.where(side_b, tidB[ssite], tidA[ssite])
    q_pos = np.where(side_b, posA[ssite], posB[ssite])
    q_tid = np.where(side_b, tidA[ssite], tidB[ssite])
    p_sup = np.where(truth[ssite] == 0, 0.03, 0.6)     # hom-ref sites: splits rarely line up
    jit = np.where(rng.random(S) < p_sup, rng.integers(-4, 5, size=S), rng.integers(-60, 61, size=S))
    jit2 = np.where(rng.random(S) < p_sup, rng.integers(-4, 5, size=S), rng.integers(-60, 61, size=S))
    plen = rng.integers(25, 80, size=S)
    qlen = rng.integers(25, 80, size=S)
    # a piece "supports" a breakend through its start (reverse side) or end (forward side)
    p_start = np.where(rng.random(S) < 0.5, p_pos + jit, p_pos + jit - plen)
    q_start = np.where(rng.random(S) < 0.5, q_pos + jit2, q_pos + jit2 - qlen)
    p_start = np.maximum(p_start, 0)
    q_start = np.maximum(q_start, 0)
    pm_ = _mapq_sampler(rng, S)
    qm_ = _mapq_sampler(rng, S)
    P = np.stack([p_tid, p_start, p_start + plen, pm_], axis=1)
    Q = np.stack([q_tid, q_start, q_start + qlen, qm_], axis=1)
    Q[soft] = (ev.TID_NONE, 1, 1, 0)                  # SplitPiece(None, 1, ..., mapq 0), parsers.py:976-981
    swap = rng.random(S) < 0.5
    Lp = np.where(swap[:, None], Q, P)
    Rp = np.where(swap[:, None], P, Q)
    sfirst = rng.random(S) < 0.85
    if S:
        starts = np.cumsum(nsplit) - nsplit
        sfirst[starts[nsplit > 0]] = True
    sflags = soft * ev.S_SOFT_CLIP | sfirst * ev.S_FIRST
    splits = np.zeros((S, ev.SPLIT_WORDS), dtype=np.int64)
    splits[:, 0:3] = Lp[:, 0:3]
    splits[:, 3:6] = Rp[:, 0:3]
    splits[:, 6] = Lp[:, 3] | (Rp[:, 3] << 8) | (sflags << 16)
    splits[:, 7] = np.cumsum(sfirst) - 1 if S else 0

    # ---------------- site rows ----------------
    sites = np.zeros((n, ev.SITE_WORDS), dtype=np.int64)
    sites[:, 0], sites[:, 1] = posA, posB
    sites[:, 2:6] = ci
    sites[:, 6], sites[:, 7] = tidA, tidB
    sites[:, 8] = var_length
    sites[:, 9] = svtype | (o1 << 2) | (o2 << 3) | (skip.astype(np.int64) << 4)
    foff = np.cumsum(nfrag) - nfrag
    soff = np.cumsum(nsplit) - nsplit
    sites[:, 10], sites[:, 11], sites[:, 12] = foff & 0xFFFFFFFF, foff >> 32, nfrag
    sites[:, 13], sites[:, 14], sites[:, 15] = soff & 0xFFFFFFFF, soff >> 32, nsplit
    batch = ev.EvidenceBatch((sites & 0xFFFFFFFF).astype(np.uint32).view(np.int32),
                             (frags & 0xFFFFFFFF).astype(np.uint32).view(np.int32),
                             (splits & 0xFFFFFFFF).astype(np.uint32).view(np.int32), libs)
    if bucket:
        batch.order = batch.length_order()
    return batch


def hazard_batch(libs=None) -> ev.EvidenceBatch:
    """Hand-built order-sensitivity vectors: one site per case."""
    libs = libs or make_libraries(1)
    cases = [(30, 10), (100, 20), (700, 10), (1000, 10), (1000, 30), (10, 10), (1, 60), (2, 0),
             (5000, 60), (3000, 20)]
    sites, frags = [], []
    off = 0
    for nfr, q in cases:
        posA, length = 1_000_000, 4000
        # every fragment: ref-spanning pair at A whose first read also covers the breakpoint
        for j in range(nfr):
            a_start = posA - 60
            frags.append((a_start, a_start + READ_LEN, a_start + 220, a_start + 321, 0, 0,
                          q | (q << 8), ev.F_HAS_A | ev.F_HAS_B | ev.F_PAIRED | ev.F_REV_B))
        sites.append((posA, posA + length + 1, 0, 0, 0, 0, 0, 0, length, ev.SV_DEL | ev.SITE_O2_REV,
                      off, 0, nfr, 0, 0, 0))
        off += nfr
    return ev.EvidenceBatch(np.array(sites, dtype=np.int32), np.array(frags, dtype=np.int32),
                            np.zeros((0, ev.SPLIT_WORDS), np.int32), libs)


def concat_batches(parts, alloc=None, bucket=True) -> ev.EvidenceBatch:
    """Concatenate site-disjoint batches (same library table) into one, fixing row offsets.

    `alloc(name, shape, dtype)` may supply the destination arrays (e.g. pinned host
    memory); default is plain numpy.
    """
    alloc = alloc or (lambda name, shape, dtype: np.empty(shape, dtype=dtype))
    ns = sum(p.n_sites for p in parts)
    nf = sum(p.n_frag for p in parts)
    nsp = sum(p.n_split for p in parts)
    sites = alloc("sites", (ns, ev.SITE_WORDS), np.int32)
    frags = alloc("frags", (nf, ev.FRAG_WORDS), np.int32)
    splits = alloc("splits", (nsp, ev.SPLIT_WORDS), np.int32)
    s0 = f0 = p0 = 0
    for p in parts:
        s = p.sites.copy()
        foff = s[:, 10:12].copy().view(np.int64).ravel() + f0
        soff = s[:, 13:15].copy().view(np.int64).ravel() + p0
        s[:, 10:12] = foff.view(np.int32).reshape(-1, 2)
        s[:, 13:15] = soff.view(np.int32).reshape(-1, 2)
        sites[s0:s0 + p.n_sites] = s
        frags[f0:f0 + p.n_frag] = p.frags
        splits[p0:p0 + p.n_split] = p.splits
        s0, f0, p0 = s0 + p.n_sites, f0 + p.n_frag, p0 + p.n_split
    out = ev.EvidenceBatch.__new__(ev.EvidenceBatch)
    out.sites, out.frags, out.splits, out.libs, out.order = sites, frags, splits, parts[0].libs, None
    if bucket:
        order = alloc("order", (ns,), np.int32)
        order[:] = out.length_order()
        out.order = order
    return out


def _gen_chunk(args):
    config, n, seed, stream, path = args
    b = generate(config, n_sites=n, seed=seed, rank=stream, bucket=False)
    if path is None:
        return b.sites, b.frags, b.splits
    np.save(path + ".sites.npy", b.sites)
    np.save(path + ".frags.npy", b.frags)
    np.save(path + ".splits.npy", b.splits)
    return None


def generate_parallel(config="del1m4lib", n_sites=None, rank=0, chunk=25_000, procs=None, alloc=None,
                      bucket=True, site_range=None) -> ev.EvidenceBatch:
    """`generate()` over independent Philox streams (one per chunk), fanned out over a
    process pool; chunks travel through /dev/shm files, not pickles.  Stream ids are
    `rank * 65536 + chunk_index`, so every rank of a multi-GPU job draws a distinct shard.
    `site_range=(lo, hi)` returns only sites [lo, hi) of that batch (same bytes as slicing the
    whole thing), generating just the chunks that cover them: how the ranks of a site-sharded
    job each build their shard of ONE global batch without moving it."""
    import multiprocessing as mp
    import tempfile
    cfg = CONFIGS[config]
    n = int(cfg["n_sites"] if n_sites is None else n_sites)
    sizes = [min(chunk, n - i) for i in range(0, n, chunk)] or [0]
    seed = BASE_SEED + cfg["seed_off"]
    libs = make_libraries(cfg["n_lib"])
    if site_range is not None:
        lo, hi = int(site_range[0]), int(site_range[1])
        c0, c1 = lo // chunk, max(lo // chunk, (hi - 1) // chunk) if hi > lo else lo // chunk
        parts = []
        sel = [(i, sizes[i]) for i in range(c0, min(c1, len(sizes) - 1) + 1)] if hi > lo else []
        if sel:
            pr = procs or min(len(sel), max(1, (os.cpu_count() or 2) - 1), 32)
            tmp = tempfile.mkdtemp(prefix="svgt_synth_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            jobs = [(config, sz, seed, rank * 65536 + i, os.path.join(tmp, "c%05d" % i)) for i, sz in sel]
            try:
                if len(jobs) > 1 and pr > 1:
                    with mp.get_context("forkserver").Pool(pr) as pool:
                        pool.map(_gen_chunk, jobs, chunksize=1)
                else:
                    for j in jobs:
                        _gen_chunk(j)
                for j in jobs:
                    parts.append(ev.EvidenceBatch(np.load(j[4] + ".sites.npy"), np.load(j[4] + ".frags.npy"),
                                                  np.load(j[4] + ".splits.npy"), libs))
            finally:
                import shutil
                shutil.rmtree(tmp, ignore_errors=True)
        if not parts:
            return generate(config, n_sites=0, seed=seed, rank=rank * 65536, bucket=bucket, libs=libs)
        whole = concat_batches(parts, None, False)
        part = whole.slice_sites(lo - c0 * chunk, hi - c0 * chunk)
        return concat_batches([part], alloc, bucket)
    if len(sizes) == 1:
        return concat_batches([generate(config, n_sites=sizes[0], seed=seed, rank=rank * 65536, bucket=False)],
                              alloc, bucket)
    procs = procs or min(len(sizes), max(1, (os.cpu_count() or 2) - 1), 32)
    tmp = tempfile.mkdtemp(prefix="svgt_synth_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    jobs = [(config, sz, seed, rank * 65536 + i, os.path.join(tmp, "c%05d" % i)) for i, sz in enumerate(sizes)]
    try:
        with mp.get_context("forkserver").Pool(procs) as pool:
            pool.map(_gen_chunk, jobs, chunksize=1)
        parts = []
        for j in jobs:
            parts.append(ev.EvidenceBatch(np.load(j[4] + ".sites.npy", mmap_mode="r"),
                                          np.load(j[4] + ".frags.npy", mmap_mode="r"),
                                          np.load(j[4] + ".splits.npy", mmap_mode="r"), libs))
        return concat_batches(parts, alloc, bucket)
    finally:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
