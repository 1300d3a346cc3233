"""GPU: the CUDA path (through the C ABI of libsvgt.so) against the CPU oracle.

The default path scores COMPACT rows (svgt_score_compact: svgt_compact_kernel + svgt_call_compact_kernel);
the wide-row thread-per-site kernels behind svgt_score_batch are run beside it as an independent
cross-check.  Integer FORMAT fields and GT bit-exact; GL bit-exact against the oracle (same LUTs, same
IEEE operation order) and within 1e-6 of the reference's golden values; SQ within 1e-9.
"""
import os

import numpy as np
import pytest

from svtyper_b200 import compact as cp, evidence as ev, native, synth
from util import assert_rows_match, INT_FIELDS

pytestmark = pytest.mark.gpu

# "c0" / "c1" / "c2" / "c3": compact path with work units chosen from the row counts / fixed full units / fixed
# 2-site units / ramped units; "c9": choice left to the library (by site count) -- all without a piece plan;
# "p1" / "p3": compact path with sites longer than 1 / 3 chunks scored in pieces (svgt_segplan_t: SEG tally kernel +
# svgt_replay_pieces_kernel); "pa": the planner's own piece length; 0, 1: wide-row kernels
PATHS = ["c0", "c1", "c2", "c3", "c9", "p1", "p3", "pa", 0, 1]


@pytest.fixture(scope="module")
def eng():
    from svtyper_b200 import engine
    e = engine.Engine(0)
    yield e
    native.set_variant(-1)
    e.close()


def gpu_rows(eng, batch, path="c0", min_aligned=20, **kw):
    if isinstance(path, str):
        cb = batch if isinstance(batch, cp.CompactBatch) else cp.compact_from_wide(batch, min_aligned=min_aligned)
        if path[0] == "p":
            dev = eng.upload(cb, min_aligned=min_aligned, piece_chunks=None if path == "pa" else int(path[1:]), **kw)
            if path != "pa":
                assert dev.plan is not None and dev.plan_info["pieces"] > 0, "the batch has no site to cut: nothing tested"
        else:
            dev = eng.upload(cb, unit_mode=-1 if path == "c9" else int(path[1]), min_aligned=min_aligned, piece_chunks=0, **kw)
    else:
        native.set_variant(path)
        dev = eng.upload(batch, min_aligned=min_aligned, **kw)
    eng.score(dev)
    return eng.rows(dev)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("assoc", [ev.ASSOC_SSO, ev.ASSOC_CLASSIC])
def test_reference_fixture(eng, oracle, fixture_batch, fixture_npz, path, assoc):
    """211 breakpoints of the reference's own test data: golden values from the reference."""
    got = gpu_rows(eng, fixture_batch, path, assoc_mode=assoc)
    exp = fixture_npz["expected_sso"]
    for k in INT_FIELDS:
        assert np.array_equal(got[k], exp[k]), k
    assert np.allclose(got["GL"], exp["GL"], rtol=0, atol=1e-6)
    assert_rows_match(got, oracle.score(fixture_batch, assoc_mode=assoc), exact_gl=True, where="fixture")


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("config,n", [("del10k", 10_000), ("mixed100k", 20_000), ("del1m4lib", 20_000),
                                      ("stress1m", 6_000)])
def test_synthetic_configs(eng, oracle, config, n, path):
    b = synth.generate(config, n_sites=n)
    got = gpu_rows(eng, b, path)
    exp = oracle.score(b, n_threads=oracle.max_threads())
    assert_rows_match(got, exp, exact_gl=True, where=config)
    if config == "stress1m":
        assert (exp["GT"] == ev.GT_SKIPPED).any() and (exp["GT"] == ev.GT_BLANK).any()
        assert (exp["GT"] == ev.GT_UNDERFLOW).any()


@pytest.mark.parametrize("path", PATHS)
def test_hazard_vectors(eng, oracle, path):
    b = synth.hazard_batch()
    got = gpu_rows(eng, b, path)
    assert_rows_match(got, oracle.score(b), exact_gl=True, where="hazard")
    assert got["RS"][0] == 26 and got["RP"][0] == 12        # sequential, not tree, sums


@pytest.mark.parametrize("path", ["c0", "c2", "p2", 0])
def test_classic_association_and_weights(eng, oracle, path):
    b = synth.generate("mixed100k", n_sites=3000, seed=77)
    for assoc in (ev.ASSOC_SSO, ev.ASSOC_CLASSIC):
        for sw, dw in ((1.0, 1.0), (2.5, 0.5), (0.0, 1.0)):
            got = gpu_rows(eng, b, path, assoc_mode=assoc, split_weight=sw, disc_weight=dw)
            exp = oracle.score(b, assoc_mode=assoc, split_weight=sw, disc_weight=dw)
            assert_rows_match(got, exp, exact_gl=True, where="assoc%d w%s/%s" % (assoc, sw, dw))


@pytest.mark.parametrize("m", [0, 5, 35])
def test_other_min_aligned(eng, oracle, m):
    """-m changes the is_ref_seq windows, the straddle anchors and the packer-evaluated hit bits of gapped reads."""
    b = synth.generate("mixed100k", n_sites=3000, seed=13)
    assert_rows_match(gpu_rows(eng, b, "c0", min_aligned=m), oracle.score(b, min_aligned=m), exact_gl=True, where="m=%d" % m)
    cb = cp.compact_from_wide(b, min_aligned=m)
    dev = eng.upload(cb, min_aligned=m + 1)
    with pytest.raises(native.SvgtError) as ei:           # rows packed for another -m are refused, not mis-scored
        eng.score(dev)
    assert ei.value.code == native.ERR_ARG


def test_sites_at_the_contig_start(eng, oracle):
    """A breakend within min_aligned of position 0: its is_ref_seq window is cut short (parsers.py:812
    max(0, pos - m)) and can never be covered; straddles and splits are unaffected."""
    b = synth.generate("mixed100k", n_sites=2000, seed=41)
    s = b.sites.copy()
    f = b.frags.copy()
    q = b.splits.copy()
    for i in range(0, 2000, 3):                           # move every third site (and its rows) next to the origin
        shift = int(s[i, 0]) - (5 + i % 17)
        foff, nf = int(s[i, 10]), int(s[i, 12])
        soff, ns = int(s[i, 13]), int(s[i, 15])
        same = s[i, 6] == s[i, 7]
        s[i, 0] -= shift
        f[foff:foff + nf, 0:2] -= shift
        if same:
            s[i, 1] -= shift
            f[foff:foff + nf, 2:4] -= shift
            q[soff:soff + ns, 1:3] -= shift
            q[soff:soff + ns, 4:6] -= shift
    b2 = ev.EvidenceBatch(s, f, q, b.libs, b.order)
    exp = oracle.score(b2)
    for path in ("c0", "c2", 0):
        assert_rows_match(gpu_rows(eng, b2, path), exp, exact_gl=True, where="contig start " + str(path))


def test_identity_order_and_ragged_tail(eng, oracle):
    for n in (1, 7, 8, 9, 31, 32, 33, 1000):
        b = synth.generate("mixed100k", n_sites=n, seed=5, bucket=False)
        assert b.order is None
        assert_rows_match(gpu_rows(eng, b), oracle.score(b), exact_gl=True, where="n=%d" % n)


def test_empty_batch(eng):
    b = synth.generate("del10k", n_sites=0)
    assert gpu_rows(eng, b).shape == (0,)
    assert gpu_rows(eng, b, 0).shape == (0,)
    assert eng.score_host(cp.compact_from_wide(b)).shape == (0,)
    assert eng.score_host(b).shape == (0,)


@pytest.mark.parametrize("path", PATHS)
def test_literal_path_for_unsafe_library(eng, oracle, path):
    """flank = mean + 3 sd within 1e-9 of an integer: the integer window rewrite is not
    provably exact, so the kernel must take the literal fp64 comparisons."""
    mean, sd, hist = synth.fixture_library()
    libs = ev.LibraryTable([(299.999999999, 50.0, hist), (mean, sd, hist)])
    assert abs(libs.lib_f64[0, 0] - 450.0) < 1e-8 and libs.lib_f64[0, 0] != 450.0
    b = synth.generate("mixed100k", n_sites=4000, seed=11, libs=libs)
    assert_rows_match(gpu_rows(eng, b, path), oracle.score(b), exact_gl=True, where="unsafe lib")


def test_p_concordant_tie_and_integral_flank(eng, oracle):
    """19 * h1 == h2 (the one case the integer test cannot decide) and an integral
    mean + 3 sd that hits the histogram as a key."""
    hist = {i: 19 for i in range(100, 400)}
    hist.update({i: 1 for i in range(400, 1400)})
    hist.update({i: 361 for i in range(1400, 1500)})
    libs = ev.LibraryTable([(300.0, 50.0, hist)])           # flank 450.0 -> nondel_L = 450
    assert libs.lib_i32[0, 2] == 450
    for cfg in ("del10k", "mixed100k"):
        b = synth.generate(cfg, n_sites=4000, seed=3, libs=libs)
        assert_rows_match(gpu_rows(eng, b), oracle.score(b), exact_gl=True, where="tie " + cfg)


def test_many_libraries_and_large_histogram(eng, oracle):
    """> SVGT_SMEM_LIBS libraries (global-memory library rows), a histogram too large for the
    shared-memory copy, and histogram counts >= 2^26 (the 32-bit 19 * h1 > h2 does not apply)."""
    libs = ev.LibraryTable([synth.gaussian_library(300 + 7 * i, 40 + i) for i in range(70)])
    b = synth.generate("del1m4lib", n_sites=2000, seed=9, libs=libs)
    assert int(((b.frags[:, 6] >> 16) & 0xFFFF).max()) >= 64
    assert_rows_match(gpu_rows(eng, b), oracle.score(b), exact_gl=True, where="70 libs")
    big = ev.LibraryTable([synth.gaussian_library(4000, 900)])
    assert big.hist.size > 6144
    b = synth.generate("del10k", n_sites=2000, seed=10, libs=big)
    for path in PATHS:
        assert_rows_match(gpu_rows(eng, b, path), oracle.score(b), exact_gl=True, where="big hist")
    mean, sd, hist = synth.fixture_library()
    huge = ev.LibraryTable([(mean, sd, {k: v * 12000 for k, v in hist.items()})])
    assert int(huge.hist.max()) >= 1 << 26
    b = synth.generate("del10k", n_sites=1500, seed=12, libs=huge)
    assert_rows_match(gpu_rows(eng, b), oracle.score(b), exact_gl=True, where="counts >= 2^26")


def test_host_buffer_path_matches_device_path(eng, oracle):
    b = synth.generate("mixed100k", n_sites=5000, seed=21)
    cb = cp.compact_from_wide(b)
    got = eng.score_host(cb)
    assert_rows_match(got, oracle.score(b), exact_gl=True, where="host path")
    assert eng.last_h2d >= cb.sites.nbytes + cb.rows.nbytes
    assert eng.last_d2h >= b.n_sites * ev.OUT_BYTES
    native.set_variant(-1)
    assert eng.score_host(b).tobytes() == got.tobytes()                    # wide compatibility entry


def test_error_flags(eng):
    b = synth.generate("del10k", n_sites=500, seed=2)
    cb = cp.compact_from_wide(b)
    # library index out of range
    bad = cp.CompactBatch(cb.sites.copy(), cb.rows.copy(), cb.libs)
    nf0 = int(bad.sites[0, 10])
    bad.rows[:nf0, 3] |= 5 << 16
    dev = eng.upload(bad)
    eng.score(dev)
    with pytest.raises(native.SvgtError) as ei:
        eng.check(dev)
    assert ei.value.code == native.ERR_LIB_INDEX
    # coordinates outside +-2^30
    bad = cp.CompactBatch(cb.sites.copy(), cb.rows.copy(), cb.libs)
    bad.sites[3, 0] = (1 << 30) + 5
    dev = eng.upload(bad)
    eng.score(dev)
    with pytest.raises(native.SvgtError) as ei:
        eng.check(dev)
    assert ei.value.code == native.ERR_RANGE
    with pytest.raises(native.SvgtError):
        eng.score_host(bad)
    # a site whose rows lie outside the row array; a launch permutation naming a site that does not exist
    bad = cp.CompactBatch(cb.sites.copy(), cb.rows.copy(), cb.libs)
    bad.sites[7, 10] = cb.n_rows + 5
    dev = eng.upload(bad)
    eng.score(dev)
    with pytest.raises(native.SvgtError) as ei:
        eng.check(dev)
    assert ei.value.code == native.ERR_ARG
    bad = cp.CompactBatch(cb.sites.copy(), cb.rows.copy(), cb.libs, order=np.arange(cb.n_sites, dtype=np.int32))
    bad.order[11] = cb.n_sites + 3
    dev = eng.upload(bad)
    eng.score(dev)
    with pytest.raises(native.SvgtError) as ei:
        eng.check(dev)
    assert ei.value.code == native.ERR_ARG
    # log10 table too small for QR + QA
    dev = eng.upload(cb)
    dev.desc.n_log = 4
    eng.score(dev)
    with pytest.raises(native.SvgtError) as ei:
        eng.check(dev)
    assert ei.value.code == native.ERR_LOG_TABLE
    # the wide entry reports the same way
    wbad = ev.EvidenceBatch(b.sites.copy(), b.frags.copy(), b.splits.copy(), b.libs)
    wbad.frags[:, 6] |= 5 << 16
    native.set_variant(0)
    dev = eng.upload(wbad)
    eng.score(dev)
    with pytest.raises(native.SvgtError) as ei:
        eng.check(dev)
    assert ei.value.code == native.ERR_LIB_INDEX


def test_idempotent_and_permutation_invariant(eng):
    """Size-independent properties: rescoring gives identical bytes whatever the work-unit mapping or kernel;
    permuting the site rows (with their offsets) permutes the output rows."""
    b = synth.generate("stress1m", n_sites=3000, seed=8)
    r1 = gpu_rows(eng, b, "c0")
    for path in PATHS:
        assert gpu_rows(eng, b, path).tobytes() == r1.tobytes(), path
    cb = cp.compact_from_wide(b)
    perm = np.random.default_rng(1).permutation(b.n_sites)
    pb = cp.CompactBatch(cb.sites[perm], cb.rows, cb.libs)
    assert gpu_rows(eng, pb, "c0").tobytes() == r1[perm].tobytes()


def test_full_size_configs_against_oracle(eng, oracle):
    """BASELINE.json configs[1] and configs[2] at their full sizes (10k DEL, 100k mixed) against the oracle
    (all host threads): these sizes run the ramped AND the 8-site work units."""
    for config, n in (("del10k", 10_000), ("mixed100k", 100_000)):
        b = synth.generate_parallel(config, n_sites=n)
        got = gpu_rows(eng, b, "c0")
        exp = oracle.score(b, n_threads=oracle.max_threads())
        assert_rows_match(got, exp, exact_gl=True, where="full-size " + config)
        assert gpu_rows(eng, b, "c1").tobytes() == got.tobytes()          # 8-site units only: same bytes
        assert gpu_rows(eng, b, "pa").tobytes() == got.tobytes()          # long sites in pieces (the default upload)


def test_stress_shape_200k_sites_against_oracle(eng, oracle):
    """BASELINE.json configs[4] shape (max_reads=10000 ragged evidence, empty and skipped sites) at 200k sites:
    every row against the oracle on all host threads (reference singlesample.py:168-185 skip, :479-496 blank)."""
    b = synth.generate_parallel("stress1m", n_sites=200_000)
    got = gpu_rows(eng, b, "c0")
    exp = oracle.score(b, n_threads=oracle.max_threads())
    assert_rows_match(got, exp, exact_gl=True, where="stress1m 200k")
    assert gpu_rows(eng, b, "pa").tobytes() == got.tobytes()              # the default upload: long sites in pieces
    assert gpu_rows(eng, b, "p9").tobytes() == got.tobytes()
    gt = exp["GT"]
    assert (gt == ev.GT_SKIPPED).sum() > 100 and (gt == ev.GT_BLANK).sum() > 100 and (gt == ev.GT_UNDERFLOW).sum() > 100


def test_million_site_shape_properties(eng, oracle):
    """The benchmark shape (configs[3]) at 400k sites -- above any small-batch path: the unit mapping must not
    change a byte, rescoring is idempotent, the pipelined host path gives the same bytes, and two 16k-site
    slices equal the oracle."""
    b = synth.generate_parallel("del1m4lib", n_sites=400_000)
    cb = cp.compact_from_wide(b)
    r = gpu_rows(eng, cb, "c0")
    assert gpu_rows(eng, cb, "c0").tobytes() == r.tobytes()
    assert gpu_rows(eng, cb, "c1").tobytes() == r.tobytes()
    assert gpu_rows(eng, cb, "c2").tobytes() == r.tobytes()
    assert eng.score_host(cb).tobytes() == r.tobytes()                    # >= 131072 sites: 8 pipelined slices
    for lo, hi in ((123_000, 139_000), (384_000, 400_000)):
        part = b.slice_sites(lo, hi)
        assert_rows_match(r[lo:hi], oracle.score(part, n_threads=oracle.max_threads()), exact_gl=True, where="1M-shape slice")


def test_pipelined_host_path(tmp_path):
    """svgt_ctx_score_host_compact overlaps H2D / kernels / D2H over site slices for large batches laid out in
    site order.  Forced on for a small batch in a child process (the threshold is read once per process): same
    bytes as the device path and the oracle; a batch NOT in site order is refused when it claims to be, and
    scored in one shot when it does not."""
    import subprocess
    import sys
    code = r'''
import sys
sys.path.insert(0, %r)
import numpy as np
from svtyper_b200 import compact as cp, engine, evidence as ev, native, synth
from oracle import oracle
eng = engine.Engine(0)
b = synth.generate("mixed100k", n_sites=6000, seed=31)
cb = cp.compact_from_wide(b)
want = oracle.score(b)
dev = eng.upload(cb); eng.score(dev); ref = eng.rows(dev)
got = eng.score_host(cb)
assert got.tobytes() == ref.tobytes()
for k in ("GT", "GQ", "DP", "RO", "AO", "QR", "QA", "RS", "AS", "ASC", "RP", "AP"):
    assert np.array_equal(got[k], want[k]), k
assert eng.last_h2d >= cb.sites.nbytes + cb.rows.nbytes
perm = np.random.default_rng(3).permutation(cb.n_sites)
pb = cp.CompactBatch(cb.sites[perm], cb.rows, cb.libs)                  # offsets no longer monotonic
try:
    eng.score_host(pb)
    raise SystemExit("a batch that is not in site order was accepted as such")
except native.SvgtError as e:
    assert e.code == native.ERR_ARG
assert eng.score_host(pb, site_order=False).tobytes() == ref[perm].tobytes()
e5 = cp.CompactBatch(cb.sites[:5], cb.rows, cb.libs)                     # fewer sites than slices: one shot
assert eng.score_host(e5).tobytes() == ref[:5].tobytes()
print("pipelined ok")
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),)
    env = dict(os.environ, SVGT_PIPELINE_MIN_SITES="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "pipelined ok" in r.stdout, r.stderr[-1500:]


def test_torch_operator_matches_engine(eng, oracle):
    """torch.ops.svgt.score_batch (device tensors in / out, current stream) gives the engine's bytes."""
    import torch
    from svtyper_b200 import engine, torch_op  # noqa: F401
    b = synth.generate("mixed100k", n_sites=3000, seed=17)
    cb = cp.compact_from_wide(b)
    ref = gpu_rows(eng, cb, "c1")
    arrs = engine.host_arrays(cb)
    dev = {k: torch.from_numpy(v.view(np.int32) if v.dtype == np.uint32 else v).cuda() for k, v in arrs.items()}
    out, status = torch.ops.svgt.score_batch(dev["sites"], dev["rows"], dev["order"], dev["lib_f64"], dev["lib_i32"], dev["hist"],
                                             dev["pm"], dev["logt"], dev["consts"], 1.0, 1.0, 20, 3, ev.ASSOC_SSO, 1)
    torch.cuda.synchronize()
    assert int(status[0]) == 0
    assert out.cpu().numpy().reshape(-1).view(ev.OUT_DTYPE).tobytes() == ref.tobytes()


def test_piece_plan_is_bit_identical_and_covers_the_host_path(eng, oracle, monkeypatch):
    """A heavy-tailed batch (config stress1m: up to 10,000 reads per site) scored with its long sites cut into pieces:
    every piece length gives the bytes of the unsplit path (and the oracle's values), through svgt_score_compact and
    through svgt_ctx_score_host_compact (which plans by itself); pieces really were used."""
    b = synth.generate("stress1m", n_sites=4000, seed=5)
    cb = cp.compact_from_wide(b)
    exp = oracle.score(b, n_threads=oracle.max_threads())
    for assoc in (ev.ASSOC_SSO, ev.ASSOC_CLASSIC):
        want = oracle.score(b, assoc_mode=assoc, n_threads=oracle.max_threads()) if assoc != ev.ASSOC_SSO else exp
        base = gpu_rows(eng, cb, "c0", assoc_mode=assoc)
        assert_rows_match(base, want, exact_gl=True, where="no plan")
        for k in (1, 2, 7, 40, None):
            dev = eng.upload(cb, assoc_mode=assoc, piece_chunks=k)
            assert dev.plan is not None and dev.plan_info["heavy_sites"] > 0
            eng.score(dev)
            got = eng.rows(dev)
            assert got.tobytes() == base.tobytes(), "pieces of <= %s chunks, assoc %d" % (k, assoc)
    monkeypatch.setenv("SVGT_PLAN_FORCE_CHUNKS", "2")
    host = eng.score_host(cb)
    assert eng.last_pieces > 0
    assert host.tobytes() == gpu_rows(eng, cb, "c0").tobytes()
    monkeypatch.delenv("SVGT_PLAN_FORCE_CHUNKS")
    host = eng.score_host(cb)                       # no plan unless asked for (SVGT_PLAN=1 / FORCE_CHUNKS)
    assert eng.last_pieces == 0
    assert_rows_match(host, exp, exact_gl=True, where="host path")


def test_bad_piece_plan_is_flagged(eng):
    """Entries / pieces that do not describe the batch raise SVGT_ERR_ARG instead of reading out of bounds."""
    import torch
    b = synth.generate("del10k", n_sites=500, seed=3)
    cb = cp.compact_from_wide(b)
    for what in ("entry", "piece_rows", "piece_scratch", "heavy"):
        dev = eng.upload(cb, piece_chunks=1)
        assert dev.plan is not None
        if what == "entry":
            dev.tensors["plan_entries"][0] = cb.n_sites + 5
        elif what == "piece_rows":
            dev.tensors["plan_pieces"][0, 2] = 1 << 20
        elif what == "piece_scratch":
            dev.tensors["plan_pieces"][0, 3] = dev.plan.scratch_chunks
        else:
            dev.tensors["plan_heavy"][0, 2] += 1
        dev.tensors["plan_scratch"].fill_(255)      # whatever an unscored piece leaves behind: lead counts of -1, NaN addends
        torch.cuda.synchronize()
        eng.score(dev)
        with pytest.raises(native.SvgtError) as ei:
            eng.rows(dev)
        assert ei.value.code == native.ERR_ARG, what
