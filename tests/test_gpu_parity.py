"""GPU: the CUDA path (through the C ABI of libsvgt.so) against the CPU oracle.

Integer FORMAT fields and GT bit-exact; GL bit-exact against the oracle (same LUTs, same
IEEE operation order) and within 1e-6 of the reference's golden values; SQ within 1e-9.
"""
import os

import numpy as np
import pytest

from svtyper_b200 import evidence as ev, native, synth
from util import assert_rows_match, INT_FIELDS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from svtyper_b200 import engine
    e = engine.Engine(0)
    yield e
    native.set_variant(-1)
    e.close()


def gpu_rows(eng, batch, variant=5, **kw):
    native.set_variant(variant)
    dev = eng.upload(batch, **kw)
    eng.score(dev)
    return eng.rows(dev)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("assoc", [ev.ASSOC_SSO, ev.ASSOC_CLASSIC])
def test_reference_fixture(eng, oracle, fixture_batch, fixture_npz, variant, assoc):
    """211 breakpoints of the reference's own test data: golden values from the reference."""
    got = gpu_rows(eng, fixture_batch, variant, assoc_mode=assoc)
    exp = fixture_npz["expected_sso"]
    for k in INT_FIELDS:
        assert np.array_equal(got[k], exp[k]), k
    assert np.allclose(got["GL"], exp["GL"], rtol=0, atol=1e-6)
    assert_rows_match(got, oracle.score(fixture_batch, assoc_mode=assoc), exact_gl=True, where="fixture")


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("config,n", [("del10k", 10_000), ("mixed100k", 20_000), ("del1m4lib", 20_000),
                                      ("stress1m", 6_000)])
def test_synthetic_configs(eng, oracle, config, n, variant):
    b = synth.generate(config, n_sites=n)
    got = gpu_rows(eng, b, variant)
    exp = oracle.score(b, n_threads=oracle.max_threads())
    assert_rows_match(got, exp, exact_gl=True, where=config)
    if config == "stress1m":
        assert (exp["GT"] == ev.GT_SKIPPED).any() and (exp["GT"] == ev.GT_BLANK).any()
        assert (exp["GT"] == ev.GT_UNDERFLOW).any()


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6, 7])
def test_hazard_vectors(eng, oracle, variant):
    b = synth.hazard_batch()
    got = gpu_rows(eng, b, variant)
    assert_rows_match(got, oracle.score(b), exact_gl=True, where="hazard")
    assert got["RS"][0] == 26 and got["RP"][0] == 12        # sequential, not tree, sums


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6, 7])
def test_classic_association_and_weights(eng, oracle, variant):
    b = synth.generate("mixed100k", n_sites=3000, seed=77)
    for assoc in (ev.ASSOC_SSO, ev.ASSOC_CLASSIC):
        for sw, dw in ((1.0, 1.0), (2.5, 0.5), (0.0, 1.0)):
            got = gpu_rows(eng, b, variant, assoc_mode=assoc, split_weight=sw, disc_weight=dw)
            exp = oracle.score(b, assoc_mode=assoc, split_weight=sw, disc_weight=dw)
            assert_rows_match(got, exp, exact_gl=True, where="assoc%d w%s/%s" % (assoc, sw, dw))


def test_identity_order_and_ragged_tail(eng, oracle):
    for n in (1, 31, 32, 33, 1000):
        b = synth.generate("mixed100k", n_sites=n, seed=5, bucket=False)
        assert b.order is None
        assert_rows_match(gpu_rows(eng, b), oracle.score(b), exact_gl=True, where="n=%d" % n)


def test_empty_batch(eng):
    b = synth.generate("del10k", n_sites=0)
    assert gpu_rows(eng, b).shape == (0,)
    assert eng.score_host(b).shape == (0,)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6, 7])
def test_literal_path_for_unsafe_library(eng, oracle, variant):
    """flank = mean + 3 sd within 1e-9 of an integer: the integer window rewrite is not
    provably exact, so the kernel must take the literal fp64 comparisons."""
    mean, sd, hist = synth.fixture_library()
    libs = ev.LibraryTable([(299.999999999, 50.0, hist), (mean, sd, hist)])
    assert abs(libs.lib_f64[0, 0] - 450.0) < 1e-8 and libs.lib_f64[0, 0] != 450.0
    b = synth.generate("mixed100k", n_sites=4000, seed=11, libs=libs)
    assert_rows_match(gpu_rows(eng, b, variant), oracle.score(b), exact_gl=True, where="unsafe lib")


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
    """> SVGT_SMEM_LIBS libraries (global-memory library rows) and a histogram too large
    for the shared-memory copy."""
    libs = ev.LibraryTable([synth.gaussian_library(300 + 7 * i, 40 + i) for i in range(70)])
    b = synth.generate("del1m4lib", n_sites=2000, seed=9, libs=libs)
    assert int(((b.frags[:, 6] >> 16) & 0xFFFF).max()) >= 64
    assert_rows_match(gpu_rows(eng, b), oracle.score(b), exact_gl=True, where="70 libs")
    big = ev.LibraryTable([synth.gaussian_library(4000, 900)])
    assert big.hist.size > 6144
    b = synth.generate("del10k", n_sites=2000, seed=10, libs=big)
    for v in (0, 1, 2, 3, 4, 5, 6, 7):
        assert_rows_match(gpu_rows(eng, b, v), oracle.score(b), exact_gl=True, where="big hist")


def test_host_buffer_path_matches_device_path(eng, oracle):
    b = synth.generate("mixed100k", n_sites=5000, seed=21)
    native.set_variant(-1)
    got = eng.score_host(b)
    assert_rows_match(got, oracle.score(b), exact_gl=True, where="host path")
    assert eng.last_h2d >= b.sites.nbytes + b.frags.nbytes + b.splits.nbytes
    assert eng.last_d2h >= b.n_sites * ev.OUT_BYTES


def test_error_flags(eng):
    b = synth.generate("del10k", n_sites=500, seed=2)
    # library index out of range
    bad = ev.EvidenceBatch(b.sites.copy(), b.frags.copy(), b.splits.copy(), b.libs)
    bad.frags[:, 6] |= 5 << 16
    dev = eng.upload(bad)
    eng.score(dev)
    with pytest.raises(native.SvgtError) as ei:
        eng.check(dev)
    assert ei.value.code == native.ERR_LIB_INDEX
    # coordinates outside +-2^30
    bad = ev.EvidenceBatch(b.sites.copy(), b.frags.copy(), b.splits.copy(), b.libs)
    bad.sites[3, 0] = (1 << 30) + 5
    dev = eng.upload(bad)
    eng.score(dev)
    with pytest.raises(native.SvgtError) as ei:
        eng.check(dev)
    assert ei.value.code == native.ERR_RANGE
    # log10 table too small for QR + QA
    dev = eng.upload(b)
    dev.desc.n_log = 4
    eng.score(dev)
    with pytest.raises(native.SvgtError) as ei:
        eng.check(dev)
    assert ei.value.code == native.ERR_LOG_TABLE
    with pytest.raises(native.SvgtError):
        eng.score_host(bad)


def test_idempotent_and_permutation_invariant(eng):
    """Size-independent properties: rescoring gives identical bytes; permuting the site
    rows (with their offsets) permutes the output rows."""
    b = synth.generate("stress1m", n_sites=3000, seed=8)
    r1 = gpu_rows(eng, b, 0)
    r2 = gpu_rows(eng, b, 1)
    assert r1.tobytes() == r2.tobytes()
    assert r1.tobytes() == gpu_rows(eng, b, 2).tobytes()
    assert r1.tobytes() == gpu_rows(eng, b, 3).tobytes()
    assert r1.tobytes() == gpu_rows(eng, b, 4).tobytes()
    perm = np.random.default_rng(1).permutation(b.n_sites)
    pb = ev.EvidenceBatch(b.sites[perm], b.frags, b.splits, b.libs)
    r3 = gpu_rows(eng, pb, 2)
    assert r3.tobytes() == r1[perm].tobytes()


def test_full_size_configs_against_oracle(eng, oracle):
    """BASELINE.json configs[1] and configs[2] at their full sizes (10k DEL, 100k mixed) against the oracle
    (all host threads), default kernel: these sizes run the ramped AND the 8-site work units."""
    for config, n in (("del10k", 10_000), ("mixed100k", 100_000)):
        b = synth.generate_parallel(config, n_sites=n)
        got = gpu_rows(eng, b, 5)
        exp = oracle.score(b, n_threads=oracle.max_threads())
        assert_rows_match(got, exp, exact_gl=True, where="full-size " + config)
        assert gpu_rows(eng, b, 6).tobytes() == got.tobytes()          # 8-site units only: same bytes


def test_million_site_shape_properties(eng, oracle):
    """The benchmark shape (configs[3]) at 400k sites -- above any small-batch path: the unit mapping must not
    change a byte (variants 5 / 6 / 7), rescoring is idempotent, and a 16k-site slice equals the oracle."""
    b = synth.generate_parallel("del1m4lib", n_sites=400_000)
    r5 = gpu_rows(eng, b, 5)
    assert gpu_rows(eng, b, 5).tobytes() == r5.tobytes()
    assert gpu_rows(eng, b, 6).tobytes() == r5.tobytes()
    assert gpu_rows(eng, b, 7).tobytes() == r5.tobytes()
    assert gpu_rows(eng, b, 2).tobytes() == r5.tobytes()               # the previous default kernel
    lo, hi = 123_000, 139_000
    part = b.slice_sites(lo, hi)
    assert_rows_match(r5[lo:hi], oracle.score(part, n_threads=oracle.max_threads()), exact_gl=True, where="1M-shape slice")


def test_pipelined_host_path(tmp_path):
    """svgt_ctx_score_host overlaps H2D / kernels / D2H over site slices for large batches.  Forced on for a
    small batch in a child process (the threshold is read once per process): same bytes as the device path
    and the oracle; a batch whose rows are NOT laid out in site order must fall back and still be right."""
    import subprocess
    import sys
    code = r'''
import sys
sys.path.insert(0, %r)
import numpy as np
from svtyper_b200 import engine, evidence as ev, native, synth
from oracle import oracle
eng = engine.Engine(0)
b = synth.generate("mixed100k", n_sites=6000, seed=31)
want = oracle.score(b)
dev = eng.upload(b); eng.score(dev); ref = eng.rows(dev)
got = eng.score_host(b)
assert got.tobytes() == ref.tobytes()
for k in ("GT", "GQ", "DP", "RO", "AO", "QR", "QA", "RS", "AS", "ASC", "RP", "AP"):
    assert np.array_equal(got[k], want[k]), k
assert eng.last_h2d >= b.sites.nbytes + b.frags.nbytes + b.splits.nbytes
perm = np.random.default_rng(3).permutation(b.n_sites)
pb = ev.EvidenceBatch(b.sites[perm], b.frags, b.splits, b.libs)        # offsets no longer monotonic
got2 = eng.score_host(pb)
assert got2.tobytes() == ref[perm].tobytes()
e = ev.EvidenceBatch(b.sites[:5], b.frags, b.splits, b.libs)           # fewer sites than slices
assert eng.score_host(e).tobytes() == ref[:5].tobytes()
print("pipelined ok")
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),)
    env = dict(os.environ, SVGT_PIPELINE_MIN_SITES="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "pipelined ok" in r.stdout, r.stderr[-1500:]
