"""CPU: the native evidence packer (libsvgt_pack.so) against its parity checker,
the Python gather path (gather.py + evidence.BatchPacker, itself pinned to the reference through the
golden VCF), on the reference's own fixture BAM: identical site / fragment / split rows for both
gather modes, with and without the too-many-reads limit; pysam-style count(); error conventions."""
import ctypes
import os
import re

import numpy as np
import pytest

from svtyper_b200 import evidence as ev
from svtyper_b200 import bamio, gather, genotype, packer, vcf
from svtyper_b200.sample import SampleInfo
from util import REPO

DATA = os.path.join(REPO, "tests", "data")
BAM = os.path.join(DATA, "NA12878.target_loci.sorted.bam")
VCF = os.path.join(DATA, "example.vcf")
LIB = os.path.join(DATA, "NA12878.bam.json")


def make_plan():
    lines = open(VCF).read().splitlines()
    header = vcf.VcfHeader().parse([l for l in lines if l.startswith("##")])
    header.ensure_svtyper_fields()
    header.add_sample("NA12878")
    plan = genotype.SitePlan()
    open_bnds = {}
    for line in lines:
        if line.startswith("#"):
            continue
        rec = vcf.VcfRecord(line.rstrip().split("\t"), header)
        if not rec.has_svtype() or rec.svtype() not in ("BND", "DEL", "DUP", "INV"):
            plan.passthrough(rec)
            continue
        if rec.svtype() == "BND":
            mate_id = rec.info["MATEID"]
            if mate_id not in open_bnds:
                open_bnds[rec.var_id] = rec
                continue
            first = open_bnds.pop(mate_id)
            plan.site(first, rec, vcf.bnd_breakpoint(first, rec, 1e10))
        else:
            plan.site(rec, None, vcf.simple_breakpoint(rec, 1e10))
    return plan


@pytest.fixture(scope="module")
def plan():
    return make_plan()


@pytest.fixture()
def sample():
    s = SampleInfo.open(BAM, LIB, None, 1000000)
    yield s
    s.close()


def test_abi_exports_every_declared_symbol():
    L = packer.lib()
    assert L.svgt_pack_abi_version() == 2
    header = open(os.path.join(REPO, "include", "svgt_pack.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = set(re.findall(r"\b(svgt_(?:pack|bam|compact|format|host)_[a-z_]+)\s*\(", header))
    assert {"svgt_bam_open", "svgt_bam_close", "svgt_bam_count", "svgt_pack_sites", "svgt_pack_rows",
            "svgt_compact_count", "svgt_compact_fill", "svgt_format_calls", "svgt_format_quals"} <= names
    for n in names:
        assert hasattr(L, n), n


@pytest.mark.parametrize("mode,max_reads", [(packer.MODE_SSO, 1000), (packer.MODE_SSO, None), (packer.MODE_SSO, 150),
                                            (packer.MODE_CLASSIC, None), (packer.MODE_CLASSIC, 1000),
                                            (packer.MODE_CLASSIC, 120)])
def test_rows_identical_to_python_gather(sample, plan, mode, max_reads):
    if mode == packer.MODE_SSO:
        g = lambda smp, bp: gather.gather_sso(smp, bp, genotype.Z, max_reads)
    else:
        g = lambda smp, bp: gather.gather_classic(smp, bp, genotype.Z, max_reads)
    want = genotype.pack_sample_python(sample, plan, g, 20)
    got = packer.pack_sample(sample, plan, mode, max_reads, genotype.Z)
    assert got.n_sites == want.n_sites == len(plan.breakpoints) == 211
    assert np.array_equal(got.sites, want.sites)
    assert np.array_equal(got.frags, want.frags)
    assert np.array_equal(got.splits, want.splits)
    assert np.array_equal(got.order, want.order)
    if max_reads is not None and max_reads < 200:
        assert (got.sites[:, 9] & ev.SITE_SKIP).any() and not (got.sites[:, 9] & ev.SITE_SKIP).all()
    assert (got.frags[:, 7] & ev.F_EXTRA).any() and got.n_split > 0      # gapped reads and split candidates occur


def test_thread_count_does_not_change_the_rows(sample, plan):
    one = packer.pack_sample(sample, plan, packer.MODE_SSO, 1000, genotype.Z, threads=1)
    for th in (2, 5, 0):
        many = packer.pack_sample(sample, plan, packer.MODE_SSO, 1000, genotype.Z, threads=th)
        assert np.array_equal(many.sites, one.sites) and np.array_equal(many.frags, one.frags)
        assert np.array_equal(many.splits, one.splits)


@pytest.mark.parametrize("num_samp", [1000000, 5000, 37])
def test_library_scan_matches_python_passes(num_samp):
    """svgt_bam_scan_libraries (one pass) against the three Python passes per library, on the fixture
    and on a constructed two-library BAM: same read length, prevalence, histogram (values AND key
    order), hence the same mean / sd and the same -l JSON."""
    import json
    bam = bamio.AlignmentFile(BAM)
    want = SampleInfo.from_bam(bam, num_samp, native=False)
    got = SampleInfo.from_bam(bam, num_samp, native=True)
    assert json.dumps(got.to_json()) == json.dumps(want.to_json())
    for a, b in zip(got.libraries, want.libraries):
        assert list(a.hist.items()) == list(b.hist.items()) and a.mean == b.mean and a.sd == b.sd
        assert a.read_length == b.read_length and a.prevalence == b.prevalence and a.readgroups == b.readgroups
    bam.close()


def test_library_scan_two_libraries(tmp_path):
    path, _ = _synthetic_bam(tmp_path, 11)
    bam = bamio.AlignmentFile(path)
    for num_samp in (100000, 50):
        want = SampleInfo.from_bam(bam, num_samp, native=False)
        got = SampleInfo.from_bam(bam, num_samp, native=True)
        assert len(got.libraries) == 2
        for a, b in zip(got.libraries, want.libraries):
            assert list(a.hist.items()) == list(b.hist.items()) and (a.mean, a.sd) == (b.mean, b.sd)
            assert (a.read_length, a.prevalence, a.name) == (b.read_length, b.prevalence, b.name)
    bam.close()


def test_count_matches_python_reader():
    nb = packer.NativeBam(BAM)
    pb = bamio.AlignmentFile(BAM)
    assert nb.references == pb.references and nb.lengths == pb.lengths
    rng = np.random.default_rng(7)
    plan = make_plan()
    for bp in plan.breakpoints[::9]:
        tid = pb.gettid(bp["A"]["chrom"])
        pos = bp["A"]["pos"]
        for _ in range(2):
            lo = max(0, pos - int(rng.integers(1, 3000)))
            hi = pos + int(rng.integers(1, 3000))
            for cb in ("all", "nofilter"):
                assert nb.count(tid, lo, hi, cb) == pb.count(pb.references[tid], lo, hi, read_callback=cb)
    assert nb.count(0, 100, 100) == 0
    nb.close()
    pb.close()


def test_errors_are_codes_not_crashes(tmp_path, sample, plan):
    with pytest.raises(packer.PackError) as e:
        packer.NativeBam(str(tmp_path / "missing.bam"))
    assert e.value.code == packer.ERR_IO
    bad = tmp_path / "bad.bam"
    bad.write_bytes(b"not a bam at all" * 10)
    with pytest.raises(packer.PackError):
        packer.NativeBam(str(bad))
    # a read group the table does not list is an error, like the reference's KeyError
    nb = packer.NativeBam(BAM)
    site = packer.fetch_windows(sample, plan.breakpoints[0], genotype.Z, packer.MODE_SSO)
    with pytest.raises(packer.PackError) as e:
        nb.pack([site], ["nope"], [0], [True], packer.MODE_SSO, None)
    assert e.value.code == packer.ERR_RG
    with pytest.raises(packer.PackError) as e:
        nb.pack([site], [], [], [True], 7, None)
    assert e.value.code == packer.ERR_ARG
    nb.close()


def test_entry_points_use_the_native_packer(sample, plan, monkeypatch):
    assert packer.usable(sample)
    monkeypatch.setenv("SVGT_PACKER", "python")
    assert not packer.usable(sample)


# ---- constructed edge cases: a synthetic BAM read by both readers -------------------------------------
def _synthetic_bam(tmp_path, seed):
    """Reads around two breakends per site with everything the gather rules branch on: gapped reads
    (D / N), soft and hard clips, SA tags (one / several entries, other contigs, unknown contig),
    fragments with 1, 2 and 3+ primaries, duplicate / secondary / supplementary / unmapped records, a
    repeated (name, flag) record, and an inactive library."""
    import bamwriter
    rng = np.random.default_rng(seed)
    refs = [("chrA", 400000), ("chrB", 300000)]
    header = ("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:chrA\tLN:400000\n@SQ\tSN:chrB\tLN:300000\n"
              "@RG\tID:rg1\tSM:S1\tLB:libX\n@RG\tID:rg2\tSM:S1\tLB:libX\n@RG\tID:rg3\tSM:S1\tLB:libY\n")
    sites, recs = [], []
    cigars = ["100M", "60M5D40M", "30M200N70M", "20S80M", "75M25S", "10H90M", "40M2I58M", "15S60M3D10M15S",
              "50M50H", "100M"]
    for s in range(40):
        posA = int(rng.integers(2000, 180000)) + 2000 * s
        length = int(rng.integers(300, 6000))
        inter = rng.random() < 0.2
        tidB, posB = (1, int(rng.integers(2000, 250000))) if inter else (0, posA + length)
        sites.append((0, posA, tidB, posB))
        for f in range(int(rng.integers(3, 40))):
            qname = "q%03d_%03d" % (s, int(rng.integers(0, 500)))          # collisions on purpose
            n_prim = int(rng.choice([1, 2, 2, 2, 2, 3, 4]))
            rg = str(rng.choice(["rg1", "rg2", "rg3"]))
            for k in range(n_prim):
                tid, anchor = (0, posA) if (k % 2 == 0 or rng.random() < 0.3) else (tidB, posB)
                pos = max(0, anchor + int(rng.integers(-700, 700)))
                flag = 0x1 | (0x10 if rng.random() < 0.5 else 0) | (0x40 if k == 0 else 0x80)
                r = rng.random()
                if r < 0.04:
                    flag |= 0x400
                elif r < 0.07:
                    flag |= 0x100
                elif r < 0.10:
                    flag |= 0x800
                elif r < 0.12:
                    flag |= 0x4
                cigar = str(rng.choice(cigars))
                tags = [("RG", "Z", rg)]
                r = rng.random()
                if r < 0.25:
                    sa_chr = str(rng.choice(["chrA", "chrA", "chrB", "chrUn"]))
                    sa = "%s,%d,%s,%s,%d,0;" % (sa_chr, max(1, pos + int(rng.integers(-3000, 3000))),
                                                rng.choice(["+", "-"]), rng.choice(["60S40M", "45M55S", "30S30M2D40M"]),
                                                int(rng.integers(0, 61)))
                    if rng.random() < 0.2:
                        sa += "chrB,500,+,50M50S,20,1;"
                    tags.append(("SA", "Z", sa))
                tags.append(("NM", "i", int(rng.integers(0, 5))))
                if rng.random() < 0.6:
                    flag |= 0x20                                           # mate on the reverse strand
                rec = dict(tid=tid, pos=pos, qname=qname, flag=flag, mapq=int(rng.integers(0, 61)), cigar=cigar,
                           l_seq=int(rng.choice([0, 100])), tags=tags, tlen=int(rng.integers(-200, 900)))
                recs.append(rec)
                if rng.random() < 0.05:
                    recs.append(dict(rec))                                 # the same (name, flag) twice
    recs.sort(key=lambda r: (r["tid"], r["pos"]))
    path = str(tmp_path / ("synth%d.bam" % seed))
    bamwriter.write_bam(path, refs, header, recs)
    return path, sites


class _Lib(object):
    def __init__(self, name, rgs, mean, sd, prevalence):
        self.name, self.readgroups, self.mean, self.sd, self.prevalence = name, rgs, mean, sd, prevalence
        self.hist = {int(mean) + d: 10 for d in range(-50, 51)}


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_constructed_edge_cases_match_python_gather(tmp_path, seed):
    from svtyper_b200.sample import SampleInfo
    path, sites = _synthetic_bam(tmp_path, seed)
    bam = bamio.AlignmentFile(path)
    libs = [_Lib("libX", ["rg1", "rg2"], 300.0, 40.0, 0.9), _Lib("libY", ["rg3"], 420.0, 60.0, 0.9)]
    for inactive in (False, True):
        if inactive:
            libs[1].prevalence = 1e-9                                    # below MIN_LIB_PREVALENCE: ignored
        smp = SampleInfo("S1", bam, libs, 0, 0)
        plan = genotype.SitePlan()
        for tidA, posA, tidB, posB in sites:
            bp = {"id": "x", "svtype": "BND" if tidB != tidA else "DEL", "var_length": posB - posA,
                  "A": {"chrom": bam.references[tidA], "pos": posA, "ci": [-5, 7], "is_reverse": False},
                  "B": {"chrom": bam.references[tidB], "pos": posB, "ci": [0, 0], "is_reverse": True}}
            plan.breakpoints.append(bp)
        for mode, mr in ((packer.MODE_SSO, None), (packer.MODE_SSO, 25), (packer.MODE_CLASSIC, None),
                         (packer.MODE_CLASSIC, 18)):
            if mode == packer.MODE_SSO:
                g = lambda s_, bp_: gather.gather_sso(s_, bp_, genotype.Z, mr)
            else:
                g = lambda s_, bp_: gather.gather_classic(s_, bp_, genotype.Z, mr)
            want = genotype.pack_sample_python(smp, plan, g, 20)
            got = packer.pack_sample(smp, plan, mode, mr, genotype.Z, threads=3)
            where = "seed %d mode %d max_reads %s inactive %s" % (seed, mode, mr, inactive)
            assert np.array_equal(got.sites, want.sites), where
            assert np.array_equal(got.frags, want.frags), where
            assert np.array_equal(got.splits, want.splits), where
            if mr is None:
                fl = got.frags[:, 7]
                assert (fl & ev.F_CONT).any() and (fl & ev.F_EXTRA).any() and (fl & ev.F_PAIRED).any(), where
                assert got.n_split > 0 and ((got.splits[:, 6] >> 16) & ev.S_SOFT_CLIP).any(), where
            else:
                assert (got.sites[:, 9] & ev.SITE_SKIP).any(), where
    bam.close()


def test_corrupted_files_never_crash(tmp_path):
    """The native reader parses untrusted files: random byte flips / truncations of the fixture BAM and BAI
    must end in a result or a PackError, never in a crash (run in a child process so a crash is visible)."""
    import subprocess
    import sys
    code = r'''
import os, random, sys
sys.path.insert(0, %r)
from svtyper_b200 import packer
src = %r
bam = open(src, "rb").read(); bai = open(src + ".bai", "rb").read()
random.seed(20261017)
d = %r
for it in range(60):
    b, i = bytearray(bam), bytearray(bai)
    mode = it %% 4
    if mode == 0:
        for _ in range(random.randint(1, 20)): b[random.randrange(len(b))] = random.randrange(256)
    elif mode == 1:
        b = b[:random.randrange(1, len(b))]
    elif mode == 2:
        for _ in range(random.randint(1, 20)): i[random.randrange(len(i))] = random.randrange(256)
    else:
        i = i[:random.randrange(1, len(i))]
    p = os.path.join(d, "f.bam")
    open(p, "wb").write(bytes(b)); open(p + ".bai", "wb").write(bytes(i))
    try:
        nb = packer.NativeBam(p)
        sites = [(0, 1000000 * k %% 200000000, 1000000 * k %% 200000000 + 3000, 0, 5000 * k, 5000 * k + 2000) for k in range(1, 20)]
        try:
            nb.pack(sites, ["x"], [0], [True], it %% 2, None if it %% 3 else 50, threads=1 + it %% 3)
            nb.count(0, 0, 250000000, "all")
        except packer.PackError:
            pass
        nb.close()
    except packer.PackError:
        pass
print("survived")
''' % (REPO, BAM, str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "survived" in r.stdout, (r.returncode, r.stderr[-500:])


def test_native_compact_converter_and_formatter_match_python(oracle):
    """libsvgt_pack.so's wide -> compact converter and FORMAT formatter against their Python specifications
    (compact.compact_from_wide, genotype.ChunkWriter's Python fallback): identical bytes / text."""
    from svtyper_b200 import compact as cp, synth
    for cfg, n, m in (("mixed100k", 6000, 20), ("stress1m", 2500, 20), ("del1m4lib", 4000, 7)):
        b = synth.generate(cfg, n_sites=n, seed=3)
        a = cp.compact_from_wide(b, min_aligned=m)
        c = packer.compact_from_wide(b, min_aligned=m, threads=3)
        assert np.array_equal(a.sites, c.sites) and np.array_equal(a.rows, c.rows) and np.array_equal(a.order, c.order)
        assert c.min_aligned == m
    bad = ev.EvidenceBatch(b.sites.copy(), b.frags.copy(), b.splits.copy(), b.libs)
    bad.frags[:, 6] |= 600 << 16
    with pytest.raises(packer.PackError):
        packer.compact_from_wide(bad)
    rows = oracle.score(synth.generate("stress1m", n_sites=3000, seed=5))
    import io
    header = vcf.VcfHeader().parse([])
    header.ensure_svtyper_fields()
    header.add_sample("S")
    for classic_mode in (False, True):
        w = genotype.ChunkWriter(header, ["S"], classic_mode)
        fast, style = w._texts(rows)
        w.native = None
        slow, style2 = w._texts(rows)
        assert fast == slow and np.array_equal(style, style2)
    # host_sq: GT / GQ / SQ recomputed from GL with the host libm are the oracle's (same libm, same expressions),
    # whatever the scorer left there
    want = rows.copy()
    junk = rows.copy()
    called = junk["GT"] >= -1
    junk["SQ"][called] *= 1.0000001
    junk["GQ"][called] = 7
    junk["GT"][called] = 1
    packer.host_sq(junk, threads=2)
    assert junk.tobytes() == want.tobytes()
    q = np.array([0.0, 0.005, 2.675, 1743.0030805692013, 99999.995, 1e-9])
    assert packer.format_quals(q) == ["%0.2f" % v for v in q.tolist()]
