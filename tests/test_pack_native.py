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
    assert L.svgt_pack_abi_version() == 1
    header = open(os.path.join(REPO, "include", "svgt_pack.h")).read()
    names = set(re.findall(r"\b(svgt_(?:pack|bam)_[a-z_]+)\s*\(", header))
    assert {"svgt_bam_open", "svgt_bam_close", "svgt_bam_count", "svgt_pack_sites", "svgt_pack_rows"} <= names
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
