"""CPU: the multi-sample classic path against the LIVE reference (oracle/_ref): `sv_genotype("a.bam,b.bam", ...)`
on two constructed BAMs -- one (site x sample) batch per chunk here, a per-sample loop there
(reference classic.py:279-284), QUAL summed across samples (:485) and reset by a sample without evidence
(:496-513); plus scripts/vcf_paste.py-style cohort merge of per-sample outputs (reference scripts/vcf_paste.py:41-117).
The scorer is the parity oracle (conftest.oracle_scorer), reading the merged COMPACT batch the product builds."""
import json
import os

import numpy as np
import pytest

from oracle import ref_loader
from svtyper_b200 import classic, compact as cp, evidence as ev, genotype, singlesample, synth

needs_ref = pytest.mark.skipif(not ref_loader.ensure(), reason="oracle/_ref not available")

REFS = [("chrA", 400000), ("chrB", 300000)]


def _write_bam(path, sample, seed, sites, covered):
    """Reads around the covered sites: FR pairs spanning a breakend, alt-orientation pairs across the junction,
    split reads with SA tags, lone reads; two libraries."""
    import bamwriter
    rng = np.random.default_rng(seed)
    header = ("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:chrA\tLN:400000\n@SQ\tSN:chrB\tLN:300000\n"
              "@RG\tID:%s_rg1\tSM:%s\tLB:libX\n@RG\tID:%s_rg2\tSM:%s\tLB:libY\n" % (sample, sample, sample, sample))
    recs = []
    for s, (posA, posB) in enumerate(sites):
        if s not in covered:
            continue
        for f in range(int(rng.integers(8, 40))):
            qname = "%s_%03d_%03d" % (sample, s, f)
            rg = "%s_rg%d" % (sample, 1 if rng.random() < 0.7 else 2)
            kind = rng.random()
            ins = int(rng.normal(300, 40))
            if kind < 0.45:                                   # reference-spanning FR pair at A or B
                anchor = posA if rng.random() < 0.5 else posB
                a = anchor - int(rng.integers(30, max(ins - 30, 31)))
                b = a + ins - 100
                pair = [(a, 0x1 | 0x20 | 0x40, "100M"), (b, 0x1 | 0x10 | 0x80, "100M")]
            elif kind < 0.75:                                 # pair across the deletion junction
                a = posA - int(rng.integers(100, 260))
                b = posB + int(rng.integers(0, 160))
                pair = [(a, 0x1 | 0x20 | 0x40, "100M"), (b, 0x1 | 0x10 | 0x80, "100M")]
            elif kind < 0.9:                                  # split read: clipped at A, supplementary part at B
                a = posA - 60
                pair = [(a, 0x1 | 0x20 | 0x40, "60M40S"), (a + 250, 0x1 | 0x10 | 0x80, "100M")]
            else:
                pair = [(posA - int(rng.integers(0, 90)), 0x1 | 0x40, "100M")]
            for k, (pos, flag, cigar) in enumerate(pair):
                tags = [("RG", "Z", rg)]
                if cigar == "60M40S":
                    tags.append(("SA", "Z", "chrA,%d,+,60S40M,60,0;" % (posB + 1)))
                recs.append(dict(tid=0, pos=max(0, pos), qname=qname, flag=flag, mapq=int(rng.choice([60, 60, 60, 37, 20, 0])),
                                 cigar=cigar, l_seq=100, tags=tags, tlen=ins if k == 0 else -ins))
    recs.sort(key=lambda r: (r["tid"], r["pos"]))
    bamwriter.write_bam(path, REFS, header, recs)


def _lib_json(path, samples):
    hist = {str(300 + d): int(1000 * np.exp(-0.5 * (d / 40.0) ** 2)) + 1 for d in range(-150, 151)}
    hist2 = {str(330 + d): int(800 * np.exp(-0.5 * (d / 55.0) ** 2)) + 1 for d in range(-200, 201)}
    doc = {}
    for s in samples:
        doc[s] = {"sample_name": s, "bam": s + ".bam", "mapped": 1000, "unmapped": 10, "libraryArray": [
            {"library_name": "libX", "readgroups": [s + "_rg1"], "read_length": 100, "mean": 300.0, "sd": 40.0,
             "prevalence": 0.7, "histogram": hist},
            {"library_name": "libY", "readgroups": [s + "_rg2"], "read_length": 100, "mean": 330.0, "sd": 55.0,
             "prevalence": 0.3, "histogram": hist2}]}
    with open(path, "w") as f:
        json.dump(doc, f)


def _vcf(path, sites, with_samples=()):
    head = ['##fileformat=VCFv4.2', '##INFO=<ID=SVTYPE,Number=1,Type=String,Description="Type of structural variant">',
            '##INFO=<ID=END,Number=1,Type=Integer,Description="End position of the variant described in this record">',
            '##INFO=<ID=CIPOS,Number=2,Type=Integer,Description="Confidence interval around POS">',
            '##INFO=<ID=CIEND,Number=2,Type=Integer,Description="Confidence interval around END">',
            '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
            "\t".join(["#CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO", "FORMAT"] + list(with_samples))]
    body = []
    for i, (posA, posB) in enumerate(sites):
        cols = ["chrA", str(posA), "sv%d" % i, "N", "<DEL>", "12.5" if i % 4 == 0 else ".", ".",
                "SVTYPE=DEL;END=%d;CIPOS=-3,4;CIEND=-2,2" % posB, "GT"] + ["./." for _ in with_samples]
        body.append("\t".join(cols))
    with open(path, "w") as f:
        f.write("\n".join(head + body) + "\n")


@pytest.fixture(scope="module")
def cohort(tmp_path_factory):
    d = tmp_path_factory.mktemp("cohort")
    rng = np.random.default_rng(7)
    sites = []
    for s in range(36):
        posA = 5000 + 9000 * s + int(rng.integers(0, 500))
        sites.append((posA, posA + int(rng.integers(400, 5000))))
    bams = []
    for k, name in enumerate(("S1", "S2", "S3")):
        covered = set(range(36)) if k == 0 else set(i for i in range(36) if (i + k) % 3 != 0)   # S2 / S3 miss a third
        path = str(d / (name + ".bam"))
        _write_bam(path, name, 100 + k, sites, covered)
        bams.append(path)
    lib = str(d / "libs.json")
    _lib_json(lib, ("S1", "S2", "S3"))
    vcf = str(d / "in.vcf")
    _vcf(vcf, sites)
    return dict(dir=d, sites=sites, bams=bams, lib=lib, vcf=vcf)


def _strip(text):
    return [l for l in text.split("\n") if not l.startswith("##fileDate=")]


@needs_ref
@pytest.mark.parametrize("sum_quals,max_reads,batch_size", [(False, None, None), (True, None, 5), (False, 30, 11)])
def test_three_sample_classic_matches_reference(cohort, oracle_scorer, tmp_path, sum_quals, max_reads, batch_size):
    ref = ref_loader.load()
    bam_string = ",".join(cohort["bams"])
    args = (20, 1, 1, 1000000, cohort["lib"], False, None, None, sum_quals, max_reads, 1e10)
    theirs = tmp_path / "ref.vcf"
    with open(cohort["vcf"]) as fin, open(theirs, "w") as fout:
        ref.classic.sv_genotype(bam_string, fin, fout, *args)
    mine = tmp_path / "mine.vcf"
    with open(cohort["vcf"]) as fin, open(mine, "w") as fout:
        classic.sv_genotype(bam_string, fin, fout, *args, batch_size=batch_size)
    got, want = _strip(open(mine).read()), _strip(open(theirs).read())
    assert got == want
    recs = [l.split("\t") for l in got if l and not l.startswith("#")]
    assert len(recs) == 36 and all(len(r) == 12 for r in recs)
    # the quirks this test exists for: a sample without evidence (blank row) next to called ones, QUAL summed over
    # the called samples and reset to 0 by a blank one that comes later
    assert any("./.:.:.:.:0" in r[10] or "./.:.:.:.:0" in r[11] for r in recs)
    assert any(r[5] == "0.00" and r[9].startswith(("0/", "1/")) for r in recs)
    if max_reads is not None:                                   # too many reads: classic writes GT ./. and nothing else
        assert any(c.startswith("./.:.:.:.:.:") for r in recs for c in r[9:])


def test_site_by_sample_batch_is_one_launch(cohort, oracle, monkeypatch, tmp_path):
    """The three samples of a chunk reach the scorer as ONE compact batch (3 x sites rows, joined library table),
    and its rows equal scoring each sample on its own."""
    seen = []

    def scorer(batch, **params):
        seen.append((batch.n_sites, batch.libs.n_lib))
        return oracle.score(cp.wide_from_compact(batch), **params)
    monkeypatch.setattr(genotype, "score", scorer)
    out = tmp_path / "o.vcf"
    with open(cohort["vcf"]) as fin, open(out, "w") as fout:
        classic.sv_genotype(",".join(cohort["bams"]), fin, fout, 20, 1, 1, 1000000, cohort["lib"], False, None, None, False,
                            None, 1e10)
    assert seen == [(3 * 36, 6)]
    merged = [l.split("\t") for l in open(out).read().split("\n") if l and not l.startswith("#")]
    for k, bam in enumerate(cohort["bams"]):
        single = tmp_path / ("s%d.vcf" % k)
        with open(cohort["vcf"]) as fin, open(single, "w") as fout:
            classic.sv_genotype(bam, fin, fout, 20, 1, 1, 1000000, cohort["lib"], False, None, None, False, None, 1e10)
        cols = [l.split("\t") for l in open(single).read().split("\n") if l and not l.startswith("#")]
        for m, s in zip(merged, cols):
            if s[8] == "GT":
                continue
            assert m[9 + k] == s[9], (k, m[2])


def test_merge_samples_rows_and_libraries(oracle):
    a = cp.compact_from_wide(synth.generate("mixed100k", n_sites=300, seed=1))
    b = cp.compact_from_wide(synth.generate("mixed100k", n_sites=300, seed=2, libs=synth.make_libraries(4)))
    m = genotype.merge_samples([a, b])
    assert m.n_sites == 600 and m.n_rows == a.n_rows + b.n_rows and m.libs.n_lib == a.libs.n_lib + b.libs.n_lib
    got = oracle.score(cp.wide_from_compact(m))
    assert got[:300].tobytes() == oracle.score(cp.wide_from_compact(a)).tobytes()
    assert got[300:].tobytes() == oracle.score(cp.wide_from_compact(b)).tobytes()
    assert sorted(m.order.tolist()) == list(range(600))
