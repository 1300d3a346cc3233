"""CPU: host-side pieces of the drop-in against the live reference (oracle/_ref):
library statistics measured from a BAM (reference parsers.py:501-583, statistics.py:40-121),
the -l JSON round trip, split-read QC helpers (reference tests/test_svtyper.py:14-56) and the
VCF edge cases of the entry points (unsupported SVTYPE pass-through, too-many-reads rows,
BND mates, sum_quals)."""
import io
import json
import os

import pytest

from oracle import ref_loader
from svtyper_b200 import classic, gather, genotype, sample, singlesample
from util import REPO

DATA = os.path.join(REPO, "tests", "data")
BAM = os.path.join(DATA, "NA12878.target_loci.sorted.bam")
VCF = os.path.join(DATA, "example.vcf")
LIB = os.path.join(DATA, "NA12878.bam.json")

needs_ref = pytest.mark.skipif(not ref_loader.ensure(), reason="oracle/_ref not available")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()



# ---- CIGAR helpers: the reference's own unit tests (tests/test_svtyper.py:14-56) -----------------
def test_cigar_helpers_match_reference_unit_tests():
    """Same vectors as the reference's TestCigarParsing (tests/test_svtyper.py:14-56)."""
    assert gather.cigar_from_string("5H3S2D1N5M3I2P2X1=") == [(5, 5), (4, 3), (2, 2), (3, 1), (0, 5), (1, 3),
                                                              (6, 2), (8, 2), (7, 1)]
    cig = gather.cigar_from_string("2S3M1D2M2I3M3S")
    assert gather.query_span(cig, True) == (3, 13, 15)
    assert gather.query_span(cig, False) == (2, 12, 15)
    assert gather.reference_end_of(1, gather.cigar_from_string("2S5M3D2M3S")) == 11
    cig = gather.cigar_from_string("2S5M3D1I1M3S")
    for rev in (True, False):
        assert gather.Piece(1, 25, None, rev, cig, 60).start_diagonal() == 23
    cig = gather.cigar_from_string("2S5M3D2I1M3S")
    for rev in (True, False):
        assert gather.Piece(1, 25, 34, rev, cig, 60).end_diagonal() == 34 - (2 + 8)


@needs_ref
def test_cigar_helpers_against_live_reference(ref):
    SR = ref.parsers.SplitRead
    for text in ("5H3S2D1M", "36M2D64M", "10S40M1I49M", "100M", "3S10M5N20M7H"):
        assert gather.cigar_from_string(text) == SR.cigarstring_to_tuple(text)
        cig = gather.cigar_from_string(text)
        assert gather.reference_end_of(1000, cig) == SR.get_reference_end_from_cigar(1000, cig)
        for rev in (False, True):
            qp = SR.SplitPiece.get_query_pos_from_cigar(list(cig), rev)
            assert gather.query_span(cig, rev) == (qp.query_start, qp.query_end, qp.query_length)
            theirs = SR.SplitPiece(1, 1000, rev, list(cig), 60)
            theirs.set_reference_end(gather.reference_end_of(1000, cig))
            mine = gather.Piece(1, 1000, gather.reference_end_of(1000, cig), rev, cig, 60)
            assert mine.start_diagonal() == SR.get_start_diagonal(theirs)
            assert mine.end_diagonal() == SR.get_end_diagonal(theirs)


# ---- library statistics from the BAM vs the reference -------------------------------------------
@needs_ref
def test_library_stats_from_bam_match_reference(ref):
    bam = sample.open_alignment(BAM)
    mine = sample.SampleInfo.from_bam(bam, 100000)
    rbam = ref.singlesample.open_alignment_file(BAM, None)
    theirs = ref.parsers.Sample.from_bam(rbam, 100000, 1e-3)
    assert mine.name == theirs.name
    assert [l.name for l in mine.libraries] == list(theirs.lib_dict.keys())
    for lib in mine.libraries:
        r = theirs.lib_dict[lib.name]
        assert lib.readgroups == r.readgroups
        assert lib.read_length == r.read_length
        assert lib.hist == dict(r.hist)
        assert lib.mean == r.mean
        assert lib.sd == pytest.approx(r.sd, rel=1e-12)      # summation order of a dict (see sample.py)
        assert lib.prevalence == r.prevalence
    assert mine.fetch_flank(3) == pytest.approx(theirs.get_fetch_flank(3), rel=1e-12)
    assert (mine.mapped, mine.unmapped) == (theirs.bam_mapped, theirs.bam_unmapped)


def test_library_json_round_trip(tmp_path):
    bam = sample.open_alignment(BAM)
    with open(LIB) as f:
        s = sample.SampleInfo.from_lib_info(bam, json.load(f))
    out = tmp_path / "lib.json"
    sample.write_sample_json([s], str(out))
    with open(out) as f:
        s2 = sample.SampleInfo.from_lib_info(sample.open_alignment(BAM), json.load(f))
    a, b = s.libraries[0], s2.libraries[0]
    assert (a.name, a.readgroups, a.read_length, a.mean, a.sd, a.prevalence, a.hist) == \
           (b.name, b.readgroups, b.read_length, b.mean, b.sd, b.prevalence, b.hist)
    assert (s.mapped, s.unmapped) == (s2.mapped, s2.unmapped)


# ---- read gathering + split candidates vs the reference ------------------------------------------
@needs_ref
def test_gathered_fragments_match_reference(ref):
    ss = ref.singlesample
    rs = ss.setup_sample(BAM, LIB, None, 1000000, 20)
    src = ss.init_vcf(VCF, rs, "/nonexistent")
    bps = ss.collect_breakpoints(src, 1e10)[:40]
    mine = sample.SampleInfo.open(BAM, LIB, None, 1000000)
    for bp in bps:
        regions = ss.get_breakpoint_regions(bp, rs, 3)
        rfr, _ = ss.gather_reads(rs.bam, bp["id"], regions, rs.rg_to_lib, rs.active_libs, 1000)
        mfr, over = gather.gather_sso(mine, bp, 3, 1000)
        assert not over and sorted(mfr) == sorted(rfr)
        for q in rfr:
            r, m = rfr[q], mfr[q]
            assert [(x.reference_start, x.flag) for x in r.primary_reads] == \
                   [(x.reference_start, x.flag) for x in m.primary_reads]
            assert len(r.split_reads) == len(m.split_reads)
            for a, b in zip(r.split_reads, m.split_reads):
                assert a.is_soft_clip == b.is_soft_clip
                for pa, pb in ((a.query_left, b.query_left), (a.query_right, b.query_right)):
                    assert (pa.chrom, pa.reference_start, pa.reference_end, pa.mapping_quality) == \
                           (pb.chrom, pb.reference_start, pb.reference_end, pb.mapping_quality)


# ---- entry-point edge cases, text-identical to the reference ------------------------------------
EDGE_VCF_EXTRA = (
    "1\t1000000\tX_ins\tN\t<INS>\t.\t.\tSVTYPE=INS;END=1000100;CIPOS=-1,1;CIEND=-1,1\tGT\t./.\n"
    "1\t1000500\tX_nosv\tN\t<DEL>\t.\t.\tEND=1000700\tGT\t./.\n")


def _edge_vcf(tmp_path):
    lines = open(VCF).read().split("\n")
    head = [l for l in lines if l.startswith("#")]
    body = [l for l in lines if l and not l.startswith("#")]
    path = tmp_path / "edge.vcf"
    path.write_text("\n".join(head + body[:25]) + "\n" + EDGE_VCF_EXTRA + "\n".join(body[25:60]) + "\n")
    return str(path)


def _strip(text):
    return [l for l in text.split("\n") if not l.startswith("##fileDate=")]


@needs_ref
@pytest.mark.parametrize("sum_quals,max_reads", [(False, 1000), (True, 1000), (False, 120)])
def test_sso_edge_cases_match_reference(ref, oracle_scorer, tmp_path, sum_quals, max_reads):
    path = _edge_vcf(tmp_path)
    args = (20, 1, 1, 1000000, LIB, False, None, sum_quals, max_reads, 1e10, None, 1000)
    theirs = tmp_path / "ref.vcf"
    with open(path) as fin, open(theirs, "w") as fout:
        ref.singlesample.sso_genotype(BAM, fin, fout, *args)
    mine = tmp_path / "mine.vcf"
    with open(path) as fin, open(mine, "w") as fout:
        singlesample.sso_genotype(BAM, fin, fout, *args)
    assert _strip(open(mine).read()) == _strip(open(theirs).read())


@needs_ref
@pytest.mark.parametrize("sum_quals,max_reads", [(False, None), (True, None), (False, 100)])
def test_classic_edge_cases_match_reference(ref, oracle_scorer, tmp_path, sum_quals, max_reads):
    path = _edge_vcf(tmp_path)
    args = (20, 1, 1, 1000000, LIB, False, None, None, sum_quals, max_reads, 1e10)
    theirs = tmp_path / "ref.vcf"
    with open(path) as fin, open(theirs, "w") as fout:
        ref.classic.sv_genotype(BAM, fin, fout, *args)
    mine = tmp_path / "mine.vcf"
    with open(path) as fin, open(mine, "w") as fout:
        classic.sv_genotype(BAM, fin, fout, *args)
    assert _strip(open(mine).read()) == _strip(open(theirs).read())


def test_bad_alignment_name_exits():
    with pytest.raises(SystemExit):
        classic.sv_genotype("reads.sam", io.StringIO(""), io.StringIO(), 20, 1, 1, 1000, None, False, None, None,
                            False, None, 1e10)


@needs_ref
def test_sso_records_with_existing_format_values_match_reference(ref, oracle_scorer, tmp_path):
    """Records that already carry FORMAT values for the sample (and a FORMAT key svtyper does not write)
    cannot take the direct text path (genotype.RowFormatter): the generic record model must give the
    reference's text for them, next to fast-path records in the same file."""
    lines = open(VCF).read().split("\n")
    head = [l for l in lines if l.startswith("##")]
    head.append('##FORMAT=<ID=XX,Number=1,Type=Integer,Description="pre-existing per-sample value">')
    chrom_line = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tNA12878"
    body = [l for l in lines if l and not l.startswith("#")][:40]
    out = []
    for i, l in enumerate(body):
        c = l.split("\t")[:8]
        if i % 3 == 0:
            c += ["GT:XX", "0/1:%d" % (i + 1)]
        elif i % 3 == 1:
            c += ["GT", "./."]
        out.append("\t".join(c))
    path = tmp_path / "fmt.vcf"
    path.write_text("\n".join(head + [chrom_line] + out) + "\n")
    args = (20, 1, 1, 1000000, LIB, False, None, False, 1000, 1e10, None, 1000)
    theirs = tmp_path / "ref.vcf"
    with open(path) as fin, open(theirs, "w") as fout:
        ref.singlesample.sso_genotype(BAM, fin, fout, *args)
    mine = tmp_path / "mine.vcf"
    with open(path) as fin, open(mine, "w") as fout:
        singlesample.sso_genotype(BAM, fin, fout, *args)
    got, want = _strip(open(mine).read()), _strip(open(theirs).read())
    assert got == want
    assert any(":XX" in l.split("\t")[8] for l in got if not l.startswith("#") and l)


@needs_ref
def test_duplicate_library_name_in_json_matches_reference(ref, oracle_scorer, tmp_path):
    """A -l JSON that lists one library name twice: the reference's name-keyed dict keeps the later entry, the
    earlier entry's read groups keep pointing at the replaced object and their reads are skipped
    (parsers.py:636-644).  Same text here."""
    doc = json.load(open(LIB))
    name = list(doc.keys())[0]
    first = doc[name]["libraryArray"][0]
    second = dict(first)
    second["readgroups"] = []
    second["mean"], second["sd"] = first["mean"] + 40.0, first["sd"] + 5.0
    doc[name]["libraryArray"] = [first, second]
    lib = tmp_path / "dup.json"
    lib.write_text(json.dumps(doc))
    lines = open(VCF).read().split("\n")
    path = tmp_path / "few.vcf"
    path.write_text("\n".join([l for l in lines if l.startswith("#")] + [l for l in lines if l and not l.startswith("#")][:30]) + "\n")
    args = (20, 1, 1, 1000000, str(lib), False, None, False, 1000, 1e10, None, 1000)
    theirs, mine = tmp_path / "ref.vcf", tmp_path / "mine.vcf"
    with open(path) as fin, open(theirs, "w") as fout:
        ref.singlesample.sso_genotype(BAM, fin, fout, *args)
    with open(path) as fin, open(mine, "w") as fout:
        singlesample.sso_genotype(BAM, fin, fout, *args)
    assert _strip(open(mine).read()) == _strip(open(theirs).read())
