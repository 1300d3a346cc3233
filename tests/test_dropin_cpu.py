"""CPU: the drop-in entry points' host plumbing (VCF parse, read gathering, packing, VCF text)
on the reference's own fixture -- BASELINE.json configs[0].

There is no CPU scoring path in the product, so these tests monkeypatch the parity ORACLE over
genotype.score (conftest.oracle_scorer) and require the output VCF to be identical to the
reference's golden file, `diff -I '^##fileDate='`-clean (reference tests/test_svtyper.py:66-89,
tests/test_singlesample.py:20-70).
"""
import os

import pytest

from svtyper_b200 import classic, genotype, singlesample
from util import REPO

DATA = os.path.join(REPO, "tests", "data")
BAM = os.path.join(DATA, "NA12878.target_loci.sorted.bam")
VCF = os.path.join(DATA, "example.vcf")
GOLD = os.path.join(DATA, "example.gt.vcf")
LIB = os.path.join(DATA, "NA12878.bam.json")


def _strip(path):
    return [l for l in open(path) if not l.startswith("##fileDate=")]


def test_classic_reproduces_golden_vcf(oracle_scorer, tmp_path):
    out = tmp_path / "classic.vcf"
    with open(VCF) as fin, open(out, "w") as fout:
        classic.sv_genotype(BAM, fin, fout, 20, 1, 1, 1000000, LIB, False, None, None, False, None, 1e10)
    assert _strip(out) == _strip(GOLD)


@pytest.mark.parametrize("cores", [None, 1])
def test_sso_reproduces_golden_vcf(oracle_scorer, tmp_path, cores):
    out = tmp_path / "sso.vcf"
    with open(VCF) as fin, open(out, "w") as fout:
        singlesample.sso_genotype(BAM, fin, fout, 20, 1, 1, 1000000, LIB, False, None, False, 1000, 1e10,
                                  cores, 1000)
    assert _strip(out) == _strip(GOLD)


def test_product_path_refuses_to_run_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    out = tmp_path / "x.vcf"
    with open(VCF) as fin, open(out, "w") as fout:
        with pytest.raises(Exception):
            singlesample.sso_genotype(BAM, fin, fout, 20, 1, 1, 1000000, LIB, False, None, False, 1000, 1e10,
                                      None, 1000)


@pytest.mark.parametrize("batch_size", [1, 7, 64, 100000])
def test_sso_batch_size_streams_the_same_vcf(oracle_scorer, tmp_path, batch_size):
    """batch_size = breakpoints per pack/score/write chunk (reference singlesample.py:723-725); BND mates that fall
    into different chunks, pass-through records and chunk boundaries must not change a byte."""
    out = tmp_path / "sso.vcf"
    with open(VCF) as fin, open(out, "w") as fout:
        singlesample.sso_genotype(BAM, fin, fout, 20, 1, 1, 1000000, LIB, False, None, False, 1000, 1e10, 2, batch_size)
    assert _strip(out) == _strip(GOLD)


def test_classic_chunked_and_cli(oracle_scorer, tmp_path):
    out = tmp_path / "classic.vcf"
    with open(VCF) as fin, open(out, "w") as fout:
        classic.sv_genotype(BAM, fin, fout, 20, 1, 1, 1000000, LIB, False, None, None, False, None, 1e10, batch_size=13)
    assert _strip(out) == _strip(GOLD)
    out2 = tmp_path / "cli.vcf"
    classic.main(["-B", BAM, "-i", VCF, "-o", str(out2), "-l", LIB])
    assert _strip(out2) == _strip(GOLD)
    out3 = tmp_path / "cli_sso.vcf"
    singlesample.main(["-B", BAM, "-i", VCF, "-o", str(out3), "-l", LIB, "--batch_size", "50", "--cores", "2"])
    assert _strip(out3) == _strip(GOLD)


def test_cli_defaults_match_the_reference():
    a = singlesample.get_args(["-B", "x.bam"])
    assert (a.min_aligned, a.num_samp, a.max_reads, a.max_ci_dist, a.split_weight, a.disc_weight, a.cores, a.batch_size,
            a.sum_quals) == (20, 1000000, 1000, 1e10, 1, 1, None, 1000, False)
    c = classic.get_args(["-B", "x.bam,y.bam"])
    assert (c.min_aligned, c.num_samp, c.max_reads, c.max_ci_dist, c.alignment_outpath, c.verbose) == (20, 1000000, None, 1e10, None, False)
