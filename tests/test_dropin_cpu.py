"""CPU: the drop-in entry points' host plumbing (VCF parse, read gathering, packing, VCF text)
on the reference's own fixture -- BASELINE.json configs[0].

There is no CPU scoring path in the product, so these tests install the parity ORACLE as the
scorer (tests may use it as the checker) and require the output VCF to be identical to the
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


@pytest.fixture()
def oracle_scorer(oracle):
    def scorer(batch, **params):
        return oracle.score(batch, **params)
    genotype.set_scorer(scorer)
    yield
    genotype.set_scorer(None)


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
    genotype.set_scorer(None)
    out = tmp_path / "x.vcf"
    with open(VCF) as fin, open(out, "w") as fout:
        with pytest.raises(Exception):
            singlesample.sso_genotype(BAM, fin, fout, 20, 1, 1, 1000000, LIB, False, None, False, 1000, 1e10,
                                      None, 1000)
