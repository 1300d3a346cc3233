"""GPU: the drop-in entry points end to end with the CUDA engine (no test scorer installed):
VCF + BAM in, VCF out, identical to the reference's golden file (reference
tests/test_svtyper.py:66-89, tests/test_singlesample.py:20-70)."""
import os

import pytest

from svtyper_b200 import classic, genotype, singlesample
from util import REPO

pytestmark = pytest.mark.gpu

DATA = os.path.join(REPO, "tests", "data")
BAM = os.path.join(DATA, "NA12878.target_loci.sorted.bam")
VCF = os.path.join(DATA, "example.vcf")
GOLD = os.path.join(DATA, "example.gt.vcf")
LIB = os.path.join(DATA, "NA12878.bam.json")


def _strip(path):
    return [l for l in open(path) if not l.startswith("##fileDate=")]


def test_classic_sv_genotype_on_gpu(tmp_path):
    out = tmp_path / "classic.vcf"
    with open(VCF) as fin, open(out, "w") as fout:
        classic.sv_genotype(BAM, fin, fout, 20, 1, 1, 1000000, LIB, False, None, None, False, None, 1e10)
    assert _strip(out) == _strip(GOLD)


@pytest.mark.parametrize("cores", [None, 1])
def test_sso_genotype_on_gpu(tmp_path, cores):
    out = tmp_path / "sso.vcf"
    with open(VCF) as fin, open(out, "w") as fout:
        singlesample.sso_genotype(BAM, fin, fout, 20, 1, 1, 1000000, LIB, False, None, False, 1000, 1e10,
                                  cores, 1000)
    assert _strip(out) == _strip(GOLD)
