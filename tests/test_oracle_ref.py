"""CPU: the C oracle against the REFERENCE run live (oracle/_ref, the py3-patched copy).

oracle/_ref is generated from /root/reference by oracle/make_ref.py where that tree
exists (the build container); it travels to the GPU box with the snapshot.  Without it
these tests skip -- the committed golden fixtures (test_oracle_golden.py) still pin the
oracle.
"""
import numpy as np
import pytest

from oracle import ref_loader
from svtyper_b200 import evidence as ev, synth
from util import assert_rows_match

pytestmark = pytest.mark.skipif(not ref_loader.ensure(), reason="oracle/_ref not available")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


@pytest.mark.parametrize("config,n", [("del10k", 600), ("mixed100k", 800), ("del1m4lib", 600),
                                      ("stress1m", 500)])
def test_oracle_vs_live_reference(oracle, ref, config, n):
    from oracle import ref_adapter
    b = synth.generate(config, n_sites=n)
    exp = ref_adapter.reference_score(ref, b)
    got = oracle.score(b)
    assert_rows_match(got, exp, exact_gl=True, where=config)


def test_hazards_vs_live_reference(oracle, ref):
    from oracle import ref_adapter
    b = synth.hazard_batch()
    assert_rows_match(oracle.score(b), ref_adapter.reference_score(ref, b), exact_gl=True, where="hazard")


def test_reference_reproduces_golden_vcf(ref, tmp_path):
    """The patched reference + this repo's BAM reader reproduce the reference's golden VCF
    (reference tests/test_singlesample.py:20-44)."""
    import os
    from util import REPO
    data = os.path.join(REPO, "tests", "data")
    out = tmp_path / "out.vcf"
    with open(os.path.join(data, "example.vcf")) as fin, open(out, "w") as fout:
        ref.singlesample.sso_genotype(os.path.join(data, "NA12878.target_loci.sorted.bam"), fin, fout, 20, 1, 1,
                                      1000000, os.path.join(data, "NA12878.bam.json"), False, None, False,
                                      1000, 1e10, None, 1000)
    strip = lambda p: [l for l in open(p) if not l.startswith("##fileDate=")]
    assert strip(out) == strip(os.path.join(data, "example.gt.vcf"))
