"""CPU: svtyper_b200.vcf_paste (cohort merge of per-sample VCFs) against reference scripts/vcf_paste.py.

The reference script is Python 2 (`except IOError, e`); where /root/reference exists it is run live from a temporary
copy with that one line patched.  Python 3 prints floats differently from Python 2's str(float), so the live leg uses
QUAL values both print alike and `py2_float_str` is pinned separately on known Python 2 outputs."""
import io
import os
import subprocess
import sys

import pytest

from svtyper_b200 import vcf_paste

REF_SCRIPT = "/root/reference/scripts/vcf_paste.py"

HEADER = "##fileformat=VCFv4.2\n##source=%s\n"
COLS = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s\n"


def write_vcf(path, source, samples, records):
    with open(path, "w") as f:
        f.write(HEADER % source)
        f.write(COLS % "\t".join(samples))
        for r in records:
            f.write("\t".join(r) + "\n")
    return str(path)


@pytest.fixture
def three(tmp_path):
    a = write_vcf(tmp_path / "a.vcf", "A", ["s1"], [
        ["1", "100", "v1", "N", "<DEL>", "12.5", ".", "SVTYPE=DEL;END=900", "GT:GQ", "0/1:40"],
        ["2", "200", "v2", "N", "<DUP>", "0.25", ".", "SVTYPE=DUP;END=999", "GT:GQ", "0/0:7"]])
    b = write_vcf(tmp_path / "b.vcf", "B", ["s2", "s3"], [
        ["1", "100", "v1", "N", "<DEL>", "100.0", "PASS", "X", "GT", "1/1", "./."],
        ["2", "200", "v2", "N", "<DUP>", "3.5", "PASS", "X", "GT", "0/1", "0/0"]])
    c = write_vcf(tmp_path / "c.vcf", "C", ["s4"], [
        ["1", "100", "v1", "N", "<DEL>", "0.0", ".", "Y", "GT:GQ:SQ", "0/0:1:0.00"],
        ["2", "200", "v2", "N", "<DUP>", "1.75", ".", "Y", "GT:GQ:SQ", "0/1:2:1.75"]])
    return a, b, c


def run_ours(paths, master=None, sum_quals=False):
    out = io.StringIO()
    vcfs = [vcf_paste.open_vcf(p) for p in paths]
    vcf_paste.svt_join(open(master) if master else None, sum_quals, vcfs, out)
    return out.getvalue()


def test_paste_semantics(three):
    a, b, c = three
    got = run_ours([a, b, c]).splitlines()
    assert got[:2] == ["##fileformat=VCFv4.2", "##source=A"]                        # header of the master (first VCF)
    assert got[2].split("\t")[9:] == ["s1", "s2", "s3", "s4"]
    r = got[3].split("\t")
    assert r[:8] == ["1", "100", "v1", "N", "<DEL>", "12.5", ".", "SVTYPE=DEL;END=900"]   # master's first 8 columns
    assert r[8] == "GT:GQ" and r[9:] == ["0/1:40", "1/1", "./.", "0/0:1:0.00"]      # FORMAT of the first input
    q = run_ours([a, b, c], sum_quals=True).splitlines()
    assert q[3].split("\t")[5] == "125.0"                   # 12.5 (master = a) + 12.5 + 100.0 + 0.0: a counts twice
    assert q[4].split("\t")[5] == "5.75"
    m = run_ours([a, c], master=b, sum_quals=True).splitlines()
    assert m[1] == "##source=B" and m[2].split("\t")[9:] == ["s1", "s4"]        # only the master's first nine columns
    assert m[3].split("\t")[5] == "112.5" and m[3].split("\t")[6] == "PASS"


def test_py2_float_str():
    f = vcf_paste.py2_float_str
    assert f(5.0) == "5.0" and f(0.1 + 0.2) == "0.3" and f(1 / 3.0) == "0.333333333333"
    assert f(1e16) == "1e+16" and f(123456789012.345) == "123456789012.0" and f(1234567890123.0) == "1.23456789012e+12"
    assert f(0.0) == "0.0" and f(float("inf")) == "inf" and f(2.5e-7) == "2.5e-07"


def test_shorter_input_is_an_error(three, tmp_path, capsys):
    a, b, _ = three
    short = write_vcf(tmp_path / "short.vcf", "S", ["s9"], [
        ["1", "100", "v1", "N", "<DEL>", "1.0", ".", "Z", "GT", "0/0"]])
    out = io.StringIO()
    with pytest.raises(SystemExit) as ei:
        vcf_paste.svt_join(None, False, [open(a), open(short)], out)
    assert ei.value.code == 1
    assert "VCF files differ in length" in capsys.readouterr().err
    assert len(out.getvalue().splitlines()) == 4            # header x2, #CHROM, the one complete record


def test_gz_inputs_and_cli(three, tmp_path):
    import gzip
    a, b, c = three
    gz = str(tmp_path / "b.vcf.gz")
    with open(b, "rb") as fi, gzip.open(gz, "wb") as fo:
        fo.write(fi.read())
    lst = tmp_path / "list.txt"
    lst.write_text("\n".join([a, gz, c]) + "\n")
    res = subprocess.run([sys.executable, "-m", "svtyper_b200.vcf_paste", "-f", str(lst), "-q"], capture_output=True,
                         text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert res.returncode == 0, res.stderr
    assert res.stdout == run_ours([a, b, c], sum_quals=True)


@pytest.mark.skipif(not os.path.exists(REF_SCRIPT), reason="reference tree not available")
@pytest.mark.parametrize("sum_quals,with_master", [(False, False), (True, False), (True, True)])
def test_against_live_reference_script(three, tmp_path, sum_quals, with_master):
    a, b, c = three
    src = open(REF_SCRIPT).read().replace("except IOError, e:", "except IOError as e:")
    script = tmp_path / "ref_vcf_paste.py"
    script.write_text(src)
    lst = tmp_path / "list.txt"
    lst.write_text("\n".join([a, b, c]) + "\n")
    cmd = [sys.executable, str(script), "-f", str(lst)] + (["-q"] if sum_quals else []) + (["-m", b] if with_master else [])
    ref = subprocess.run(cmd, capture_output=True, text=True)
    assert ref.returncode == 0, ref.stderr
    assert run_ours([a, b, c], master=b if with_master else None, sum_quals=sum_quals) == ref.stdout
