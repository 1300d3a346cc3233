"""Cohort merge of per-sample genotyped VCFs: paste the sample columns of several VCFs side by side.

Host-side companion of the single-sample entry point (`svtyper-sso` writes one VCF per sample; the reference joins
them with scripts/vcf_paste.py).  Same observable behaviour as reference scripts/vcf_paste.py:41-117 `svt_join`:

  * the `##` header lines and the first eight columns of every record come from the master VCF (default: the first
    VCF of the list, re-opened -- :45-53, :84-86);
  * the `#CHROM` line is the master's first nine columns plus every input's sample names (:66-78);
  * FORMAT is taken from the FIRST input of each record, not from the master (:99-104);
  * with `sum_quals` the QUAL column is the master's QUAL PLUS every input's QUAL (:87, :106, :108-109) -- so with
    the default master the first VCF counts twice, as in the reference -- printed the way Python 2 prints a float
    (`str(float)`: 12 significant digits, always a `.0` or an exponent);
  * inputs shorter than the master end the run with the reference's message and exit status 1 (:93-95).

Text only: no parsing of the sample columns, one `str.split` per line; `.gz` inputs are read through gzip.
"""
from __future__ import annotations

import gzip
import sys

MAX_SPLIT = 9


def py2_float_str(x):
    """Python 2's str(float): repr with 12 significant digits (`'%.12g'` plus a trailing `.0` for integral values)."""
    s = "%.12g" % x
    if "." not in s and "e" not in s and "n" not in s:      # 'inf' / 'nan' contain an n
        s += ".0"
    return s


def open_vcf(path):
    return gzip.open(path, "rt") if path.endswith(".gz") else open(path, "r")


def _header(handle, echo=None):
    """The fields of the column-header line.  Master (`echo` given): its `##` lines are echoed and the first other
    line is taken, whatever it is (:57-66).  Inputs: everything up to the first line that starts with a single `#` is
    skipped, `[]` at end of file (:68-78)."""
    for text in handle:
        if text.startswith("##"):
            if echo is not None:
                echo.write(text.rstrip() + "\n")
        elif echo is not None or text.startswith("#"):
            return text.rstrip().split("\t", MAX_SPLIT)
    return []


def svt_join(master, sum_quals, vcf_list, out=None):
    """master: open text file or None (then the first VCF's path is re-opened); vcf_list: open text files."""
    out = sys.stdout if out is None else out
    master = open_vcf(vcf_list[0].name) if master is None else master
    inputs = list(vcf_list)
    try:
        names = _header(master, echo=out)[:9]               # the master keeps only its first nine columns
        for handle in inputs:
            names += _header(handle)[9:]
        out.write("\t".join(names) + "\n")
        pending = []
        for record in master:
            fixed = record.rstrip().split("\t", MAX_SPLIT)[:8]
            total = float(fixed[5])                          # the master's own QUAL is part of the sum
            samples, fmt = [], None
            for handle in inputs:
                text = handle.readline()
                if not text:
                    out.write("".join(pending))
                    sys.stderr.write("\nError: VCF files differ in length\n")
                    sys.exit(1)
                cols = text.rstrip().split("\t", MAX_SPLIT)
                fmt = cols[8] if fmt is None else fmt        # FORMAT of the first input, not of the master
                total += float(cols[5])
                samples += cols[9:]
            if sum_quals:
                fixed[5] = py2_float_str(total)
            pending.append("\t".join(fixed + ([fmt] if fmt is not None else []) + samples) + "\n")
            if len(pending) >= 4096:
                out.write("".join(pending))
                pending = []
        out.write("".join(pending))
    finally:
        master.close()
        for handle in inputs:
            handle.close()


def get_args(argv=None):
    import argparse
    parser = argparse.ArgumentParser(prog="svtyper-paste", description="Paste VCFs from multiple samples")
    parser.add_argument("-m", "--master", type=argparse.FileType("r"), default=None,
                        help="VCF file to set first 8 columns of variant info [first file in vcf_list]")
    parser.add_argument("-q", "--sum_quals", required=False, action="store_true",
                        help="Sum QUAL scores of input VCFs as output QUAL score")
    parser.add_argument("-f", "--vcf_list", required=True, help="Line-delimited list of VCF files to paste")
    return parser.parse_args(argv)


def main(argv=None):
    args = get_args(argv)
    with open(args.vcf_list, "r") as f:
        vcf_list = [open_vcf(line.rstrip()) for line in f]
    svt_join(args.master, args.sum_quals, vcf_list)


def cli():
    try:
        sys.exit(main())
    except IOError as e:
        if e.errno != 32:       # ignore SIGPIPE, as the reference does
            raise


if __name__ == "__main__":
    cli()
