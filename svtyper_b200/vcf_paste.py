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


def svt_join(master, sum_quals, vcf_list, out=None):
    """master: open text file or None (then the first VCF's path is re-opened); vcf_list: open text files."""
    out = sys.stdout if out is None else out
    if master is None:
        master = open_vcf(vcf_list[0].name)
    try:
        master_line = ""
        while True:                                         # header
            master_line = master.readline()
            if not master_line or master_line[:2] != "##":
                break
            out.write(master_line.rstrip() + "\n")
        out_v = master_line.rstrip().split("\t", MAX_SPLIT)[:9]
        for vcf in vcf_list:                                # sample names
            while True:
                line = vcf.readline()
                if not line:
                    break
                if line[:2] == "##":
                    continue
                if line[0] == "#":
                    out_v = out_v + line.rstrip().split("\t", MAX_SPLIT)[9:]
                    break
        out.write("\t".join(out_v) + "\n")
        lines = []
        while True:                                         # body
            master_line = master.readline()
            if not master_line:
                break
            out_v = master_line.rstrip().split("\t", MAX_SPLIT)[:8]
            qual = float(out_v[5])
            fmt = None
            for vcf in vcf_list:
                line = vcf.readline()
                if not line:
                    out.write("".join(lines))
                    sys.stderr.write("\nError: VCF files differ in length\n")
                    sys.exit(1)
                line_v = line.rstrip().split("\t", MAX_SPLIT)
                if fmt is None:
                    fmt = line_v[8]
                    out_v.append(fmt)
                qual += float(line_v[5])
                out_v = out_v + line_v[9:]
            if sum_quals:
                out_v[5] = py2_float_str(qual)
            lines.append("\t".join(out_v) + "\n")
            if len(lines) >= 4096:
                out.write("".join(lines))
                lines = []
        out.write("".join(lines))
    finally:
        master.close()
        for vcf in vcf_list:
            vcf.close()


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
