"""Drop-in for `svtyper.singlesample.sso_genotype` (reference svtyper/singlesample.py:764-816) and the
`svtyper-sso` console entry point (reference singlesample.py:18-50, :818-853).

Same signature and VCF contract.  The reference's serial and `mp.Pool` modes differ only in who gathers and
scores which breakpoint; both produce the same VCF.  Here `batch_size` is the number of breakpoints per
pack -> score -> write chunk (singlesample.py:723-725 `partition_all(batch_size, ...)`: chunks are pipelined, so
host memory is bounded by two chunks) and `cores` the number of host threads of the native read gathering
(singlesample.py:746 `mp.Pool(cores)`; None = serial in the reference, one thread per hardware thread here --
the output does not depend on it).  Every chunk is scored by the CUDA engine.
"""
from __future__ import annotations

import argparse
import os
import sys

from . import gather, genotype, packer, vcf, version
from .sample import SampleInfo, write_sample_json


def _open_again(vcf_in):
    """The reference re-opens the VCF by name for each of its passes (singlesample.py:581,587); one streaming
    pass is enough here, from a fresh handle when the input is a real file, else from the handle itself."""
    path = os.path.abspath(getattr(vcf_in, "name", "<stdin>") or "<stdin>")
    if os.path.basename(path) == "<stdin>" or not os.path.isfile(path):
        return vcf_in, False
    return open(path, "r"), True


def sso_genotype(bam_string,
                 vcf_in,
                 vcf_out,
                 min_aligned,
                 split_weight,
                 disc_weight,
                 num_samp,
                 lib_info_path,
                 debug,
                 ref_fasta,
                 sum_quals,
                 max_reads,
                 max_ci_dist,
                 cores,
                 batch_size):
    if vcf_in is None:
        return
    bam_path = os.path.abspath(bam_string)
    if not (bam_path.endswith(".bam") or bam_path.endswith(".cram")):
        sys.exit("Error: %s is not a valid alignment file (*.bam or *.cram)\n" % bam_path)
    sample = SampleInfo.open(bam_path, lib_info_path, ref_fasta, num_samp)
    if lib_info_path is not None and not os.path.exists(lib_info_path):
        write_sample_json([sample], lib_info_path)

    stream, owned = _open_again(vcf_in)
    try:
        # the reference reads only the '##' lines into the header and then appends the BAM's
        # sample: the output has exactly one sample column (singlesample.py:112-124)
        meta, first = [], None
        for line in stream:
            if line.startswith("##"):
                meta.append(line)
            else:
                first = line
                break
        header = vcf.VcfHeader().parse(meta)
        header.ensure_svtyper_fields()
        header.add_sample(sample.name)
        vcf_out.write(header.render() + "\n")

        def body():
            if first is not None:
                yield first
            for line in stream:
                yield line

        chunk = int(batch_size) if batch_size else genotype.DEFAULT_BATCH
        threads = int(cores) if cores else 0
        plans = genotype.walk_records(body(), header, sum_quals, max_ci_dist, max(chunk, 1), {})
        genotype.run_pipeline(
            [sample], plans, lambda lines: vcf_out.write("\n".join(lines) + "\n") if lines else None,
            lambda smp, bp: gather.gather_sso(smp, bp, genotype.Z, max_reads), packer.MODE_SSO, False,
            min_aligned, split_weight, disc_weight, max_reads, header, threads=threads)
    finally:
        if owned:
            stream.close()
        sample.close()


# --------------------------------------------------------------------------------------------
# command line (reference singlesample.py:18-50, :818-853)
def get_args(argv=None):
    parser = argparse.ArgumentParser(formatter_class=argparse.RawTextHelpFormatter, description="\
svtyper\n\
author: " + version.__author__ + "\n\
version: " + version.__version__ + "\n\
description: Compute genotype of structural variants based on breakpoint depth on a SINGLE sample")
    parser.add_argument('-i', '--input_vcf', metavar='FILE', type=argparse.FileType('r'), default=None, help='VCF input (default: stdin)')
    parser.add_argument('-o', '--output_vcf', metavar='FILE', type=argparse.FileType('w'), default=sys.stdout, help='output VCF to write (default: stdout)')
    parser.add_argument('-B', '--bam', metavar='FILE', type=str, required=True, help='BAM or CRAM file(s), comma-separated if genotyping multiple samples')
    parser.add_argument('-T', '--ref_fasta', metavar='FILE', type=str, required=False, default=None, help='Indexed reference FASTA file (recommended for reading CRAM files)')
    parser.add_argument('-S', '--split_bam', type=str, required=False, help=argparse.SUPPRESS)
    parser.add_argument('-l', '--lib_info', metavar='FILE', dest='lib_info_path', type=str, required=False, default=None, help='create/read JSON file of library information')
    parser.add_argument('-m', '--min_aligned', metavar='INT', type=int, required=False, default=20, help='minimum number of aligned bases to consider read as evidence [20]')
    parser.add_argument('-n', dest='num_samp', metavar='INT', type=int, required=False, default=1000000, help='number of reads to sample from BAM file for building insert size distribution [1000000]')
    parser.add_argument('-q', '--sum_quals', action='store_true', required=False, help='add genotyping quality to existing QUAL (default: overwrite QUAL field)')
    parser.add_argument('--max_reads', metavar='INT', type=int, default=1000, required=False, help='maximum number of reads to assess at any variant (reduces processing time in high-depth regions, default: 1000)')
    parser.add_argument('--max_ci_dist', metavar='INT', type=int, default=1e10, required=False, help='maximum size of a confidence interval before 95%% CI is used intead (default: 1e10)')
    parser.add_argument('--split_weight', metavar='FLOAT', type=float, required=False, default=1, help='weight for split reads [1]')
    parser.add_argument('--disc_weight', metavar='FLOAT', type=float, required=False, default=1, help='weight for discordant paired-end reads [1]')
    parser.add_argument('--debug', action='store_true', help=argparse.SUPPRESS)
    parser.add_argument('--cores', type=int, metavar='INT', required=False, default=None, help='number of host threads gathering reads (default: one per hardware thread)')
    parser.add_argument('--batch_size', type=int, metavar='INT', required=False, default=1000, help='number of breakpoints per pack/score/write chunk')
    args = parser.parse_args(argv)
    # if no input, check if part of pipe and if so, read stdin.
    if args.input_vcf is None:
        if not sys.stdin.isatty():
            args.input_vcf = sys.stdin
    return args


def main(argv=None):
    args = get_args(argv)
    if args.split_bam is not None:
        sys.stderr.write('Warning: --split_bam (-S) is deprecated. Ignoring %s.\n' % args.split_bam)
    sso_genotype(args.bam, args.input_vcf, args.output_vcf, args.min_aligned, args.split_weight, args.disc_weight,
                 args.num_samp, args.lib_info_path, args.debug, args.ref_fasta, args.sum_quals, args.max_reads,
                 args.max_ci_dist, args.cores, args.batch_size)


def cli():
    try:
        sys.exit(main())
    except IOError as e:
        if e.errno != 32:  # ignore SIGPIPE
            raise


if __name__ == '__main__':
    cli()
