"""Drop-in for `svtyper.classic.sv_genotype` (reference svtyper/classic.py:107-533) and the `svtyper` console
entry point (reference classic.py:14-50, :534-583).

Same positional/keyword signature, same VCF-in / VCF-out text contract, one or several comma-separated BAMs.
The records are walked in chunks; per chunk every sample's evidence is packed, the (site x sample) batch is
scored by the CUDA engine in ONE launch pair and the FORMAT text of all its rows is produced at once (see
genotype.py).  QUAL sums the samples' SQ in order and a sample without evidence resets it to 0, exactly as the
reference's per-sample loop does (classic.py:279-284, :485, :496-513).

`alignment_outpath` (-w, the diagnostic BAM of supporting reads, classic.py:161-166) is not implemented: the
evidence rows carry no read sequences to write back; it is refused rather than silently ignored.
"""
from __future__ import annotations

import argparse
import logging
import os
import sys

from . import gather, genotype, packer, vcf, version
from .sample import SampleInfo, write_sample_json


def sv_genotype(bam_string,
                vcf_in,
                vcf_out,
                min_aligned,
                split_weight,
                disc_weight,
                num_samp,
                lib_info_path,
                debug,
                alignment_outpath,
                ref_fasta,
                sum_quals,
                max_reads,
                max_ci_dist,
                batch_size=None,
                cores=None):
    for path in bam_string.split(","):
        if not (path.endswith(".bam") or path.endswith(".cram")):
            sys.stderr.write("Error: %s is not a valid alignment file (*.bam or *.cram)\n" % path)
            sys.exit(1)
    if alignment_outpath is not None:
        raise NotImplementedError("alignment_outpath (-w evidence BAM dump) is not part of the accelerated path")
    if vcf_in is None:
        sys.stderr.write("Warning: VCF not found.\n")
    samples = [SampleInfo.open(path, lib_info_path, ref_fasta, num_samp) for path in bam_string.split(",")]
    if lib_info_path is not None and not os.path.isfile(lib_info_path):
        logging.info("Writing library metrics to %s..." % lib_info_path)
        write_sample_json(samples, lib_info_path)
    if vcf_in is None:
        return

    header_lines, first = [], None
    for line in vcf_in:
        if line.startswith("#"):
            header_lines.append(line)
        elif line.strip():
            first = line
            break
    if first is None:                 # the reference emits its header with the first record
        vcf_in.close()
        vcf_out.close()
        return
    header = vcf.VcfHeader().parse(header_lines)
    header.ensure_svtyper_fields()
    for s in samples:
        if s.name not in header.samples:
            header.add_sample(s.name)
    vcf_out.write(header.render() + "\n")

    def body():
        yield first
        for line in vcf_in:
            yield line

    open_bnds = {}
    chunk = int(batch_size) if batch_size else 16 * genotype.DEFAULT_BATCH
    plans = genotype.walk_records(body(), header, sum_quals, max_ci_dist, max(chunk, 1), open_bnds)
    genotype.run_pipeline(
        samples, plans, lambda lines: vcf_out.write("\n".join(lines) + "\n") if lines else None,
        lambda smp, bp: gather.gather_classic(smp, bp, genotype.Z, max_reads), packer.MODE_CLASSIC, True,
        min_aligned, split_weight, disc_weight, max_reads, header, threads=int(cores) if cores else 0)
    if open_bnds:
        logging.warning("Unpaired breakends found in file. These will not be present in output.")
    vcf_in.close()
    vcf_out.close()
    for s in samples:
        s.close()
    return


# --------------------------------------------------------------------------------------------
# command line (reference classic.py:14-50, :534-583)
def get_args(argv=None):
    parser = argparse.ArgumentParser(formatter_class=argparse.RawTextHelpFormatter, description="\
svtyper\n\
author: " + version.__author__ + "\n\
version: " + version.__version__ + "\n\
description: Compute genotype of structural variants based on breakpoint depth")
    parser.add_argument('-i', '--input_vcf', metavar='FILE', type=argparse.FileType('r'), default=None, help='VCF input (default: stdin)')
    parser.add_argument('-o', '--output_vcf', metavar='FILE', type=argparse.FileType('w'), default=sys.stdout, help='output VCF to write (default: stdout)')
    parser.add_argument('-B', '--bam', metavar='FILE', type=str, required=True, help='BAM or CRAM file(s), comma-separated if genotyping multiple samples')
    parser.add_argument('-T', '--ref_fasta', metavar='FILE', type=str, required=False, default=None, help='Indexed reference FASTA file (recommended for reading CRAM files)')
    parser.add_argument('-S', '--split_bam', type=str, required=False, help=argparse.SUPPRESS)
    parser.add_argument('-l', '--lib_info', metavar='FILE', dest='lib_info_path', type=str, required=False, default=None, help='create/read JSON file of library information')
    parser.add_argument('-m', '--min_aligned', metavar='INT', type=int, required=False, default=20, help='minimum number of aligned bases to consider read as evidence [20]')
    parser.add_argument('-n', dest='num_samp', metavar='INT', type=int, required=False, default=1000000, help='number of reads to sample from BAM file for building insert size distribution [1000000]')
    parser.add_argument('-q', '--sum_quals', action='store_true', required=False, help='add genotyping quality to existing QUAL (default: overwrite QUAL field)')
    parser.add_argument('--max_reads', metavar='INT', type=int, default=None, required=False, help='maximum number of reads to assess at any variant (reduces processing time in high-depth regions, default: unlimited)')
    parser.add_argument('--max_ci_dist', metavar='INT', type=int, default=1e10, required=False, help='maximum size of a confidence interval before 95%% CI is used intead (default: 1e10)')
    parser.add_argument('--split_weight', metavar='FLOAT', type=float, required=False, default=1, help='weight for split reads [1]')
    parser.add_argument('--disc_weight', metavar='FLOAT', type=float, required=False, default=1, help='weight for discordant paired-end reads [1]')
    parser.add_argument('-w', '--write_alignment', metavar='FILE', dest='alignment_outpath', type=str, required=False, default=None, help='write relevant reads to BAM file (not implemented in this build)')
    parser.add_argument('--debug', action='store_true', help=argparse.SUPPRESS)
    parser.add_argument('--verbose', action='store_true', default=False, help='Report status updates')
    args = parser.parse_args(argv)
    # if no input, check if part of pipe and if so, read stdin.
    if args.input_vcf is None:
        if not sys.stdin.isatty():
            args.input_vcf = sys.stdin
    return args


def set_up_logging(verbose):
    logging.basicConfig(format='%(message)s', level=logging.INFO if verbose else logging.WARNING)


def main(argv=None):
    args = get_args(argv)
    set_up_logging(args.verbose)
    if args.split_bam is not None:
        sys.stderr.write('Warning: --split_bam (-S) is deprecated. Ignoring %s.\n' % args.split_bam)
    sv_genotype(args.bam, args.input_vcf, args.output_vcf, args.min_aligned, args.split_weight, args.disc_weight,
                args.num_samp, args.lib_info_path, args.debug, args.alignment_outpath, args.ref_fasta, args.sum_quals,
                args.max_reads, args.max_ci_dist)


def cli():
    try:
        sys.exit(main())
    except IOError as e:
        if e.errno != 32:  # ignore SIGPIPE
            raise


if __name__ == '__main__':
    cli()
