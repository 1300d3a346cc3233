"""Drop-in for `svtyper.classic.sv_genotype` (reference svtyper/classic.py:107-533).

Same positional/keyword signature, same VCF-in / VCF-out text contract; the per-breakpoint
evidence tally and Bayesian call run on the GPU in one batch per sample (see genotype.py).
`alignment_outpath` (the diagnostic evidence-BAM dump, classic.py:161-166) is outside the
accelerated path and is refused rather than silently ignored.
"""
from __future__ import annotations

import logging
import os
import sys

from . import evidence as ev
from . import gather, genotype, packer, vcf
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
                max_ci_dist):
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

    header_lines, body = vcf.split_header_and_body(vcf_in)
    if not body:                      # the reference emits its header with the first record
        vcf_in.close()
        vcf_out.close()
        return
    header = vcf.VcfHeader().parse(header_lines)
    header.ensure_svtyper_fields()
    for s in samples:
        if s.name not in header.samples:
            header.add_sample(s.name)

    # ---- walk the records: pass-through lines and genotyped sites, in output order ----
    plan = genotype.SitePlan()
    open_bnds = {}
    for line in body:
        rec = vcf.VcfRecord(line.rstrip().split("\t"), header)
        if not sum_quals:
            rec.qual = 0
        if not rec.has_svtype():
            genotype.warn("Warning: SVTYPE missing at variant %s. Skipping.\n" % rec.var_id)
            plan.passthrough(rec)
            continue
        svtype = rec.svtype()
        if svtype not in ("BND", "DEL", "DUP", "INV"):
            genotype.warn("Warning: Unsupported SVTYPE at variant %s (%s). Skipping.\n" % (rec.var_id, svtype))
            plan.passthrough(rec)
            continue
        if svtype == "BND":
            mate_id = rec.info["MATEID"]
            if mate_id not in open_bnds:
                open_bnds[rec.var_id] = rec
                continue
            first = open_bnds.pop(mate_id)
            plan.site(first, rec, vcf.bnd_breakpoint(first, rec, max_ci_dist))
        else:
            plan.site(rec, None, vcf.simple_breakpoint(rec, max_ci_dist))

    # ---- one batch per sample through the engine ----
    rows = {}
    for s in samples:
        batch = genotype.pack_sample(
            s, plan, lambda smp, bp: gather.gather_classic(smp, bp, genotype.Z, max_reads), min_aligned,
            mode=packer.MODE_CLASSIC, max_reads=max_reads)
        rows[s.name] = genotype.score(batch, min_aligned=min_aligned, split_slop=genotype.SPLIT_SLOP,
                                      split_weight=split_weight, disc_weight=disc_weight,
                                      assoc_mode=ev.ASSOC_CLASSIC)

    # ---- write ----
    vcf_out.write(header.render() + "\n")
    for kind, rec, mate, idx in plan.entries:
        if kind == "site":
            for s in samples:
                genotype.apply_row(rec, s.name, rows[s.name][idx], classic=True)
        vcf_out.write(rec.render() + "\n")
        if mate is not None:
            mate.adopt_calls(rec)
            vcf_out.write(mate.render() + "\n")
    if open_bnds:
        logging.warning("Unpaired breakends found in file. These will not be present in output.")
    vcf_in.close()
    vcf_out.close()
    for s in samples:
        s.close()
    return
