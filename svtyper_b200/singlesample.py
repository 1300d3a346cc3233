"""Drop-in for `svtyper.singlesample.sso_genotype` (reference svtyper/singlesample.py:764-816).

Same signature and VCF contract.  The reference's serial and `mp.Pool` modes differ only in
who gathers and scores which breakpoint; both produce the same VCF.  Here `cores` / `batch_size`
keep their meaning as host-side fan-out of the (CPU) read gathering, while every batch of
breakpoints is scored by the CUDA engine.
"""
from __future__ import annotations

import os
import sys

from . import evidence as ev
from . import gather, genotype, packer, vcf
from .sample import SampleInfo, write_sample_json


def _read_vcf_text(vcf_in):
    path = os.path.abspath(vcf_in.name)
    if os.path.basename(path) == "<stdin>":
        return list(vcf_in)
    with open(path, "r") as f:
        return list(f)


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
    lines = _read_vcf_text(vcf_in)
    bam_path = os.path.abspath(bam_string)
    if not (bam_path.endswith(".bam") or bam_path.endswith(".cram")):
        sys.exit("Error: %s is not a valid alignment file (*.bam or *.cram)\n" % bam_path)
    sample = SampleInfo.open(bam_path, lib_info_path, ref_fasta, num_samp)
    if lib_info_path is not None and not os.path.exists(lib_info_path):
        write_sample_json([sample], lib_info_path)

    # the reference reads only the '##' lines into the header and then appends the BAM's
    # sample: the output has exactly one sample column (singlesample.py:112-124)
    meta = []
    for line in lines:
        if line.startswith("##"):
            meta.append(line)
        else:
            break
    header = vcf.VcfHeader().parse(meta)
    header.ensure_svtyper_fields()
    header.add_sample(sample.name)
    body = [l for l in lines if not l.startswith("#")]

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
        if rec.svtype() not in ("BND", "DEL", "DUP", "INV"):
            genotype.warn("Warning: Unsupported SVTYPE at variant %s (%s). Skipping.\n" % (rec.var_id, rec.svtype()))
            plan.passthrough(rec)
            continue
        if rec.svtype() == "BND":
            mate_id = rec.info["MATEID"]
            if mate_id not in open_bnds:
                open_bnds[rec.var_id] = rec
                continue
            first = open_bnds.pop(mate_id)
            plan.site(first, rec, vcf.bnd_breakpoint(first, rec, max_ci_dist))
        else:
            plan.site(rec, None, vcf.simple_breakpoint(rec, max_ci_dist))

    batch = genotype.pack_sample(
        sample, plan, lambda smp, bp: gather.gather_sso(smp, bp, genotype.Z, max_reads), min_aligned,
        mode=packer.MODE_SSO, max_reads=max_reads)
    rows = genotype.score(batch, min_aligned=min_aligned, split_slop=genotype.SPLIT_SLOP,
                          split_weight=split_weight, disc_weight=disc_weight, assoc_mode=ev.ASSOC_SSO)

    vcf_out.write(header.render() + "\n")
    fast = genotype.RowFormatter(header, sample.name, rows)
    out = []
    for kind, rec, mate, idx in plan.entries:
        if kind == "site" and fast.eligible(rec) and (mate is None or fast.eligible(mate)):
            qual, fmt, call = fast.columns(rec, idx)
            out.append(fast.line(rec, qual, fmt, call))
            if mate is not None:                    # BND mates share one genotype (singlesample.py:648-652)
                out.append(fast.line(mate, qual, fmt, call))
            continue
        if kind == "site":
            genotype.apply_row(rec, sample.name, rows[idx], classic=False)
        out.append(rec.render())
        if mate is not None:
            mate.adopt_calls(rec)
            out.append(mate.render())
    if out:
        vcf_out.write("\n".join(out) + "\n")
    sample.close()
