#!/usr/bin/env python
"""Generate tests/golden/fixture_evidence.npz from the REFERENCE ITSELF.

Runs only where the reference is available (this build container:
/root/reference -> oracle/_ref via oracle/make_ref.py).  The committed .npz is what
travels to the GPU box.

For every breakpoint of the reference's own fixture (tests/data/example.vcf +
NA12878.target_loci.sorted.bam + NA12878.bam.json, reference
tests/test_singlesample.py:20-44) it records

  * the evidence rows produced by svtyper_b200.evidence.BatchPacker from the
    reference's own `gather_reads` output (reference singlesample.py:187-205), and
  * the reference's own answers: `tally_variant_read_fragments` counts
    (singlesample.py:355), `bayes_gt` log-likelihoods (statistics.py:23) and the
    `bayesian_genotype` FORMAT fields (singlesample.py:406), plus the classic
    entry point's debug dump (classic.py:415-421,450-451) for the classic
    association order.
"""
from __future__ import annotations

import contextlib
import io
import math
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import ref_loader  # noqa: E402
from svtyper_b200 import evidence as ev  # noqa: E402

DATA = os.path.join(REPO, "tests", "data")
VCF = os.path.join(DATA, "example.vcf")
BAM = os.path.join(DATA, "NA12878.target_loci.sorted.bam")
LIBJSON = os.path.join(DATA, "NA12878.bam.json")

EXPECT_DTYPE = np.dtype([
    ("counts", "<f8", (5,)),      # ref_seq, alt_seq, alt_clip, ref_span, alt_span (post-zeroing)
    ("GL", "<f8", (3,)), ("SQ", "<f8"),
    ("GT", "<i4"), ("GQ", "<i4"), ("DP", "<i4"), ("RO", "<i4"), ("AO", "<i4"),
    ("QR", "<i4"), ("QA", "<i4"), ("RS", "<i4"), ("AS", "<i4"), ("ASC", "<i4"),
    ("RP", "<i4"), ("AP", "<i4"),
])


def expected_row(ref, bp, counts, result):
    """Translate the reference's result dict into the numeric golden row."""
    row = np.zeros((), dtype=EXPECT_DTYPE)
    fm = result["formats"]
    row["counts"] = [counts[k] for k in ("ref_seq", "alt_seq", "alt_clip", "ref_span", "alt_span")]
    if fm["GL"] == ".":
        row["GT"], row["GQ"] = ev.GT_BLANK, -1
        return row
    for k in ("DP", "RO", "AO", "QR", "QA", "RS", "AS", "ASC", "RP", "AP"):
        row[k] = fm[k]
    row["GL"] = ref.statistics.bayes_gt(fm["QR"], fm["QA"], bp["svtype"] == "DUP")
    if fm["GT"] == "./.":
        row["GT"], row["GQ"], row["SQ"] = ev.GT_UNDERFLOW, -1, 0.0
    else:
        row["GT"] = {"0/0": 0, "0/1": 1, "1/1": 2}[fm["GT"]]
        row["GQ"] = fm["GQ"]
        row["SQ"] = fm["SQ"]
    return row


def main(out_path=os.path.join(HERE, "fixture_evidence.npz")):
    ref = ref_loader.load()
    ss = ref.singlesample
    sample = ss.setup_sample(BAM, LIBJSON, None, 1000000, 20)
    lib_names = list(sample.lib_dict.keys())
    libs = ev.LibraryTable([(sample.lib_dict[n].mean, sample.lib_dict[n].sd, sample.lib_dict[n].hist)
                            for n in lib_names])
    src_vcf = ss.init_vcf(VCF, sample, "/nonexistent-scratch")
    breakpoints = ss.collect_breakpoints(src_vcf, 1e10)

    packer = ev.BatchPacker(sample.bam.gettid, libs)
    rows, ids, svtypes = [], [], []
    for bp in breakpoints:
        regions = ss.get_breakpoint_regions(bp, sample, 3)
        frags, many = ss.gather_reads(sample.bam, bp["id"], regions, sample.rg_to_lib,
                                      sample.active_libs, 1000)
        assert not many
        packer.add_site(bp, frags, lib_index_of=lambda f: lib_names.index(f.lib.name))
        counts = ss.tally_variant_read_fragments(3, 20, bp, frags, False)
        if sum(counts.values()) == 0:
            result = ss.blank_genotype_result()
        else:
            result = ss.bayesian_genotype(bp, counts, 1, 1, False)
        rows.append(expected_row(ref, bp, counts, result))
        ids.append(bp["id"])
        svtypes.append(bp["svtype"])
    batch = packer.finish()

    # classic entry point: capture its debug dump (pre-zeroing counts + raw GL)
    buf = io.StringIO()
    with open(VCF) as inf, open(os.devnull, "w") as outf, contextlib.redirect_stdout(buf):
        ref.classic.sv_genotype(BAM, inf, outf, 20, 1, 1, 1000000, LIBJSON, True, None, None,
                                False, None, 1e10)
    blocks = buf.getvalue().split("--------------------------\n")[1:]
    classic = np.zeros(len(blocks), dtype=[("raw_counts", "<f8", (5,)), ("GL", "<f8", (3,)),
                                           ("has_gl", "<i4")])
    for i, blk in enumerate(blocks):
        vals = dict(re.findall(r"^(\w+): (\S+)$", blk, flags=re.M))
        classic["raw_counts"][i] = [float(vals[k]) for k in
                                    ("ref_seq", "alt_seq", "alt_clip", "ref_span", "alt_span")]
        m = re.search(r"^\(([^)]*)\)$", blk, flags=re.M)
        if m:
            classic["GL"][i] = [float(x) for x in m.group(1).split(",")]
            classic["has_gl"][i] = 1
    assert len(blocks) == len(breakpoints), (len(blocks), len(breakpoints))

    np.savez_compressed(
        out_path, sites=batch.sites, frags=batch.frags, splits=batch.splits,
        lib_f64=libs.lib_f64, lib_i32=libs.lib_i32, hist=libs.hist,
        lib_mean_sd=np.array([(sample.lib_dict[n].mean, sample.lib_dict[n].sd) for n in lib_names]),
        expected_sso=np.array(rows, dtype=EXPECT_DTYPE), expected_classic=classic,
        ids=np.array(ids), svtypes=np.array(svtypes))
    print("wrote %s: %d sites, %d fragment rows, %d split rows" %
          (out_path, batch.n_sites, batch.n_frag, batch.n_split))


if __name__ == "__main__":
    main()
