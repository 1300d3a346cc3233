#!/usr/bin/env python
"""Whole-function measurement of the drop-in: `sso_genotype(bam, vcf_in, vcf_out, ...)` of this repo
(native evidence packer + CUDA engine) beside the reference's own `sso_genotype` (oracle/_ref, py3-patched,
serial and with its own process pool) on the reference's fixture BAM and a VCF made of the fixture's 212
records replicated K times (IDs made unique).  Output VCFs must be identical (modulo ##fileDate).
Prints one JSON line.  Needs a GPU for this repo's arm; `--no-gpu` patches the oracle over genotype.score (CPU check
of the plumbing only, not a product configuration).  `--batch-size` / `--cores` are passed to this repo's arm."""
import argparse
import io
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
DATA = os.path.join(REPO, "tests", "data")
BAM = os.path.join(DATA, "NA12878.target_loci.sorted.bam")
VCF = os.path.join(DATA, "example.vcf")
LIB = os.path.join(DATA, "NA12878.bam.json")


def replicated_vcf(k, path):
    lines = open(VCF).read().splitlines()
    head = [l for l in lines if l.startswith("#")]
    body = [l for l in lines if not l.startswith("#")]
    with open(path, "w") as f:
        f.write("\n".join(head) + "\n")
        for r in range(k):
            for l in body:
                c = l.split("\t")
                c[2] = "%s_r%d" % (c[2], r)
                c[7] = ";".join(("MATEID=%s_r%d" % (kv[7:], r)) if kv.startswith("MATEID=") else kv
                                for kv in c[7].split(";"))
                f.write("\t".join(c) + "\n")
    return len(body) * k


def strip(text):
    return [l for l in text.splitlines() if not l.startswith("##fileDate=")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--ref-cores", type=int, default=0, help="cores for the reference's pool (0 = all)")
    ap.add_argument("--no-gpu", action="store_true")
    ap.add_argument("--batch-size", type=int, default=1000)
    ap.add_argument("--cores", type=int, default=0)
    args = ap.parse_args()
    tmp = "/tmp/svgt_dropin_%d.vcf" % os.getpid()
    n_rec = replicated_vcf(args.reps, tmp)

    from svtyper_b200 import genotype, singlesample
    if args.no_gpu:
        from oracle import oracle
        from svtyper_b200 import compact as cp
        genotype.score = lambda batch, **p: oracle.score(cp.wide_from_compact(batch), **p)

    def ours():
        out = io.StringIO()
        with open(tmp) as fin:
            singlesample.sso_genotype(BAM, fin, out, 20, 1, 1, 1000000, LIB, False, None, False, 1000, 1e10,
                                      args.cores or None, args.batch_size)
        return out.getvalue()
    ours()                                              # warm: CUDA context, library load
    t0 = time.perf_counter()
    mine = ours()
    t_ours = time.perf_counter() - t0

    from oracle import ref_loader
    import resource
    res = {"metric": "sso_genotype wall time, whole function", "records": n_rec, "reps_of_fixture": args.reps,
           "batch_size": args.batch_size, "max_rss_mb": resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1024.0,
           "ours": {"seconds": t_ours, "records_per_s": n_rec / t_ours,
                    "scorer": "oracle (CPU check)" if args.no_gpu else "CUDA engine"}}
    if ref_loader.ensure():
        ref = ref_loader.load()
        cores = args.ref_cores or os.cpu_count()
        for label, c in (("reference_serial", None), ("reference_pool", cores)):
            out_path = tmp + "." + label
            t0 = time.perf_counter()
            with open(tmp) as fin, open(out_path, "w") as fout:
                ref.singlesample.sso_genotype(BAM, fin, fout, 20, 1, 1, 1000000, LIB, False, None, False, 1000, 1e10,
                                              c, 1000)
            dt = time.perf_counter() - t0
            theirs = open(out_path).read()
            os.unlink(out_path)
            res[label] = {"seconds": dt, "records_per_s": n_rec / dt, "cores": c or 1,
                          "identical_to_ours": strip(theirs) == strip(mine)}
        res["speedup_vs_reference_pool"] = res["reference_pool"]["seconds"] / t_ours
        res["speedup_vs_reference_serial"] = res["reference_serial"]["seconds"] / t_ours
        res["note"] = ("the reference runs on this repo's stdlib BAM reader as its pysam stand-in (pysam is not "
                       "installed), so its read gathering is slower than with htslib; its scoring is its own")
    os.unlink(tmp)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
