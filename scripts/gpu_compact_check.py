"""GPU box: parity of the compact-row kernel against the oracle on every shape, then a timing
of the benchmark shape next to the wide-row kernel.  Usage: python scripts/gpu_compact_check.py [--sites N]"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from svtyper_b200 import compact as cp, engine, evidence as ev, native, synth   # noqa: E402
from oracle import oracle                                                        # noqa: E402
from util import assert_rows_match                                               # noqa: E402


def rows_compact(eng, b, unit_mode=0, **kw):
    cb = b if isinstance(b, cp.CompactBatch) else cp.compact_from_wide(b)
    dev = eng.upload(cb, unit_mode=unit_mode, **kw)
    eng.score(dev)
    return eng.rows(dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=1_000_000)
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--config", default="del1m4lib")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--cache", default="", help="directory holding / receiving the generated batch as .npy files")
    ap.add_argument("--no-wide", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--tag", default="")
    ap.add_argument("--pieces", default="off", help="comma list of timed variants of the compact batch, each "
                    "<pieces>[/<unit_mode>]: piece plan auto | off | <max chunks>; unit_mode used with the plan (1 full units, 3 ramp)")
    args = ap.parse_args()
    import torch
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    eng = engine.Engine(0)
    fails = 0
    if not args.skip_parity:
        cases = []
        for config, n in (("del10k", 10_000), ("mixed100k", 20_000), ("del1m4lib", 20_000), ("stress1m", 6_000)):
            cases.append((config, synth.generate(config, n_sites=n)))
        cases.append(("hazard", synth.hazard_batch()))
        z = np.load(os.path.join(REPO, "tests", "golden", "fixture_evidence.npz"))
        from util import batch_from_npz
        cases.append(("fixture", batch_from_npz(z)))
        mean, sd, hist = synth.fixture_library()
        libs = ev.LibraryTable([(299.999999999, 50.0, hist), (mean, sd, hist)])
        cases.append(("unsafe-lib", synth.generate("mixed100k", n_sites=4000, seed=11, libs=libs)))
        h2 = {i: 19 for i in range(100, 400)}
        h2.update({i: 1 for i in range(400, 1400)})
        h2.update({i: 361 for i in range(1400, 1500)})
        tl = ev.LibraryTable([(300.0, 50.0, h2)])
        cases.append(("tie-del", synth.generate("del10k", n_sites=4000, seed=3, libs=tl)))
        cases.append(("tie-mixed", synth.generate("mixed100k", n_sites=4000, seed=3, libs=tl)))
        l70 = ev.LibraryTable([synth.gaussian_library(300 + 7 * i, 40 + i) for i in range(70)])
        cases.append(("70libs", synth.generate("del1m4lib", n_sites=2000, seed=9, libs=l70)))
        big = ev.LibraryTable([synth.gaussian_library(4000, 900)])
        cases.append(("bighist", synth.generate("del10k", n_sites=2000, seed=10, libs=big)))
        for nn in (1, 31, 33, 1000):
            cases.append(("n=%d" % nn, synth.generate("mixed100k", n_sites=nn, seed=5, bucket=False)))
        for name, b in cases:
            for assoc in (ev.ASSOC_SSO, ev.ASSOC_CLASSIC):
                exp = oracle.score(b, assoc_mode=assoc, n_threads=oracle.max_threads())
                for um in (0, 1, 2):
                    try:
                        got = rows_compact(eng, b, um, assoc_mode=assoc)
                        assert_rows_match(got, exp, exact_gl=True, where="%s assoc%d um%d" % (name, assoc, um))
                    except Exception as e:      # noqa: BLE001
                        fails += 1
                        print("FAIL %s assoc=%d unit_mode=%d: %s" % (name, assoc, um, str(e)[:300]), flush=True)
            print("checked", name, flush=True)
        b = synth.generate("mixed100k", n_sites=6000, seed=31)
        cb = cp.compact_from_wide(b)
        exp = oracle.score(b)
        got = eng.score_host(cb)
        try:
            assert_rows_match(got, exp, exact_gl=True, where="host path")
        except Exception as e:      # noqa: BLE001
            fails += 1
            print("FAIL host path:", str(e)[:300])
        print("PARITY FAILS:", fails, flush=True)

    # ---- timing
    t0 = time.time()
    cdir = args.cache and os.path.join(args.cache, "%s_%d" % (args.config, args.sites))
    wide = None
    if cdir and os.path.exists(os.path.join(cdir, "rows.npy")):
        libs = synth.make_libraries(synth.CONFIGS[args.config]["n_lib"])
        cb = cp.CompactBatch(np.load(os.path.join(cdir, "sites.npy")), np.load(os.path.join(cdir, "rows.npy")), libs,
                             np.load(os.path.join(cdir, "order.npy")))
    else:
        wide = synth.generate_parallel(args.config, n_sites=args.sites)
        cb = cp.compact_from_wide(wide)
        if cdir:
            os.makedirs(cdir, exist_ok=True)
            np.save(os.path.join(cdir, "sites.npy"), cb.sites)
            np.save(os.path.join(cdir, "rows.npy"), cb.rows)
            np.save(os.path.join(cdir, "order.npy"), cb.order)
        if args.no_wide:
            wide = None
    print("gen %.1fs: sites %d compact rows %d (%d frag + %d split); alg bytes compact %.3f GB survey %.3f GB" % (
        time.time() - t0, cb.n_sites, cb.n_rows, cb.n_frag, cb.n_split, cb.algorithmic_bytes() / 1e9, cb.survey_bytes() / 1e9),
        flush=True)
    res = {"tag": args.tag, "lib": os.path.basename(native.LIB_PATH)}
    plist = args.pieces.split(",")
    runs = [("compact" if i == 0 else "compact_pieces_" + pc, cb, pc) for i, pc in enumerate(plist)]
    if wide is not None:
        runs.append(("wide-lean", wide, None))
    for label, batch, pc in runs:
        kw = {}
        if pc is not None:
            pcs, _, um = pc.partition("/")
            kw["piece_chunks"] = None if pcs == "auto" else 0 if pcs == "off" else int(pcs)
            if um:
                kw["piece_unit_mode"] = int(um)
        dev = eng.upload(batch, **kw)
        for _ in range(3):
            eng.score(dev)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b_ in evs:
            a.record()
            eng.score(dev)
            b_.record()
        torch.cuda.synchronize()
        eng.check(dev)
        ms = [a.elapsed_time(b_) for a, b_ in evs]
        res[label] = {"ms_avg": sum(ms) / len(ms), "ms_min": min(ms)}
        if pc is not None:
            res[label].update({"pieces": pc, "plan": dev.plan_info, "unit_mode": int(dev.desc.unit_mode),
                               "frac_own_B": cb.algorithmic_bytes() / (res[label]["ms_avg"] * 1e-3) / 1e9 / 6545.9,
                               "frac_survey_B": cb.survey_bytes() / (res[label]["ms_avg"] * 1e-3) / 1e9 / 6545.9})
        rows = eng.rows(dev)
        res[label]["gt_hist"] = np.bincount(rows["GT"] + 3, minlength=6).tolist()
        if label == "compact":
            keep = rows
        else:
            res[label]["identical_to_first"] = bool(keep.tobytes() == rows.tobytes())
        del dev
    k = res["compact"]["ms_avg"] * 1e-3
    res["compact"]["frac_own_B"] = cb.algorithmic_bytes() / k / 1e9 / 6545.9
    res["compact"]["frac_survey_B"] = cb.survey_bytes() / k / 1e9 / 6545.9
    res["compact"]["Msites_per_s"] = cb.n_sites / k / 1e6
    import hashlib
    res["rows_md5"] = hashlib.md5(keep.tobytes()).hexdigest()
    if args.no_e2e:
        print(json.dumps(res), flush=True)
        with open(os.path.join(REPO, "gpurun_out", "compact_check_%s.json" % (args.tag or "x")), "w") as f:
            json.dump(res, f)
        return 1 if fails else 0
    # end to end through the host API (pinned)
    arrs = engine.host_arrays(cb)
    pin = {}
    for kx, a in arrs.items():
        t = torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).pin_memory()
        pin[kx] = t
    out = torch.empty((cb.n_sites, 80), dtype=torch.uint8, pin_memory=True)
    eng.score_host(cb, arrays=pin, out=out)
    t1 = time.perf_counter()
    for _ in range(3):
        eng.score_host(cb, arrays=pin, out=out)
    dt = (time.perf_counter() - t1) / 3
    res["e2e"] = {"ms": dt * 1e3, "Msites_per_s": cb.n_sites / dt / 1e6, "h2d": eng.last_h2d, "kernel_ms_inside": eng.last_kernel_ms,
                  "same_rows": bool(out.numpy().reshape(-1).view(ev.OUT_DTYPE).tobytes() == keep.tobytes())}
    print(json.dumps(res), flush=True)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", "compact_check_%s.json" % (args.tag or "base")), "w") as f:
        json.dump({"fails": fails, "res": res}, f)
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
