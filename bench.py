#!/usr/bin/env python
"""bench.py -- SV breakpoints genotyped / second on N B200s (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the scoring path over one resident batch of synthetic evidence:
config "del1m4lib" (BASELINE.json configs[3]: 1M DEL breakpoints, 4 read-group libraries with
distinct insert-size histograms, the config the 1->8 GPU site-shard metric is quoted on; it fits
one GPU).  Each rank holds its own `--sites` breakpoints (weak scaling); with N > 1 every step
ends with one NCCL gather of the 80-byte output rows to rank 0.  One JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "sv_breakpoints_genotyped_per_sec"
UNIT = "breakpoints/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="del1m4lib")
    ap.add_argument("--sites", type=int, default=1_000_000, help="breakpoints per GPU")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=0, help="sites in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", type=int, default=-1)
    return ap.parse_args()


def workload_name(args):
    return "%s: %d DEL breakpoints/GPU, 4 libraries, <=1000 reads/site (BASELINE.json configs[3])" % (
        args.config, args.sites) if args.config == "del1m4lib" else "%s: %d breakpoints/GPU" % (args.config, args.sites)


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.004):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.sm_max = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None),
                "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(sites):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture
    (profiles/traffic.json), valid only for the workload size it was captured at."""
    try:
        with open(os.path.join(REPO, "profiles", "traffic.json")) as f:
            tr = json.load(f)
        return tr if int(tr.get("workload_sites", -1)) == int(sites) else None
    except Exception:
        return None


def cpu_sample_of(batch, n):
    return batch.slice_sites(0, min(n, batch.n_sites))


# ------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (oracle/_ref), all host cores."""
    if rank != 0:
        return
    from svtyper_b200 import synth
    from oracle import cpu_baseline
    cores = os.cpu_count() or 1
    n = args.cpu_sample or min(args.sites, 512 * cores)
    batch = synth.generate(args.config, n_sites=n, rank=0, bucket=False)
    pool = cpu_baseline.ReferencePool(batch, cores)
    for _ in range(max(args.warmup, 1)):
        pool.step()
    t = 0.0
    for _ in range(args.steps):
        dt, _rows = pool.step()
        t += dt
    pool.close()
    value = n * args.steps / t
    sample = "%d sites of %s per step (generator stream 0), %d steps" % (n, args.config, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample_sites_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": pool.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import numpy as np
    from svtyper_b200 import evidence as ev, synth

    # ---- CPU baseline first (before CUDA is initialised in this process); rank 0, N = 1 only
    cpu = None
    cpu_rows = None
    cpu_sample = None
    cores = os.cpu_count() or 1
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline
        n = args.cpu_sample or min(args.sites, 512 * cores)
        # the sample = the first n sites of chunk 0 of rank 0's workload (same generator stream)
        first_chunk = min(25_000, args.sites)
        head = synth.generate(args.config, n_sites=first_chunk, seed=synth.BASE_SEED + synth.CONFIGS[args.config]["seed_off"],
                              rank=0, bucket=False)
        cpu_sample = head.slice_sites(0, min(n, head.n_sites))
        pool = cpu_baseline.ReferencePool(cpu_sample, cores)
        pool.step()                                   # untimed: imports, library tables, adapter objects
        passes, dt = 0, 0.0
        while dt < 10.0 and passes < 200:             # ~10 s of wall time on all cores
            t1, cpu_rows = pool.step()
            dt += t1
            passes += 1
        pool.close()
        cpu = {"value": cpu_sample.n_sites * passes / dt, "unit": UNIT, "cores": cores, "kind": pool.kind,
               "sample": "first %d sites of the workload, %d passes over %d worker processes (contiguous site "
                         "batches), %.1f s wall; only the reference's tally_variant_read_fragments + "
                         "bayesian_genotype are timed" % (cpu_sample.n_sites, passes, cores, dt)}

    import torch
    import torch.distributed as dist
    from svtyper_b200 import engine, native
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.variant >= 0:
        native.set_variant(args.variant)
    variant = native.set_variant(args.variant if args.variant >= 0 else -1)

    # ---- synthetic evidence, generated straight into pinned host memory
    pinned = {}

    def alloc(name, shape, dtype):
        t = torch.empty(tuple(shape), dtype=torch.int32, pin_memory=True)
        pinned[name] = t
        return t.numpy()

    procs = max(2, min(32, cores // max(world, 1) - 1))
    t_gen = time.time()
    batch = synth.generate_parallel(args.config, n_sites=args.sites, rank=rank, procs=procs, alloc=alloc)
    t_gen = time.time() - t_gen

    eng = engine.Engine(local_rank)
    dev = eng.upload(batch)
    stream = torch.cuda.current_stream()
    # N > 1: the one collective of the path -- a gather of the 80-byte rows to rank 0 -- is
    # double-buffered, so the gather of step i runs on NCCL's stream under the kernels of step i+1
    outs = [dev.out, torch.empty_like(dev.out)] if world > 1 else [dev.out]
    gathered = [[torch.empty_like(dev.out) for _ in range(world)] for _ in outs] if (world > 1 and rank == 0) \
        else [None for _ in outs]
    pending = [None for _ in outs]

    def step(i):
        b = i % len(outs)
        if pending[b] is not None:
            pending[b].wait()                 # the stream (not the host) waits for the gather that read outs[b]
            pending[b] = None
        eng.score(dev, stream, out=outs[b])
        if world > 1:
            pending[b] = dist.gather(outs[b], gathered[b], dst=0, async_op=True)

    def drain():
        for b in range(len(outs)):
            if pending[b] is not None:
                pending[b].wait()
                pending[b] = None

    for i in range(max(args.warmup, 3)):
        step(i)
    drain()
    torch.cuda.synchronize()
    eng.check(dev)

    # ---- timed region: K steps, CUDA events on the launching stream, max over ranks
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    launches0 = eng.launches
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0.record(stream)
    for i in range(args.steps):
        b = i % len(outs)
        if pending[b] is not None:
            pending[b].wait()
            pending[b] = None
        k_ev[i][0].record(stream)
        eng.score(dev, stream, out=outs[b])
        k_ev[i][1].record(stream)
        if world > 1:
            pending[b] = dist.gather(outs[b], gathered[b], dst=0, async_op=True)
    drain()
    ev1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    kern_ms = [a.elapsed_time(b) for a, b in k_ev]
    launches = eng.launches - launches0
    if world > 1:
        tt = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    eng.check(dev)

    # ---- parity on the CPU sample (same sites, scored by the reference above)
    parity = None
    if cpu_rows is not None:
        got = eng.rows(dev)[:cpu_sample.n_sites]
        ok = all(np.array_equal(got[k], cpu_rows[k]) for k in
                 ("GT", "GQ", "DP", "RO", "AO", "QR", "QA", "RS", "AS", "ASC", "RP", "AP"))
        ok = ok and bool(np.allclose(got["GL"], cpu_rows["GL"], rtol=0, atol=1e-6))
        parity = {"sites": int(cpu_sample.n_sites), "int_fields_bit_exact_and_GL_1e-6": bool(ok)}

    # ---- end to end through the host-buffer C ABI call (pinned host -> H2D -> kernel -> D2H)
    arrs = engine.host_arrays(batch)
    out_pinned = torch.empty((batch.n_sites, ev.OUT_BYTES), dtype=torch.uint8, pin_memory=True)
    eng.score_host(batch, arrays=arrs, out=out_pinned)            # warm: allocates the staging buffers
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        eng.score_host(batch, arrays=arrs, out=out_pinned)
    t_e2e = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([t_e2e], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item())
    e2e_value = world * batch.n_sites * args.e2e_steps / t_e2e
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(eng.last_h2d),
           "d2h_bytes_per_step": int(eng.last_d2h), "ms_per_step": 1e3 * t_e2e / args.e2e_steps,
           "kernel_ms_inside": float(eng.last_kernel_ms), "api": "svgt_ctx_score_host (Engine.score_host), pinned host buffers"}

    if rank == 0:
        peak, peak_src = measured_peak()
        alg = batch.algorithmic_bytes()
        k_avg = sum(kern_ms) / len(kern_ms)
        achieved = alg / (k_avg * 1e-3) / 1e9
        tr = ncu_traffic(batch.n_sites)
        line = {
            "metric": METRIC, "value": world * batch.n_sites * args.steps / (ms_total * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "sites_per_gpu": batch.n_sites,
                       "fragment_rows_per_gpu": batch.n_frag, "split_rows_per_gpu": batch.n_split,
                       "algorithmic_bytes_per_gpu": alg, "l2": "inputs (%.2f GB) larger than L2" % (alg / 1e9),
                       "kernel_variant": variant, "gather": "nccl gather of 80 B rows to rank 0 every step, double-buffered under the next step" if world > 1 else "none",
                       "gen_seconds": round(t_gen, 1)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (tr or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                         "kernel_ms_avg": k_avg, "kernel_ms_min": min(kern_ms),
                         "frac_of_nominal_8TBs": achieved / 8000.0},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "parity": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
