#!/usr/bin/env python
"""bench.py -- SV breakpoints genotyped / second on N B200s (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the scoring path (svgt_compact_kernel + svgt_call_compact_kernel through the C ABI) over
one resident batch of synthetic evidence in the compact 16-byte-row schema: config "del1m4lib"
(BASELINE.json configs[3]: 1M DEL breakpoints, 4 read-group libraries with distinct insert-size histograms;
it fits one GPU).  With N > 1 the run measures BOTH readings of "1 -> 8 B200 site shard":

  weak    every rank holds its own `--sites` breakpoints (per-GPU work fixed);
  strong  ONE global batch of `--sites` breakpoints, cut into contiguous site ranges balanced by evidence
          rows (svtyper_b200/shard.py), rank r scoring range r.

In both, the 80-byte output rows of every rank land in a buffer rank 0 exports through CUDA IPC, with one flag
word per rank and no collective inside the step (`--gather`, several may be listed, the first is the line's):
  dma   (default) each rank's copy engine forwards a finished step's rows over NVLink on a side stream, under the
        NEXT step's kernels; the flag follows the copy in stream order;
  peer  the call kernel stores its rows straight into rank 0's buffer (the transfer then sits at the end of the step);
  nccl  one NCCL gather per step, double-buffered (round 1's route; also the fallback when IPC mapping fails).
`--scaling` picks which of weak / strong is the line's `value` (default weak); the other is reported under its own
key.  One JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
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
INT_FIELDS = ("GT", "GQ", "DP", "RO", "AO", "QR", "QA", "RS", "AS", "ASC", "RP", "AP")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="del1m4lib")
    ap.add_argument("--sites", type=int, default=1_000_000, help="breakpoints per GPU (weak) / in the global batch (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=0, help="sites in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the site-sharded (strong) measurement")
    ap.add_argument("--strong-only", action="store_true",
                    help="N > 1: only rank 0 builds the global batch, the other ranks just their shard (for shapes whose "
                         "per-rank copy would not fit the host: stress1m at 1M sites); implies --scaling strong")
    ap.add_argument("--gather", default="dma", help="comma-separated list of dma | peer | nccl; the first is the line's")
    return ap.parse_args()


def workload_name(args):
    if args.config == "del1m4lib":
        return "del1m4lib: %d DEL breakpoints, 4 libraries, <=1000 reads/site (BASELINE.json configs[3])" % args.sites
    if args.config == "stress1m":
        return "stress1m: %d sites, max_reads=10000 ragged evidence incl. empty and skipped sites (BASELINE.json configs[4])" % args.sites
    return "%s: %d breakpoints" % (args.config, args.sites)


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


def ncu_traffic(config, sites):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json);
    only valid for the workload it was captured on, and only reported together with the capture's git hash."""
    try:
        with open(os.path.join(REPO, "profiles", "traffic.json")) as f:
            tr = json.load(f)
        if int(tr.get("workload_sites", -1)) == int(sites) and tr.get("config") == config and tr.get("git"):
            return tr
    except Exception:
        pass
    return None


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
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample_sites_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": pool.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def to_compact(wide, alloc):
    """Wide synthetic rows -> the compact schema (native converter: threaded, no temporaries; numpy twin otherwise)."""
    from svtyper_b200 import compact as cp, packer
    if packer.available():
        return packer.compact_from_wide(wide, min_aligned=20, alloc=alloc)
    return cp.compact_from_wide(wide, alloc=alloc)


class RowGather(object):
    """Where every rank's output rows end up: rank 0's buffer (see the module docstring for the three routes).

    dma / peer: rank 0 owns a CUDA-IPC-exported buffer (two halves, alternating by step) plus one flag word per
    rank; every rank maps both.  nccl: one dist.gather of the padded shards per step.
    """

    def __init__(self, mode, rank, world, counts, device):
        import torch
        import torch.distributed as dist
        from svtyper_b200 import native
        self.rank, self.world, self.counts = rank, world, list(counts)
        self.offsets = [sum(self.counts[:r]) for r in range(world)]
        self.total = sum(self.counts)
        self.mode = mode if world > 1 else "none"
        self.lib = native.lib()
        self.torch = torch
        self.buf = self.flags = None
        self._owned = []
        self.err = None
        if self.mode in ("peer", "dma"):
            ok = 1
            handles = [None, None]
            try:
                if rank == 0:
                    b, f = ctypes.c_void_p(), ctypes.c_void_p()
                    hb, hf = ctypes.create_string_buffer(64), ctypes.create_string_buffer(64)
                    native.check(self.lib.svgt_shared_alloc(2 * max(self.total, 1) * 80, ctypes.byref(b), hb))
                    native.check(self.lib.svgt_shared_alloc(4096, ctypes.byref(f), hf))
                    self.buf, self.flags = b.value, f.value
                    self._owned = [b.value, f.value]
                    handles = [hb.raw, hf.raw]
            except Exception as e:      # noqa: BLE001
                ok = 0
                self.err = str(e)
            dist.broadcast_object_list(handles, src=0)
            if rank != 0 and handles[0] is not None:
                try:
                    b, f = ctypes.c_void_p(), ctypes.c_void_p()
                    native.check(self.lib.svgt_shared_open(handles[0], ctypes.byref(b)))
                    native.check(self.lib.svgt_shared_open(handles[1], ctypes.byref(f)))
                    self.buf, self.flags = b.value, f.value
                except Exception as e:  # noqa: BLE001
                    ok = 0
                    self.err = str(e)
            if handles[0] is None:
                ok = 0
            t = torch.tensor([ok], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if int(t.item()) == 0:
                self.mode = "nccl"
        n_mine = self.counts[rank]
        if self.mode == "dma":
            self.outs = [torch.zeros((max(n_mine, 1), 80), dtype=torch.uint8, device=device) for _ in range(2)]
            self.side = torch.cuda.Stream(device=device)
            self.sent = [None, None]
        if self.mode == "nccl":
            pad = max(self.counts)
            self.pad = pad
            self.send = [torch.zeros((max(pad, 1), 80), dtype=torch.uint8, device=device) for _ in range(2)]
            self.recv = [[torch.empty_like(self.send[0]) for _ in range(world)] for _ in range(2)] if rank == 0 else [None, None]
            self.pending = [None, None]
        self.step_no = 0

    def _dst(self, half):
        return self.buf + (half * self.total + self.offsets[self.rank]) * 80

    def arm(self, desc, half):
        """Point this rank's launch descriptor at where the coming step's final rows go."""
        self.step_no += 1
        if self.mode == "peer":
            desc.out_final = self._dst(half)
            desc.done_flag = self.flags + 4 * self.rank
            desc.done_value = self.step_no
        else:
            desc.out_final = None
            desc.done_flag = None

    def out_tensor(self, dev_out, half):
        if self.mode == "nccl":
            return self.send[half]
        if self.mode == "dma":
            return self.outs[half]
        return dev_out

    def before_score(self, stream, half):
        if self.mode == "nccl" and self.pending[half] is not None:
            self.pending[half].wait()
            self.pending[half] = None
        if self.mode == "dma" and self.sent[half] is not None:
            stream.wait_event(self.sent[half])          # the copy that read outs[half] two steps ago
            self.sent[half] = None

    def after_score(self, stream, half):
        import torch.distributed as dist
        from svtyper_b200 import native
        if self.mode == "peer" and self.rank == 0:
            native.check(self.lib.svgt_wait_flags(ctypes.c_void_p(self.flags), self.world, self.step_no,
                                                  ctypes.c_void_p(stream.cuda_stream)))
        elif self.mode == "dma":
            done = self.torch.cuda.Event()
            done.record(stream)
            self.side.wait_event(done)
            side = ctypes.c_void_p(self.side.cuda_stream)
            native.check(self.lib.svgt_peer_copy(ctypes.c_void_p(self._dst(half)), ctypes.c_void_p(self.outs[half].data_ptr()),
                                                 self.counts[self.rank] * 80, side))
            native.check(self.lib.svgt_set_flag(ctypes.c_void_p(self.flags + 4 * self.rank), self.step_no, side))
            ev = self.torch.cuda.Event()
            ev.record(self.side)
            self.sent[half] = ev
        elif self.mode == "nccl":
            self.pending[half] = dist.gather(self.send[half], self.recv[half], dst=0, async_op=True)

    def drain(self, stream):
        """Everything sent so far has landed on rank 0 (enqueued on `stream`)."""
        from svtyper_b200 import native
        if self.mode == "nccl":
            for h in (0, 1):
                self.before_score(stream, h)
        elif self.mode == "dma":
            for h in (0, 1):
                if self.sent[h] is not None:
                    stream.wait_event(self.sent[h])
            if self.rank == 0 and self.step_no:
                native.check(self.lib.svgt_wait_flags(ctypes.c_void_p(self.flags), self.world, self.step_no,
                                                      ctypes.c_void_p(stream.cuda_stream)))

    def gathered_rows(self, half):
        """rank 0: all ranks' rows of the last step written into `half`, as OUT_DTYPE numpy rows."""
        import numpy as np
        import torch
        from svtyper_b200 import evidence as ev
        if self.rank != 0:
            return None
        if self.mode in ("peer", "dma"):
            from svtyper_b200 import native
            host = np.empty(self.total * 80, dtype=np.uint8)
            torch.cuda.synchronize()
            native.check(self.lib.svgt_memcpy_d2h(ctypes.c_void_p(host.ctypes.data),
                                                  ctypes.c_void_p(self.buf + half * self.total * 80), self.total * 80))
            return host.view(ev.OUT_DTYPE).copy()
        parts = [self.recv[half][r][:self.counts[r]].cpu().numpy() for r in range(self.world)]
        return np.concatenate(parts, axis=0).reshape(-1).view(ev.OUT_DTYPE).copy()

    def close(self):
        if self.mode in ("peer", "dma"):
            self.torch.cuda.synchronize()
            if self.rank == 0:
                for p in self._owned:
                    self.lib.svgt_shared_free(ctypes.c_void_p(p))
            else:
                self.lib.svgt_shared_close(ctypes.c_void_p(self.buf))
                self.lib.svgt_shared_close(ctypes.c_void_p(self.flags))


def timed_steps(eng, dev, gather, steps, warmup, world, local_rank, stream):
    """W warm-up + K timed steps of one resident batch; returns (ms_total max over ranks, per-step kernel ms,
    clocks, launches).  Every step ends with this rank's rows in rank 0's buffer."""
    import torch
    import torch.distributed as dist

    def step(i):
        half = i & 1
        gather.before_score(stream, half)
        gather.arm(dev.desc, half)
        eng.score(dev, stream, out=gather.out_tensor(dev.out, half))
        gather.after_score(stream, half)

    for i in range(max(warmup, 3)):
        step(i)
    gather.drain(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    eng.check(dev)
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    launches0 = eng.launches
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0.record(stream)
    for i in range(steps):
        half = i & 1
        gather.before_score(stream, half)
        gather.arm(dev.desc, half)
        k_ev[i][0].record(stream)
        eng.score(dev, stream, out=gather.out_tensor(dev.out, half))
        k_ev[i][1].record(stream)
        gather.after_score(stream, half)
    gather.drain(stream)
    ev1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    kern_ms = [a.elapsed_time(b) for a, b in k_ev]
    if world > 1:
        tt = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    eng.check(dev)
    return ms_total, kern_ms, clocks, eng.launches - launches0, (steps - 1) & 1


def measure_strong(args, rank, world, local_rank, device, eng, stream, batch, procs, modes, own_rows):
    """ONE global batch (rank 0's), cut by shard.shard_bounds into contiguous site ranges balanced by evidence rows;
    rank r regenerates and scores range r; all rows land on rank 0.  Returns the `strong` dict (rank 0's is complete)."""
    import torch
    import torch.distributed as dist
    from svtyper_b200 import compact as cp, shard, synth
    bounds = [0] * (world + 1)
    if rank == 0:
        bounds = shard.shard_bounds(batch, world)
    dist.broadcast_object_list(bounds, src=0)
    lo, hi = bounds[rank], bounds[rank + 1]
    if rank == 0:
        my = batch.slice_sites(lo, hi)
    else:
        w = synth.generate_parallel(args.config, n_sites=args.sites, rank=0, procs=procs, site_range=(lo, hi), bucket=False)
        my = to_compact(w, None)
        del w
    my.order = my.length_order()
    t_rows = torch.zeros(world, dtype=torch.int64, device=device)
    t_rows[rank] = my.n_rows
    dist.all_reduce(t_rows)
    rows_per_rank = [int(x) for x in t_rows.tolist()]
    sdev = eng.upload(my)
    scounts = [bounds[r + 1] - bounds[r] for r in range(world)]
    sg = RowGather(modes[0], rank, world, scounts, device)
    s_ms, s_kern, s_clocks, s_launches, s_half = timed_steps(eng, sdev, sg, args.steps, args.warmup, world, local_rank, stream)
    s_parity = None
    if rank == 0:
        g = sg.gathered_rows(s_half)
        s_parity = {"sites": int(g.shape[0]),
                    "sharded_rows_byte_identical_to_one_gpu": bool(g.tobytes() == own_rows.tobytes())}
    mean_rows = sum(rows_per_rank) / float(world)
    strong = {"value": args.sites * args.steps / (s_ms * 1e-3), "unit": UNIT, "ms_per_step": s_ms / args.steps,
              "sites_total": int(args.sites), "sites_per_rank": scounts, "rows_per_rank": rows_per_rank,
              "row_imbalance": (max(rows_per_rank) / mean_rows - 1.0) if mean_rows else 0.0,
              "kernel_ms_avg_rank0": sum(s_kern) / len(s_kern), "gather": sg.mode, "parity": s_parity,
              "unit_mode": int(sdev.desc.unit_mode), "clocks": s_clocks, "gpu_launches": s_launches,
              "partition": "shard.shard_bounds: contiguous site ranges balanced by evidence rows"}
    sg.close()
    del sdev
    return strong


def run_strong_only(args, rank, world, local_rank, device, eng, stream, batch, procs, t_gen):
    """--strong-only: the global batch lives on rank 0 alone (host and GPU); prints a line whose value is the sharded run."""
    import numpy as np
    import torch.distributed as dist
    from svtyper_b200 import evidence as ev
    modes = [m.strip() for m in args.gather.split(",") if m.strip()] or ["dma"]
    own_rows, one_ms, oracle_parity = None, None, None
    if rank == 0:
        import torch
        dev = eng.upload(batch)
        for _ in range(3):
            eng.score(dev, stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            eng.score(dev, stream)
        e1.record(stream)
        torch.cuda.synchronize()
        one_ms = e0.elapsed_time(e1) / args.steps
        own_rows = eng.rows(dev)
        del dev
        if not args.no_cpu_baseline:                  # the oracle (C restatement, all host threads) on a slice of the batch
            from oracle import oracle
            from svtyper_b200 import compact as cp
            n = min(batch.n_sites, args.cpu_sample or 20000)
            want = oracle.score(cp.wide_from_compact(batch.slice_sites(0, n)), n_threads=oracle.max_threads())
            ok = all(np.array_equal(own_rows[:n][k], want[k]) for k in INT_FIELDS) and np.array_equal(own_rows[:n]["GL"], want["GL"])
            gt = own_rows["GT"]
            oracle_parity = {"sites": int(n), "int_fields_and_GL_bit_exact_vs_oracle": bool(ok),
                             "skipped_rows": int((gt == ev.GT_SKIPPED).sum()), "blank_rows": int((gt == ev.GT_BLANK).sum()),
                             "underflow_rows": int((gt == ev.GT_UNDERFLOW).sum())}
    strong = measure_strong(args, rank, world, local_rank, device, eng, stream, batch, procs, modes, own_rows)
    if rank == 0:
        peak, peak_src = measured_peak()
        alg, surv = batch.algorithmic_bytes(), batch.survey_bytes()
        line = {
            "metric": METRIC, "value": strong["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": strong["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "sites_total": batch.n_sites, "schema": "compact (16 B rows, 48 B site rows)",
                       "fragment_rows": batch.n_frag, "split_rows": batch.n_split, "algorithmic_bytes": alg,
                       "survey_8d_bytes": surv, "gather": strong["gather"], "gen_seconds": round(t_gen, 1)},
            "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
                         "one_gpu_ms_per_step": one_ms, "achieved": alg / (one_ms * 1e-3) / 1e9,
                         "frac": alg / (one_ms * 1e-3) / 1e9 / peak, "frac_on_survey_8d_bytes": surv / (one_ms * 1e-3) / 1e9 / peak,
                         "traffic": None, "note": "one GPU scoring the whole batch (rank 0), for reference"},
            "one_gpu": {"value": batch.n_sites / (one_ms * 1e-3), "ms_per_step": one_ms},
            "cpu_baseline": None, "e2e": None, "gpu_launches": strong["gpu_launches"], "clocks": strong["clocks"],
            "parity": oracle_parity, "strong": strong,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def run_ours(args, rank, world, local_rank):
    import numpy as np
    from svtyper_b200 import compact as cp, evidence as ev, shard, synth

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
                         "batches), %.1f s inside the reference's tally_variant_read_fragments + "
                         "bayesian_genotype (slowest worker of each pass; nothing else is timed)" % (cpu_sample.n_sites, passes, cores, dt)}
        # the same functions on ONE core ("also report single-core"): the first 512 sites, ~3 s
        one = cpu_sample.slice_sites(0, min(512, cpu_sample.n_sites))
        pool1 = cpu_baseline.ReferencePool(one, 1)
        pool1.step()
        p1, d1 = 0, 0.0
        while d1 < 3.0 and p1 < 50:
            d1 += pool1.step()[0]
            p1 += 1
        pool1.close()
        cpu["single_core"] = {"value": one.n_sites * p1 / d1, "unit": UNIT, "cores": 1, "kind": pool1.kind,
                              "sample": "first %d sites, %d passes, %.1f s" % (one.n_sites, p1, d1)}

    import torch
    import torch.distributed as dist
    from svtyper_b200 import engine, native
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    # ---- synthetic evidence, converted to the compact schema straight into pinned host memory
    pinned = {}

    def alloc(name, shape, dtype):
        t = torch.empty(tuple(shape), dtype=torch.int32, pin_memory=True)
        pinned[name] = t
        return t.numpy()

    strong_only = bool(args.strong_only and world > 1)
    if strong_only:
        args.scaling = "strong"
    procs = max(2, min(32, cores // max(world, 1) - 1))
    t_gen = time.time()
    if strong_only and rank != 0:
        batch = None                                  # this rank only ever sees its shard of rank 0's batch (below)
    else:
        wide = synth.generate_parallel(args.config, n_sites=args.sites, rank=rank,
                                       procs=max(2, min(32, cores - 2 * world)) if strong_only else procs)
        batch = to_compact(wide, alloc)
        del wide
    t_gen = time.time() - t_gen

    eng = engine.Engine(local_rank)
    stream = torch.cuda.current_stream()
    if strong_only:
        return run_strong_only(args, rank, world, local_rank, device, eng, stream, batch, procs, t_gen)
    dev = eng.upload(batch)

    # ---- weak: every rank its own batch, all rows to rank 0 (once per requested gather route)
    counts = [batch.n_sites] * world
    modes = [m.strip() for m in args.gather.split(",") if m.strip()] or ["dma"]
    dev.desc.out_final = None                         # one plain pass: this rank's rows in its own buffer
    dev.desc.done_flag = None
    eng.score(dev, stream)
    own_rows = eng.rows(dev)
    weak_runs = []
    for mi, mode in enumerate(modes if world > 1 else modes[:1]):
        gather = RowGather(mode, rank, world, counts, device)
        r_ms, r_kern, r_clocks, r_launches, r_half = timed_steps(eng, dev, gather, args.steps, args.warmup, world, local_rank,
                                                                 stream)
        ok = None
        if world > 1 and rank == 0:
            g = gather.gathered_rows(r_half)
            ok = bool(g[:batch.n_sites].tobytes() == own_rows.tobytes() and g.shape[0] == sum(counts))
        weak_runs.append({"gather": gather.mode, "requested": mode, "value": world * batch.n_sites * args.steps / (r_ms * 1e-3),
                          "ms_per_step": r_ms / args.steps, "kernel_ms_avg_rank0": sum(r_kern) / len(r_kern),
                          "gathered_rows_match": ok})
        if mi == 0:
            ms_total, kern_ms, clocks, launches, gather_mode, weak_gather_ok = r_ms, r_kern, r_clocks, r_launches, gather.mode, ok
        gather.close()
        dev.desc.out_final = None
        dev.desc.done_flag = None
    weak_value = world * batch.n_sites * args.steps / (ms_total * 1e-3)

    # ---- parity on the CPU sample (same sites, scored by the reference above)
    parity = None
    if cpu_rows is not None:
        got = own_rows[:cpu_sample.n_sites]
        ok = all(np.array_equal(got[k], cpu_rows[k]) for k in INT_FIELDS)
        ok = ok and bool(np.allclose(got["GL"], cpu_rows["GL"], rtol=0, atol=1e-6))
        parity = {"sites": int(cpu_sample.n_sites), "int_fields_bit_exact_and_GL_1e-6": bool(ok)}

    # ---- strong: rank 0's batch is THE global batch; rank r scores its row-balanced site range of it
    strong = None
    if world > 1 and not args.no_strong:
        strong = measure_strong(args, rank, world, local_rank, device, eng, stream, batch, procs, modes, own_rows)

    # ---- end to end through the host-buffer C ABI call (pinned host -> H2D -> kernels -> D2H)
    arrs = engine.host_arrays(batch)
    for k in ("sites", "rows"):
        arrs[k] = pinned[k]
    for k, a in list(arrs.items()):
        if k not in ("sites", "rows"):
            t = torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).pin_memory()
            arrs[k] = t
    out_pinned = torch.empty((batch.n_sites, ev.OUT_BYTES), dtype=torch.uint8, pin_memory=True)
    eng.score_host(batch, arrays=arrs, out=out_pinned)            # warm: allocates the staging buffers
    l0 = eng.launches
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        eng.score_host(batch, arrays=arrs, out=out_pinned)
    t_e2e = time.perf_counter() - t0
    e2e_launches = eng.launches - l0
    if world > 1:
        tt = torch.tensor([t_e2e], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item())
    e2e_same = bool(out_pinned.numpy().reshape(-1).view(ev.OUT_DTYPE).tobytes() == own_rows.tobytes())
    e2e_value = world * batch.n_sites * args.e2e_steps / t_e2e
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(eng.last_h2d),
           "d2h_bytes_per_step": int(eng.last_d2h), "ms_per_step": 1e3 * t_e2e / args.e2e_steps,
           "kernel_ms_inside": float(eng.last_kernel_ms), "gpu_launches": e2e_launches,
           "rows_identical_to_resident_path": e2e_same,
           "api": "svgt_ctx_score_host_compact (Engine.score_host), pinned host buffers, 8 site slices pipelined over 3 streams"}

    if rank == 0:
        peak, peak_src = measured_peak()
        alg = batch.algorithmic_bytes()
        surv = batch.survey_bytes()
        k_avg = sum(kern_ms) / len(kern_ms)
        achieved = alg / (k_avg * 1e-3) / 1e9
        tr = ncu_traffic(args.config, batch.n_sites)
        primary_strong = args.scaling == "strong" and strong is not None
        value = strong["value"] if primary_strong else weak_value
        ms_step = strong["ms_per_step"] if primary_strong else ms_total / args.steps
        weak = {"value": weak_value, "unit": UNIT, "ms_per_step": ms_total / args.steps, "sites_per_gpu": batch.n_sites,
                "gather": gather_mode, "gathered_rows_match": weak_gather_ok, "routes": weak_runs}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if primary_strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "sites_per_gpu": batch.n_sites, "schema": "compact (16 B rows, 48 B site rows)",
                       "fragment_rows_per_gpu": batch.n_frag, "split_rows_per_gpu": batch.n_split,
                       "algorithmic_bytes_per_gpu": alg, "survey_8d_bytes_per_gpu": surv,
                       "l2": "inputs (%.2f GB) larger than L2" % (alg / 1e9),
                       "gather": {"dma": "each rank's copy engine forwards a finished step's 80 B rows into rank 0's CUDA-IPC-mapped buffer over NVLink under the next step's kernels, one flag per rank; no collective in the step",
                                  "peer": "call kernels store their 80 B rows straight into rank 0's CUDA-IPC-mapped buffer over NVLink, one flag per rank; no collective in the step",
                                  "nccl": "nccl gather of 80 B rows to rank 0 every step, double-buffered under the next step",
                                  "none": "none"}[gather_mode],
                       "gen_seconds": round(t_gen, 1)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (tr or {}).get("dram_bytes_per_launch"),
                         "traffic_source": ({"file": "profiles/traffic.json", "git": tr["git"], "capture": tr.get("capture")} if tr else None),
                         "peak_source": peak_src, "kernel_ms_avg": k_avg, "kernel_ms_min": min(kern_ms),
                         "algorithmic_bytes": "48 B/site + 16 B/row (fragment and split rows) + 80 B/site out",
                         "frac_on_survey_8d_bytes": surv / (k_avg * 1e-3) / 1e9 / peak,
                         "survey_8d_formula": "48 + 16 F + 32 S + 80 bytes per site",
                         "frac_of_nominal_8TBs": achieved / 8000.0},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "parity": parity,
        }
        if world > 1:
            line["weak"] = weak
            line["strong"] = strong
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
