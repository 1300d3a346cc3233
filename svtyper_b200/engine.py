"""Host-side driver of the scoring kernels: EvidenceBatch -> device -> OUT_DTYPE rows.

PyTorch is used for what it is good at here -- device memory, streams, pinned host
buffers, torch.distributed -- and nothing else: every byte of arithmetic happens in
libsvgt.so (svtyper_b200/csrc), reached through the C ABI of include/svgt.h.

Two call shapes, both replacing the reference's per-breakpoint
`tally_variant_read_fragments` + `bayesian_genotype` pair
(reference svtyper/singlesample.py:523-536) for a whole batch:

  Engine.score_host(batch)      host numpy arrays in, OUT_DTYPE numpy rows out
                                (svgt_ctx_score_host: H2D, kernel, D2H)
  Engine.upload(batch) + Engine.score(dev)   device-resident, asynchronous
                                (svgt_score_batch)
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import compact as cp
from . import evidence as ev
from . import native

_LUT_CACHE = {}


def luts_for(n_log):
    """Host LUTs (CPython math), cached by rounded-up table size."""
    size = 1 << max(10, int(n_log - 1).bit_length())
    if size not in _LUT_CACHE:
        _LUT_CACHE[size] = ev.build_luts(size)
    return _LUT_CACHE[size]


def _torch():
    import torch
    return torch


class PinnedArena(object):
    """Reusable pinned (page-locked) host buffers, one per array name: `alloc(name, shape, dtype)` returns a numpy
    view of the first prod(shape) int32 words, growing the buffer geometrically when a batch is larger than any
    before.  The packers / converters write compact rows straight into it and svgt_ctx_score_host_compact copies
    from it asynchronously at full PCIe rate (pageable memory would be staged by the driver, at about half)."""

    def __init__(self):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("pinned host memory needs a CUDA device")
        self._torch = torch
        self._bufs = {}

    def alloc(self, name, shape, dtype):
        if np.dtype(dtype) != np.int32:
            return np.empty(shape, dtype=dtype)
        n = int(np.prod(shape)) if len(shape) else 1
        t = self._bufs.get(name)
        if t is None or t.numel() < n:
            t = self._torch.empty(max(n + n // 2, 1024), dtype=self._torch.int32, pin_memory=True)
            self._bufs[name] = t
        return t.numpy()[:n].reshape(shape)


class DeviceBatch(object):
    """An EvidenceBatch resident in HBM plus its svgt_batch_t descriptor."""

    def __init__(self, tensors, desc, n_sites, algorithmic_bytes):
        self.tensors = tensors          # keeps the device memory alive
        self.desc = desc                # native.SvgtBatch / native.SvgtCBatch with device pointers
        self.compact = isinstance(desc, native.SvgtCBatch)
        self.n_sites = n_sites
        self.algorithmic_bytes = algorithmic_bytes
        self.out = None
        self.status = None
        self.plan = None                # native.SvgtSegPlan (kept alive: desc.plan points at it), or None
        self.plan_info = None           # {"max_chunks", "pieces", "heavy_sites", "scratch_bytes"} of that plan


def _descriptor(ptr, batch, n_log, min_aligned, split_slop, split_weight, disc_weight, assoc_mode):
    d = native.SvgtBatch()
    d.sites, d.n_sites = ptr["sites"], batch.n_sites
    d.frags, d.n_frag = ptr["frags"], batch.n_frag
    d.splits, d.n_split = ptr["splits"], batch.n_split
    d.order = ptr.get("order")
    d.lib_f64, d.lib_i32, d.n_lib = ptr["lib_f64"], ptr["lib_i32"], batch.libs.n_lib
    d.hist, d.n_hist = ptr["hist"], int(batch.libs.hist.size)
    d.pm, d.logt, d.n_log, d.consts = ptr["pm"], ptr["logt"], int(n_log), ptr["consts"]
    d.min_aligned, d.split_slop, d.assoc_mode = int(min_aligned), int(split_slop), int(assoc_mode)
    d.split_weight, d.disc_weight = float(split_weight), float(disc_weight)
    return d


def _cdescriptor(ptr, batch, n_log, min_aligned, split_slop, split_weight, disc_weight, assoc_mode,
                 unit_mode=0, flags=0):
    d = native.SvgtCBatch()
    d.sites, d.n_sites = ptr["sites"], batch.n_sites
    d.rows, d.n_rows = ptr["rows"], batch.n_rows
    d.order = ptr.get("order")
    d.lib_f64, d.lib_i32, d.n_lib = ptr["lib_f64"], ptr["lib_i32"], batch.libs.n_lib
    d.hist, d.n_hist = ptr["hist"], int(batch.libs.hist.size)
    d.hist_max = int(batch.libs.hist.max()) if batch.libs.hist.size else 1
    d.pm, d.logt, d.n_log, d.consts = ptr["pm"], ptr["logt"], int(n_log), ptr["consts"]
    d.min_aligned, d.split_slop, d.assoc_mode = int(min_aligned), int(split_slop), int(assoc_mode)
    d.unit_mode = max(int(unit_mode), 0)
    d.split_weight, d.disc_weight = float(split_weight), float(disc_weight)
    d.flags = int(flags)
    d.rows_min_aligned = int(batch.min_aligned)
    return d


def host_arrays(batch, split_weight=1.0, disc_weight=1.0):
    """name -> contiguous numpy array for every array field of svgt_batch_t / svgt_cbatch_t."""
    pm, logt, consts = luts_for(batch.log_table_size(split_weight, disc_weight))
    if isinstance(batch, cp.CompactBatch):
        arrs = {"sites": batch.sites, "rows": batch.rows}
    else:
        arrs = {"sites": batch.sites, "frags": batch.frags, "splits": batch.splits}
    arrs.update({
        "lib_f64": batch.libs.lib_f64, "lib_i32": batch.libs.lib_i32, "hist": batch.libs.hist,
        "pm": pm, "logt": logt, "consts": consts,
    })
    if batch.order is not None:
        arrs["order"] = batch.order
    return {k: np.ascontiguousarray(v) for k, v in arrs.items()}


class Engine(object):
    """One scoring engine per process / GPU."""

    def __init__(self, device=None):
        torch = _torch()
        self._lib = native.lib()            # raises if libsvgt.so is missing
        if not torch.cuda.is_available() or self._lib.svgt_device_count() <= 0:
            raise native.SvgtError(native.ERR_NO_DEVICE, "no CUDA device; svtyper_b200 has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self._ctx = ctypes.c_void_p()
        native.check(self._lib.svgt_ctx_create(self.device.index, ctypes.byref(self._ctx)))
        self.last_h2d = self.last_d2h = 0
        self.last_kernel_ms = 0.0
        self.launches = 0                   # kernels launched through this engine
        self.last_pieces = 0                # pieces the last score_host call cut its long sites into

    def close(self):
        if self._ctx:
            self._lib.svgt_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- host buffers in/out
    def score_host(self, batch, min_aligned=20, split_slop=3, split_weight=1.0, disc_weight=1.0,
                   assoc_mode=ev.ASSOC_SSO, arrays=None, out=None, site_order=True, unit_mode=0):
        """Score a host batch (CompactBatch: the default path; EvidenceBatch: wide rows, the
        compatibility entry); returns OUT_DTYPE rows (numpy).

        `arrays` may carry pre-built (e.g. pinned) host arrays from `host_arrays()`;
        `out` a pre-allocated (e.g. pinned) uint8/OUT_DTYPE buffer of n_sites rows.
        `site_order`: the compact rows are laid out in site order (true for every packer and
        converter of this package), which lets large batches be pipelined in site slices.
        """
        arrs = arrays if arrays is not None else host_arrays(batch, split_weight, disc_weight)
        ptr = {k: (v.data_ptr() if hasattr(v, "data_ptr") else v.ctypes.data) for k, v in arrs.items()}
        n_log = arrs["logt"].numel() if hasattr(arrs["logt"], "numel") else arrs["logt"].size
        if out is None:
            out = np.zeros(batch.n_sites, dtype=ev.OUT_DTYPE)
        optr = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
        if isinstance(batch, cp.CompactBatch):
            if unit_mode == 0:
                unit_mode = batch.suggest_unit_mode()
            desc = _cdescriptor(ptr, batch, n_log, min_aligned, split_slop, split_weight, disc_weight, assoc_mode,
                                unit_mode, native.LAYOUT_SITE_ORDER if site_order else 0)
            rc = self._lib.svgt_ctx_score_host_compact(self._ctx, ctypes.byref(desc), ctypes.c_void_p(optr))
            launches = 2 if batch.n_sites else 0
            npc = ctypes.c_int64()
            self._lib.svgt_ctx_last_pieces(self._ctx, ctypes.byref(npc))
            self.last_pieces = npc.value
            if npc.value:
                launches += 1               # svgt_replay_pieces_kernel
        else:
            desc = _descriptor(ptr, batch, n_log, min_aligned, split_slop, split_weight, disc_weight, assoc_mode)
            rc = self._lib.svgt_ctx_score_host(self._ctx, ctypes.byref(desc), ctypes.c_void_p(optr))
            launches = self._lib.svgt_launches_per_batch(ctypes.byref(desc))
        h2d, d2h, ms = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_float()
        self._lib.svgt_ctx_last_traffic(self._ctx, ctypes.byref(h2d), ctypes.byref(d2h))
        self._lib.svgt_ctx_last_kernel_ms(self._ctx, ctypes.byref(ms))
        self.last_h2d, self.last_d2h, self.last_kernel_ms = h2d.value, d2h.value, ms.value
        if isinstance(batch, cp.CompactBatch) and h2d.value and batch.n_sites >= 131072 and site_order:
            launches *= 8               # pipelined: one launch pair per site slice
        self.launches += launches
        native.check(rc)
        return out

    # ---------------------------------------------------------------- device resident
    def upload(self, batch, min_aligned=20, split_slop=3, split_weight=1.0, disc_weight=1.0,
               assoc_mode=ev.ASSOC_SSO, unit_mode=0, piece_chunks=0, piece_unit_mode=1):
        """unit_mode: 0 = pick from the batch's row counts (CompactBatch.suggest_unit_mode), else svgt_cbatch_t's
        values (1 full units, 2 two-site units, 3 ramped units); -1 = leave the choice to the library (by site count).
        piece_chunks (compact batches): 0 = no piece plan (the default: measured, a plan does not pay on any shape
        of BASELINE.json -- profiles/README.md), None = let the library plan pieces for sites too long for one warp
        (svgt_plan_count's policy), k > 0 = pieces of at most k 32-row chunks; piece_unit_mode: the unit_mode used
        with a plan (1: full 6-entry units -- no entry is longer than a piece, so there is no tail to ramp for)."""
        torch = _torch()
        arrs = host_arrays(batch, split_weight, disc_weight)
        tens = {}
        for k, a in arrs.items():
            if a.dtype == np.uint32:
                a = a.view(np.int32)
            t = torch.from_numpy(a) if a.size else torch.zeros(4, dtype=torch.from_numpy(a).dtype)
            tens[k] = t.to(self.device)
        ptr = {k: t.data_ptr() for k, t in tens.items()}
        if isinstance(batch, cp.CompactBatch):
            if unit_mode == 0:                      # auto: decided from the sites' row counts (we have them here)
                unit_mode = batch.suggest_unit_mode()
            desc = _cdescriptor(ptr, batch, arrs["logt"].size, min_aligned, split_slop, split_weight,
                                disc_weight, assoc_mode, unit_mode, native.LAYOUT_SITE_ORDER)
        else:
            desc = _descriptor(ptr, batch, arrs["logt"].size, min_aligned, split_slop, split_weight,
                               disc_weight, assoc_mode)
        dev = DeviceBatch(tens, desc, batch.n_sites, batch.algorithmic_bytes())
        if isinstance(batch, cp.CompactBatch) and piece_chunks != 0 and unit_mode != 2 and batch.n_sites:
            with torch.cuda.device(self.device):    # the planner sizes pieces for this device's resident warps
                pl = native.plan_pieces(batch.sites, min_aligned, split_slop, 0, int(piece_chunks or 0))
            if pl is not None:
                for k in ("entries", "pieces", "heavy"):
                    tens["plan_" + k] = torch.from_numpy(pl[k]).to(self.device)
                nbytes = pl["scratch_chunks"] * native.PLAN_CHUNK_BYTES
                tens["plan_scratch"] = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
                sp = native.SvgtSegPlan()
                sp.entries, sp.n_entries = tens["plan_entries"].data_ptr(), int(pl["entries"].shape[0])
                sp.pieces, sp.n_pieces = tens["plan_pieces"].data_ptr(), int(pl["pieces"].shape[0])
                sp.heavy, sp.n_heavy = tens["plan_heavy"].data_ptr(), int(pl["heavy"].shape[0])
                sp.scratch, sp.scratch_chunks = tens["plan_scratch"].data_ptr(), int(pl["scratch_chunks"])
                dev.plan = sp
                desc.plan = ctypes.pointer(sp)
                desc.unit_mode = int(piece_unit_mode)
                dev.plan_info = {"max_chunks": pl["max_chunks"], "pieces": sp.n_pieces, "heavy_sites": sp.n_heavy,
                                 "scratch_bytes": int(nbytes)}
        dev.out = torch.zeros((max(batch.n_sites, 1), ev.OUT_BYTES), dtype=torch.uint8, device=self.device)
        dev.status = torch.zeros(4, dtype=torch.int32, device=self.device)
        return dev

    def score(self, dev, stream=None, out=None):
        """Asynchronously score a DeviceBatch on `stream` (default: torch's current stream).

        Returns the output tensor (`out` or dev.out: uint8 [n_sites, 80] on the device).  Call
        `check(dev)` after a synchronisation to surface per-site error flags.
        """
        torch = _torch()
        s = torch.cuda.current_stream(self.device) if stream is None else stream
        out = dev.out if out is None else out
        fn = self._lib.svgt_score_compact if dev.compact else self._lib.svgt_score_batch
        with torch.cuda.device(self.device):
            rc = fn(ctypes.byref(dev.desc), ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(dev.status.data_ptr()),
                    ctypes.c_void_p(s.cuda_stream))
        native.check(rc)
        if dev.compact:
            self.launches += (3 if dev.plan is not None else 2) if dev.n_sites else 0
        else:
            self.launches += self._lib.svgt_launches_per_batch(ctypes.byref(dev.desc))
        return out

    def check(self, dev):
        st = dev.status.cpu().numpy()
        if st[0] != 0:
            raise native.SvgtError(int(st[0]), "scoring kernel flagged %d thread(s)" % int(st[2]))

    def rows(self, dev):
        """Device results -> OUT_DTYPE numpy rows (synchronises)."""
        self.check(dev)
        raw = dev.out[:dev.n_sites].cpu().numpy()
        return raw.reshape(-1).view(ev.OUT_DTYPE).copy()
