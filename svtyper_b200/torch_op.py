"""`torch.ops.svgt.score_batch`: the scoring path as a registered PyTorch operator (inner seam).

Device tensors in, one device tensor of 80-byte output rows out, asynchronous on the current CUDA stream: a thin
wrapper over the C ABI's svgt_score_compact (include/svgt.h) for callers who keep their evidence in torch tensors.
Nothing is computed by PyTorch.  Per-site error flags (svgt_err) come back in the second result, a 4-word int32
status tensor (status[0] = first error code, 0 = none), because the call does not synchronise.

    out, status = torch.ops.svgt.score_batch(sites_i32[Ns, 12], rows_i32[Nr, 4], order_i32[Ns] or None,
                                             lib_f64[Nl, 4], lib_i32[Nl, 4], hist_i32[Nh], pm_f64[256],
                                             logt_f64[Nlog], consts_f64[32], split_weight, disc_weight,
                                             min_aligned, split_slop, assoc_mode, unit_mode)
    out: uint8 [Ns, 80]  = evidence.OUT_DTYPE rows (GL[3] f64, SQ f64, GT GQ DP RO AO QR QA RS AS ASC RP AP int32)
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import native


@torch.library.custom_op("svgt::score_batch", mutates_args=(), device_types="cuda")
def score_batch(sites: torch.Tensor, rows: torch.Tensor, order: Optional[torch.Tensor], lib_f64: torch.Tensor,
                lib_i32: torch.Tensor, hist: torch.Tensor, pm: torch.Tensor, logt: torch.Tensor, consts: torch.Tensor,
                split_weight: float, disc_weight: float, min_aligned: int, split_slop: int, assoc_mode: int,
                unit_mode: int) -> Tuple[torch.Tensor, torch.Tensor]:
    for name, t, dt in (("sites", sites, torch.int32), ("rows", rows, torch.int32), ("lib_f64", lib_f64, torch.float64),
                        ("lib_i32", lib_i32, torch.int32), ("hist", hist, torch.int32), ("pm", pm, torch.float64),
                        ("logt", logt, torch.float64), ("consts", consts, torch.float64)):
        if t.dtype != dt or not t.is_contiguous() or not t.is_cuda:
            raise ValueError("svgt::score_batch: %s must be a contiguous CUDA tensor of %s" % (name, dt))
    n_sites = sites.shape[0]
    d = native.SvgtCBatch()
    d.sites, d.n_sites = sites.data_ptr(), n_sites
    d.rows, d.n_rows = rows.data_ptr(), rows.shape[0]
    d.order = None if order is None else order.data_ptr()
    d.lib_f64, d.lib_i32, d.n_lib = lib_f64.data_ptr(), lib_i32.data_ptr(), lib_f64.shape[0]
    d.hist, d.n_hist, d.hist_max = hist.data_ptr(), hist.numel(), 0          # 0: the kernel checks the counts itself
    d.pm, d.logt, d.n_log, d.consts = pm.data_ptr(), logt.data_ptr(), logt.numel(), consts.data_ptr()
    d.min_aligned, d.rows_min_aligned, d.split_slop = int(min_aligned), int(min_aligned), int(split_slop)
    d.assoc_mode, d.unit_mode = int(assoc_mode), int(unit_mode)
    d.split_weight, d.disc_weight = float(split_weight), float(disc_weight)
    out = torch.zeros((max(n_sites, 1), 80), dtype=torch.uint8, device=sites.device)
    status = torch.zeros(4, dtype=torch.int32, device=sites.device)
    with torch.cuda.device(sites.device):
        stream = torch.cuda.current_stream(sites.device)
        native.check(native.lib().svgt_score_compact(ctypes.byref(d), ctypes.c_void_p(out.data_ptr()),
                                                     ctypes.c_void_p(status.data_ptr()), ctypes.c_void_p(stream.cuda_stream)))
    return out[:n_sites], status


@score_batch.register_fake
def _(sites, rows, order, lib_f64, lib_i32, hist, pm, logt, consts, split_weight, disc_weight, min_aligned, split_slop,
      assoc_mode, unit_mode):
    return sites.new_empty((sites.shape[0], 80), dtype=torch.uint8), sites.new_empty((4,), dtype=torch.int32)
