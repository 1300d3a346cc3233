"""Small end-to-end exercise of every kernel of libsvgt.so for `compute-sanitizer --tool memcheck`:
tally kernel (plain and with a piece plan), replay kernel, both call kernel variants are too big a batch to run
under the sanitizer, so: the small-batch call kernel, the wide-row cross-check kernels, the host path, and
malformed piece plans (which must be flagged, not read out of bounds).  Prints MEMCHECK-SCRIPT-OK at the end."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from svtyper_b200 import compact as cp, engine, evidence as ev, native, synth   # noqa: E402

eng = engine.Engine(0)
b = synth.generate("stress1m", n_sites=int(os.environ.get("SITES", "300")), seed=5)
cb = cp.compact_from_wide(b)
base = {}
for assoc in (ev.ASSOC_SSO, ev.ASSOC_CLASSIC):
    dev = eng.upload(cb, assoc_mode=assoc)
    eng.score(dev)
    base[assoc] = eng.rows(dev).tobytes()
    for k in (1, 3):
        dev = eng.upload(cb, assoc_mode=assoc, piece_chunks=k)
        assert dev.plan is not None
        eng.score(dev)
        assert eng.rows(dev).tobytes() == base[assoc], (assoc, k)
    for um in (1, 2, 3):
        dev = eng.upload(cb, assoc_mode=assoc, unit_mode=um)
        eng.score(dev)
        assert eng.rows(dev).tobytes() == base[assoc], (assoc, um)
os.environ["SVGT_PLAN_FORCE_CHUNKS"] = "2"
assert eng.score_host(cb).tobytes() == base[ev.ASSOC_SSO] and eng.last_pieces > 0
del os.environ["SVGT_PLAN_FORCE_CHUNKS"]
assert eng.score_host(cb).tobytes() == base[ev.ASSOC_SSO]
for v in native.VARIANTS:
    native.set_variant(v)
    dev = eng.upload(b)
    eng.score(dev)
    eng.rows(dev)
native.set_variant(-1)
import torch
for what in ("entry", "piece_rows", "piece_scratch", "heavy"):
    dev = eng.upload(cb, piece_chunks=1)
    if what == "entry":
        dev.tensors["plan_entries"][0] = cb.n_sites + 5
    elif what == "piece_rows":
        dev.tensors["plan_pieces"][0, 2] = 1 << 20
    elif what == "piece_scratch":
        dev.tensors["plan_pieces"][0, 3] = dev.plan.scratch_chunks
    else:
        dev.tensors["plan_heavy"][0, 2] += 1
    dev.tensors["plan_scratch"].fill_(255)      # whatever an unscored piece leaves behind: lead counts of -1, NaN addends
    torch.cuda.synchronize()
    eng.score(dev)
    try:
        eng.rows(dev)
        raise AssertionError("malformed plan not flagged: " + what)
    except native.SvgtError as e:
        assert e.code == native.ERR_ARG, what
print("MEMCHECK-SCRIPT-OK")
