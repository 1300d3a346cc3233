"""CPU: libsvgt.so loads and exports every symbol include/svgt.h declares; argument errors
and the no-device error come back as codes (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from svtyper_b200 import build, native
from util import REPO


@pytest.fixture(scope="module")
def lib():
    build.build_native()
    return native.lib()


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "svgt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svgt_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    declared = _declared_symbols()
    assert sorted(native.SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_struct_size(lib):
    assert lib.svgt_abi_version() == native.ABI_VERSION


def test_header_is_plain_c_and_layout_matches_ctypes(tmp_path):
    """include/svgt.h compiles as C; sizeof/offsetof agree with the ctypes mirror."""
    import subprocess
    src = tmp_path / "layout.c"
    fields = [f[0] for f in native.SvgtBatch._fields_]
    body = "".join('printf("%%zu\\n", offsetof(svgt_batch_t, %s));' % f for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "svgt.h"\n'
                   'int main(void){printf("%zu\\n", sizeof(svgt_batch_t));printf("%zu\\n", sizeof(svgt_out_row_t));'
                   + body + 'return 0;}')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"),
                           str(src), "-o", str(exe)])
    vals = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert vals[0] == ctypes.sizeof(native.SvgtBatch)
    assert vals[1] == 80
    assert vals[2:] == [getattr(native.SvgtBatch, f).offset for f in fields]


def test_compact_struct_layout_matches_ctypes(tmp_path):
    import subprocess
    src = tmp_path / "layout.c"
    fields = [f[0] for f in native.SvgtCBatch._fields_]
    body = "".join('printf("%%zu\\n", offsetof(svgt_cbatch_t, %s));' % f for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "svgt.h"\n'
                   'int main(void){printf("%zu\\n", sizeof(svgt_cbatch_t));' + body + 'return 0;}')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"),
                           str(src), "-o", str(exe)])
    vals = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert vals[0] == ctypes.sizeof(native.SvgtCBatch)
    assert vals[1:] == [getattr(native.SvgtCBatch, f).offset for f in fields]


def test_null_batch_is_an_argument_error(lib):
    rc = lib.svgt_score_batch(None, None, None, None)
    assert rc == native.ERR_ARG
    assert b"null" in lib.svgt_last_error()
    assert lib.svgt_launches_per_batch(None) == native.ERR_ARG
    assert lib.svgt_score_compact(None, None, None, None) == native.ERR_ARG
    assert lib.svgt_ctx_score_host_compact(None, None, None) == native.ERR_ARG


def test_compact_argument_checks(lib):
    """Host-side validation of a compact batch needs no GPU: min_aligned the rows were packed for, unit_mode."""
    import numpy as np
    b = native.SvgtCBatch()
    z = np.zeros(64, dtype=np.float64)
    b.pm = b.logt = b.consts = z.ctypes.data
    b.n_log = 8
    b.min_aligned, b.rows_min_aligned = 20, 25
    assert lib.svgt_score_compact(ctypes.byref(b), None, None, None) == native.ERR_ARG
    assert b"min_aligned" in lib.svgt_last_error()
    b.rows_min_aligned = 20
    b.unit_mode = 7
    assert lib.svgt_score_compact(ctypes.byref(b), None, None, None) == native.ERR_ARG


def test_set_variant(lib):
    assert lib.svgt_set_variant(1) == 1
    assert lib.svgt_set_variant(0) == 0
    assert lib.svgt_set_variant(99) == native.ERR_ARG
    assert lib.svgt_set_variant(2) == native.ERR_ARG
    assert lib.svgt_set_variant(-1) in (0, 1)


def test_no_cpu_fallback(lib):
    """Without a CUDA device the compute entry points refuse to run."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    assert lib.svgt_device_count() == 0
    ctx = ctypes.c_void_p()
    assert lib.svgt_ctx_create(0, ctypes.byref(ctx)) == native.ERR_NO_DEVICE
    from svtyper_b200 import engine
    with pytest.raises(native.SvgtError):
        engine.Engine()


def test_torch_operator_is_registered():
    """SURVEY 8b's inner seam as a torch op: registered with the documented schema; CPU tensors are refused
    (there is no CPU implementation to dispatch to)."""
    import torch
    from svtyper_b200 import torch_op  # noqa: F401
    op = torch.ops.svgt.score_batch
    schema = str(op.default._schema)
    assert "Tensor sites" in schema and "Tensor? order" in schema and "-> (Tensor, Tensor)" in schema
    z = torch.zeros((1, 12), dtype=torch.int32)
    with pytest.raises(Exception):
        op(z, torch.zeros((0, 4), dtype=torch.int32), None, torch.zeros((1, 4), dtype=torch.float64),
           torch.zeros((1, 4), dtype=torch.int32), torch.zeros(4, dtype=torch.int32), torch.zeros(256, dtype=torch.float64),
           torch.zeros(8, dtype=torch.float64), torch.zeros(32, dtype=torch.float64), 1.0, 1.0, 20, 3, 0, 0)
