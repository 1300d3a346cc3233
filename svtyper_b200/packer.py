"""ctypes front-end of the native evidence packer (libsvgt_pack.so, include/svgt_pack.h).

`pack_sample(sample, plan, mode, max_reads)` is the native twin of
`genotype.pack_sample` (gather.gather_sso / gather.gather_classic + evidence.BatchPacker): one C
call gathers the reads around every planned breakpoint of a BAM (BGZF + BAI reader, the pysam
semantics of SURVEY.md 8c), groups them into fragments, runs the split-candidate QC and emits the
32-byte fragment / split rows the scoring kernel streams.  The Python path stays as its parity
checker (tests/test_pack_native.py) and serves inputs the native reader does not open (CRAM, a
pysam handle without a .bai).

Reference: gather_reads svtyper/classic.py:54-100, svtyper/singlesample.py:139-205;
SamFragment.add_read svtyper/parsers.py:748-768; SplitRead.is_valid svtyper/parsers.py:959-1058.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import evidence as ev

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "svgt_pack.cpp")
HEADER = os.path.join(HERE, "..", "include", "svgt_pack.h")
LIB_PATH = os.path.join(HERE, "libsvgt_pack.so")

MODE_SSO, MODE_CLASSIC = 0, 1
OK, ERR_ARG, ERR_IO, ERR_RG, ERR_RECORD = 0, -1, -2, -3, -4

_lib = None


class PackError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "svgt_pack error %d: %s" % (code, msg))
        self.code = code


class Site(ctypes.Structure):
    _fields_ = [("tidA", ctypes.c_int32), ("begA", ctypes.c_int32), ("endA", ctypes.c_int32),
                ("tidB", ctypes.c_int32), ("begB", ctypes.c_int32), ("endB", ctypes.c_int32)]


class Count(ctypes.Structure):
    _fields_ = [("n_frag_rows", ctypes.c_int32), ("n_split_rows", ctypes.c_int32),
                ("skip", ctypes.c_int32), ("n_fragments", ctypes.c_int32)]


class LibScan(ctypes.Structure):
    _fields_ = [("read_length", ctypes.c_int64), ("lib_records", ctypes.c_int64),
                ("records_seen", ctypes.c_int64), ("n_hist", ctypes.c_int64)]


def _stale():
    return (not os.path.exists(LIB_PATH) or
            any(os.path.getmtime(p) > os.path.getmtime(LIB_PATH) for p in (SRC, HEADER)))


def build(force=False):
    """g++ -> svtyper_b200/libsvgt_pack.so (host code only, links zlib).

    Normally run once by `__graft_entry__.build()` / at install time.  When a process finds the library
    missing or older than its source it rebuilds it here, safely for concurrent callers (torchrun ranks,
    worker processes): the compiler writes a private temporary file that is renamed over the target in one
    step, under an exclusive lock, so nobody ever maps a half-written library."""
    if not (force or _stale()):
        return LIB_PATH
    import fcntl
    import tempfile
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or _stale():                      # somebody else may have built it while we waited
                fd, tmp = tempfile.mkstemp(prefix=".libsvgt_pack.", suffix=".so", dir=HERE)
                os.close(fd)
                try:
                    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-shared", "-fPIC", "-o", tmp, SRC,
                                           "-lz", "-lpthread"])
                    os.replace(tmp, LIB_PATH)
                finally:
                    if os.path.exists(tmp):
                        os.unlink(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


_lib_error = None


def available():
    """True when libsvgt_pack.so can be built / loaded here (g++, zlib headers, a writable package directory
    or a prebuilt library); otherwise the Python gather path serves the request."""
    global _lib_error
    if _lib is not None:
        return True
    if _lib_error is not None:
        return False
    try:
        lib()
        return True
    except (OSError, subprocess.CalledProcessError, PermissionError) as e:
        _lib_error = e
        return False


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        L.svgt_pack_abi_version.restype = ctypes.c_int
        L.svgt_pack_last_error.restype = ctypes.c_char_p
        L.svgt_bam_open.restype = ctypes.c_int
        L.svgt_bam_open.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
        L.svgt_bam_close.restype = ctypes.c_int
        L.svgt_bam_close.argtypes = [ctypes.c_void_p]
        L.svgt_bam_n_references.restype = ctypes.c_int
        L.svgt_bam_n_references.argtypes = [ctypes.c_void_p]
        L.svgt_bam_reference_name.restype = ctypes.c_char_p
        L.svgt_bam_reference_name.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.svgt_bam_reference_length.restype = ctypes.c_int64
        L.svgt_bam_reference_length.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.svgt_bam_count.restype = ctypes.c_int64
        L.svgt_bam_count.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
        L.svgt_pack_sites.restype = ctypes.c_int
        L.svgt_pack_sites.argtypes = [ctypes.c_void_p, ctypes.POINTER(Site), ctypes.c_int64,
                                      ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_int32), ctypes.c_int32,
                                      ctypes.POINTER(ctypes.c_uint8), ctypes.c_int32, ctypes.c_int32, ctypes.c_int64,
                                      ctypes.c_int32, ctypes.POINTER(Count)]
        L.svgt_bam_scan_libraries.restype = ctypes.c_int
        L.svgt_bam_scan_libraries.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_char_p),
                                              ctypes.POINTER(ctypes.c_int32), ctypes.c_int32, ctypes.c_int32,
                                              ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(LibScan)]
        L.svgt_bam_scan_hist.restype = ctypes.c_int
        L.svgt_bam_scan_hist.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.POINTER(ctypes.c_int32)),
                                         ctypes.POINTER(ctypes.POINTER(ctypes.c_int64)), ctypes.POINTER(ctypes.c_int64)]
        L.svgt_pack_rows.restype = ctypes.c_int
        L.svgt_pack_rows.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.POINTER(ctypes.c_int32)),
                                     ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.POINTER(ctypes.c_int32)),
                                     ctypes.POINTER(ctypes.c_int64)]
        L.svgt_compact_count.restype = ctypes.c_int
        L.svgt_compact_count.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                         ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
        L.svgt_compact_fill.restype = ctypes.c_int
        L.svgt_compact_fill.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                        ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p]
        L.svgt_format_calls.restype = ctypes.c_int
        L.svgt_format_calls.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p,
                                        ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        L.svgt_host_sq.restype = ctypes.c_int
        L.svgt_host_sq.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32]
        L.svgt_format_quals.restype = ctypes.c_int
        L.svgt_format_quals.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        if L.svgt_pack_abi_version() != 2:
            raise OSError("libsvgt_pack.so ABI %d != expected 2 (stale build?)" % L.svgt_pack_abi_version())
        _lib = L
    return _lib


def _check(rc):
    if rc < 0:
        raise PackError(int(rc), lib().svgt_pack_last_error().decode("utf-8", "replace"))
    return rc


class NativeBam(object):
    """An indexed BAM opened by the native reader."""

    def __init__(self, path, bai=None):
        self._h = ctypes.c_void_p()
        _check(lib().svgt_bam_open(os.fsencode(path), os.fsencode(bai) if bai else None, ctypes.byref(self._h)))
        n = lib().svgt_bam_n_references(self._h)
        self.references = tuple(lib().svgt_bam_reference_name(self._h, i).decode("ascii") for i in range(n))
        self.lengths = tuple(int(lib().svgt_bam_reference_length(self._h, i)) for i in range(n))

    def close(self):
        if self._h:
            lib().svgt_bam_close(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def count(self, tid, beg, end, read_callback="nofilter"):
        return int(_check(lib().svgt_bam_count(self._h, tid, int(beg), int(end), 1 if read_callback == "all" else 0)))

    def pack(self, sites, rg_names, rg_lib, lib_active, mode, max_reads, threads=0):
        """sites: [(tidA, begA, endA, tidB, begB, endB)] -> (counts[n,4] int32, frags[n,8], splits[n,8])."""
        n = len(sites)
        flat = np.ascontiguousarray(np.asarray(sites, dtype=np.int32).reshape(-1, 6)) if n else np.zeros((1, 6), np.int32)
        arr = ctypes.cast(flat.ctypes.data, ctypes.POINTER(Site))
        counts = (Count * max(n, 1))()
        names = (ctypes.c_char_p * max(len(rg_names), 1))(*[r.encode("ascii") for r in rg_names])
        libs = (ctypes.c_int32 * max(len(rg_lib), 1))(*rg_lib)
        active = (ctypes.c_uint8 * max(len(lib_active), 1))(*[1 if a else 0 for a in lib_active])
        _check(lib().svgt_pack_sites(self._h, arr, n, names, libs, len(rg_names), active, len(lib_active), mode,
                                     -1 if max_reads is None else int(max_reads), int(threads), counts))
        fp, sp = ctypes.POINTER(ctypes.c_int32)(), ctypes.POINTER(ctypes.c_int32)()
        nf, ns = ctypes.c_int64(), ctypes.c_int64()
        _check(lib().svgt_pack_rows(self._h, ctypes.byref(fp), ctypes.byref(nf), ctypes.byref(sp), ctypes.byref(ns)))
        frags = (np.ctypeslib.as_array(fp, shape=(nf.value, ev.FRAG_WORDS)).copy() if nf.value
                 else np.zeros((0, ev.FRAG_WORDS), np.int32))
        splits = (np.ctypeslib.as_array(sp, shape=(ns.value, ev.SPLIT_WORDS)).copy() if ns.value
                  else np.zeros((0, ev.SPLIT_WORDS), np.int32))
        cnt = np.frombuffer(counts, dtype=np.int32).reshape(-1, 4)[:n].copy()
        return cnt, frags, splits


def _path_of(bam):
    name = getattr(bam, "filename", "") or ""
    return name.decode() if isinstance(name, bytes) else str(name)


def usable_path(bam):
    """Can the native reader open this alignment file?  (An indexed .bam on disk.)"""
    if os.environ.get("SVGT_PACKER", "").lower() == "python":
        return False
    path = _path_of(bam)
    if not path.endswith(".bam") or not os.path.exists(path):
        return False
    if not (os.path.exists(path + ".bai") or os.path.exists(os.path.splitext(path)[0] + ".bai")):
        return False
    return available()


def usable(sample):
    """Can the native reader serve this sample?"""
    return usable_path(sample.bam)


def scan_libraries(bam, readgroups_per_lib, num_samp, read_length_reads=10000, prevalence_records=100000):
    """One native pass over the head of the BAM -> [(read_length, lib_records, records_seen,
    {template length: count} in first-seen order)] per library (reference parsers.py:501-576)."""
    nb = NativeBam(_path_of(bam))
    try:
        names, libs = [], []
        for i, groups in enumerate(readgroups_per_lib):
            for g in groups:
                names.append(g.encode("ascii"))
                libs.append(i)
        n_lib = len(readgroups_per_lib)
        c_names = (ctypes.c_char_p * max(len(names), 1))(*names)
        c_libs = (ctypes.c_int32 * max(len(libs), 1))(*libs)
        out = (LibScan * max(n_lib, 1))()
        _check(lib().svgt_bam_scan_libraries(nb._h, c_names, c_libs, len(names), n_lib, int(num_samp),
                                             int(read_length_reads), int(prevalence_records), out))
        res = []
        for i in range(n_lib):
            kp, cp = ctypes.POINTER(ctypes.c_int32)(), ctypes.POINTER(ctypes.c_int64)()
            n = ctypes.c_int64()
            _check(lib().svgt_bam_scan_hist(nb._h, i, ctypes.byref(kp), ctypes.byref(cp), ctypes.byref(n)))
            hist = {int(kp[j]): int(cp[j]) for j in range(n.value)}
            res.append((int(out[i].read_length), int(out[i].lib_records), int(out[i].records_seen), hist))
        return res
    finally:
        nb.close()


def fetch_windows(sample, breakpoint, z, mode, flank=None):
    """The two fetch windows as the integers the BAM layer sees (reference classic.py:62-78,
    singlesample.py:139-157): the flank arithmetic is fp64, the reader truncates."""
    bam = sample.bam
    if flank is None:
        flank = sample.fetch_flank(z)
    out = []
    for side in ("A", "B"):
        end = breakpoint[side]
        tid = bam.gettid(end["chrom"])
        if tid < 0:
            raise ValueError("invalid contig %r" % (end["chrom"],))
        length = bam.lengths[tid]
        lo = max(end["pos"] + end["ci"][0] - flank, 0)
        hi = min(end["pos"] + end["ci"][1] + flank, length)
        out.extend((tid, max(0, int(lo)), int(hi)))
    return tuple(out)


def pack_sample(sample, plan, mode, max_reads, z=3, threads=0):
    """Native gather + pack of every planned site of one sample -> EvidenceBatch."""
    nb = getattr(sample, "native_bam", None)               # one native handle (header + index) per sample, reused
    if nb is None:                                          # by every chunk of the VCF; SampleInfo.close() closes it
        path = sample.bam.filename.decode() if isinstance(sample.bam.filename, bytes) else str(sample.bam.filename)
        nb = sample.native_bam = NativeBam(path)
    flank = sample.fetch_flank(z)                           # one value per sample: max(mean + z * sd) over its libraries
    sites = [fetch_windows(sample, bp, z, mode, flank) for bp in plan.breakpoints]
    rg_names = list(sample.rg_to_lib.keys())
    rg_lib = [sample.rg_to_lib[r] for r in rg_names]
    n_lib = len(sample.libraries)
    active = [i in sample.active for i in range(n_lib)]
    cnt, frags, splits = nb.pack(sites, rg_names, rg_lib, active, mode, max_reads, threads)
    rows = np.zeros((len(plan.breakpoints), ev.SITE_WORDS), dtype=np.int64)
    f_off = s_off = 0
    for i, bp in enumerate(plan.breakpoints):
        A, B = bp["A"], bp["B"]
        meta = ev.SVTYPE_CODE[bp["svtype"]]
        if A["is_reverse"]:
            meta |= ev.SITE_O1_REV
        if B["is_reverse"]:
            meta |= ev.SITE_O2_REV
        n_f, n_s, skip = int(cnt[i, 0]), int(cnt[i, 1]), int(cnt[i, 2])
        if skip:
            meta |= ev.SITE_SKIP
        rows[i] = (int(A["pos"]), int(B["pos"]), int(A["ci"][0]), int(A["ci"][1]), int(B["ci"][0]), int(B["ci"][1]),
                   sites[i][0], sites[i][3], int(bp.get("var_length", 0) or 0), meta,
                   f_off & 0xFFFFFFFF, f_off >> 32, n_f, s_off & 0xFFFFFFFF, s_off >> 32, n_s)
        f_off += n_f
        s_off += n_s
    batch = ev.EvidenceBatch(rows.astype(np.int32), frags, splits, sample.library_table())
    batch.order = batch.length_order()
    return batch


# ---------------------------------------------------------------------------------------------
# wide -> compact rows and FORMAT text, natively (the numpy / Python forms are their parity checkers)
def compact_from_wide(batch, min_aligned=20, alloc=None, threads=0):
    """compact.compact_from_wide() in C (threaded over sites); `alloc(name, shape, dtype)` may supply the
    destination arrays, e.g. pinned host memory the engine copies from directly."""
    from . import compact as cp
    alloc = alloc or (lambda name, shape, dtype: np.empty(shape, dtype=dtype))
    n = batch.n_sites
    sites = np.ascontiguousarray(batch.sites, dtype=np.int32)
    frags = np.ascontiguousarray(batch.frags, dtype=np.int32)
    splits = np.ascontiguousarray(batch.splits, dtype=np.int32)
    row_off = np.zeros(n + 1, dtype=np.int64)
    counts = np.zeros((max(n, 1), 2), dtype=np.int32)
    _check(lib().svgt_compact_count(sites.ctypes.data, n, frags.ctypes.data, batch.n_frag, splits.ctypes.data, batch.n_split,
                                    int(min_aligned), int(threads), row_off.ctypes.data, counts.ctypes.data))
    out_sites = alloc("sites", (n, cp.CSITE_WORDS), np.int32)
    out_rows = alloc("rows", (int(row_off[n]), cp.CROW_WORDS), np.int32)
    _check(lib().svgt_compact_fill(sites.ctypes.data, n, frags.ctypes.data, batch.n_frag, splits.ctypes.data, batch.n_split,
                                   int(min_aligned), int(threads), row_off.ctypes.data, counts.ctypes.data,
                                   out_sites.ctypes.data, out_rows.ctypes.data if out_rows.size else None))
    out = cp.CompactBatch.__new__(cp.CompactBatch)
    out.sites, out.rows, out.libs, out.order, out.min_aligned = out_sites, out_rows, batch.libs, None, int(min_aligned)
    if batch.order is not None:
        order = alloc("order", (n,), np.int32)
        order[:] = out.length_order()
        out.order = order
    return out


FIELD_IDS = {k: i for i, k in enumerate(("GT", "GQ", "SQ", "GL", "DP", "RO", "AO", "QR", "QA", "RS", "AS", "ASC", "RP", "AP", "AB"))}
STYLE_FULL, STYLE_BLANK, STYLE_DOTS = 0, 1, 2


def format_calls(rows, order, style, threads=0):
    """Sample-column text of scored OUT_DTYPE rows: list of str, one per row (see include/svgt_pack.h)."""
    n = int(rows.shape[0])
    if n == 0:
        return []
    rows = np.ascontiguousarray(rows)
    ids = np.array([FIELD_IDS[k] for k in order], dtype=np.int32)
    style = np.ascontiguousarray(style, dtype=np.uint8)
    stride = 16 * 15 + 64 + 96
    buf = np.empty(n * stride, dtype=np.uint8)
    lens = np.empty(n, dtype=np.int32)
    _check(lib().svgt_format_calls(rows.ctypes.data, n, ids.ctypes.data, len(ids), style.ctypes.data, int(threads),
                                   buf.ctypes.data, stride, lens.ctypes.data))
    raw = buf.tobytes()
    return [raw[i * stride:i * stride + l].decode("ascii") for i, l in enumerate(lens.tolist())]


def host_sq(rows, threads=0):
    """GT / GQ / SQ of scored OUT_DTYPE rows recomputed in place from GL with the host libm (see svgt_pack.h)."""
    if rows.shape[0]:
        assert rows.flags["C_CONTIGUOUS"] and rows.flags["WRITEABLE"]
        _check(lib().svgt_host_sq(rows.ctypes.data, int(rows.shape[0]), int(threads)))
    return rows


def format_quals(qual):
    """'%0.2f' of every value, as a list of str."""
    q = np.ascontiguousarray(qual, dtype=np.float64)
    n = int(q.shape[0])
    if n == 0:
        return []
    stride = 32
    big = np.abs(q) >= 1e20
    if big.any() or not np.isfinite(q).all():
        return ["%0.2f" % v for v in q.tolist()]
    buf = np.empty(n * stride, dtype=np.uint8)
    lens = np.empty(n, dtype=np.int32)
    _check(lib().svgt_format_quals(q.ctypes.data, n, buf.ctypes.data, stride, lens.ctypes.data))
    raw = buf.tobytes()
    return [raw[i * stride:i * stride + l].decode("ascii") for i, l in enumerate(lens.tolist())]
