"""Import the patched reference package (oracle/_ref/svtyper) for tests/bench baselines.

TEST INFRASTRUCTURE ONLY — never imported by the product path.

The reference needs two third-party modules at import time that this image
lacks: `pysam` (reference `svtyper/classic.py:4`, `singlesample.py:12`) and
`cytoolz.itertoolz.partition_all` (`singlesample.py:5`).  When they are not
installed we register stand-ins in `sys.modules` *before* importing the
reference: `pysam` -> a shim over this repo's own BAM reader
(`svtyper_b200.bamio`), `cytoolz.itertoolz` -> a 4-line `partition_all`.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def _install_standins():
    try:
        import pysam  # noqa: F401
    except ImportError:
        repo = os.path.dirname(HERE)
        if repo not in sys.path:
            sys.path.insert(0, repo)
        from svtyper_b200 import bamio
        shim = types.ModuleType("pysam")
        shim.AlignmentFile = bamio.AlignmentFile
        shim.AlignedSegment = bamio.AlignedSegment
        shim.__standin__ = True
        sys.modules["pysam"] = shim
    try:
        from cytoolz.itertoolz import partition_all  # noqa: F401
    except ImportError:
        def partition_all(n, seq):
            seq = list(seq)
            for i in range(0, len(seq), n):
                yield tuple(seq[i:i + n])
        pkg = types.ModuleType("cytoolz")
        sub = types.ModuleType("cytoolz.itertoolz")
        sub.partition_all = partition_all
        pkg.itertoolz = sub
        sys.modules["cytoolz"] = pkg
        sys.modules["cytoolz.itertoolz"] = sub


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "svtyper", "singlesample.py"))


def ensure(build_if_possible: bool = True) -> bool:
    """Make sure oracle/_ref exists (regenerating from /root/reference if present)."""
    if available():
        return True
    if build_if_possible and os.path.isdir("/root/reference/svtyper"):
        from . import make_ref
        make_ref.make_ref()
        return available()
    return False


def load():
    """Return the reference modules as a namespace (classic, singlesample, parsers, ...)."""
    if not available():
        raise ImportError("oracle/_ref not built; run `python oracle/make_ref.py` where "
                          "/root/reference exists")
    _install_standins()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ns = types.SimpleNamespace()
    for name in ("statistics", "parsers", "utils", "classic", "singlesample"):
        setattr(ns, name, importlib.import_module("svtyper." + name))
    return ns
