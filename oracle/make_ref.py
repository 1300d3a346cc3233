#!/usr/bin/env python
"""Generate an importable (Python 3) copy of the reference `svtyper` package.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path imports this.

The reference (hall-lab/svtyper v0.7.1) is Python 2.7.  This container only has
Python 3.12, no lib2to3 and no pysam/cytoolz, so the reference cannot be
imported as shipped.  This script reads the sources *where they lie* under
`/root/reference/svtyper` and writes a mechanically patched copy into the
git-ignored `oracle/_ref/svtyper/` (never committed; see .gitignore).  The
transformations are purely syntactic py2 -> py3 rewrites plus the one py2
semantic that py3 turned into an exception (`None > 0.5` is `False` in py2,
reference `svtyper/parsers.py:882`):

  * `print x, y`            -> `print(x, y)`       (classic.py debug prints)
  * `xrange(`               -> `range(`
  * `except E, e:`          -> `except E as e:`
  * `lambda(x):`            -> `lambda x:`
  * `map(int, ...)` results that are indexed -> `list(map(...))`
    (`confidence_interval`, parsers.py:11-15)
  * `return p > 0.5`        -> `return (p is not None) and (p > 0.5)`

`pysam` and `cytoolz` are supplied at import time by `oracle/ref_loader.py`
(a stand-in module built on this repo's own BAM reader, and a 4-line
`partition_all`).  The result reproduces the reference's golden file
`tests/data/example.gt.vcf` byte-for-byte (see tests/test_oracle_ref.py).
"""
from __future__ import annotations

import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_SRC = "/root/reference/svtyper"
DEFAULT_DST = os.path.join(HERE, "_ref", "svtyper")

_PRINT_STMT = re.compile(r"^(\s*)print (?!\()(.*)$")
_EXCEPT_COMMA = re.compile(r"except\s+(\w+)\s*,\s*(\w+)\s*:")


def _patch_source(name: str, text: str) -> str:
    out = []
    for line in text.split("\n"):
        m = _PRINT_STMT.match(line)
        if m and "print_function" not in text:
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        line = line.replace("xrange(", "range(")
        line = _EXCEPT_COMMA.sub(r"except \1 as \2:", line)
        line = line.replace("lambda(x):", "lambda x:")
        out.append(line)
    text = "\n".join(out)
    if name == "parsers.py":
        # confidence_interval indexes the result of map() (py2 list)
        text = text.replace(
            "ci = map(int, var.info[tag].split(','))",
            "ci = list(map(int, var.info[tag].split(',')))")
        text = text.replace(
            "return map(int, var.info[alt_tag].split(','))",
            "return list(map(int, var.info[alt_tag].split(',')))")
        # py2: None > 0.5 is False (ZeroDivisionError branch of p_concordant)
        assert "return p > 0.5" in text
        text = text.replace("return p > 0.5",
                            "return (p is not None) and (p > 0.5)")
    return text


def make_ref(src: str = DEFAULT_SRC, dst: str = DEFAULT_DST) -> str:
    """Write the patched package; returns the directory to put on sys.path."""
    if not os.path.isdir(src):
        raise FileNotFoundError(src)
    os.makedirs(dst, exist_ok=True)
    for fn in sorted(os.listdir(src)):
        if not fn.endswith(".py"):
            continue
        with open(os.path.join(src, fn), "r") as f:
            text = f.read()
        patched = _patch_source(fn, text)
        compile(patched, fn, "exec")  # must be valid py3
        with open(os.path.join(dst, fn), "w") as f:
            f.write(patched)
    return os.path.dirname(dst)


def have_ref(dst: str = DEFAULT_DST) -> bool:
    return os.path.isfile(os.path.join(dst, "singlesample.py"))


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else DEFAULT_SRC
    print(make_ref(src))
