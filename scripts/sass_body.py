#!/usr/bin/env python
"""Offline: compile svgt_coop.cu with -DSVGT_MARK (MEMBAR markers around the phase-A body) and
count SASS instructions per pipe class inside the marked region(s) of svgt_tally_kernel<8,0>."""
import collections, os, re, subprocess, sys, tempfile
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(HERE, "svtyper_b200", "csrc", "svgt_coop.cu")
out = os.path.join(tempfile.gettempdir(), "svgt_mark.cubin")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false", "-std=c++17",
                       "-DSVGT_MARK", "-cubin", "-o", out, src])
sass = subprocess.run(["cuobjdump", "-sass", out], capture_output=True, text=True).stdout
ALU = {"ISETP", "LOP3", "SEL", "SHF", "PLOP3", "PRMT", "FSEL", "VIMNMX", "LEA", "IADD3", "VIADD", "FSETP", "ISCADD",
       "POPC", "FLO", "BREV", "IABS", "VOTE", "R2P", "P2R", "FMNMX", "IMNMX", "DSETP", "MOV", "CS2R", "S2R"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2"}
cur = None; regions = []; inreg = False; buf = []
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); inreg = False; continue
    if cur is None or "svgt_tally_kernelILi8ELi0E" not in cur:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(2).split(".")[0]
    if op == "MEMBAR":
        if inreg:
            regions.append(buf); buf = []
        inreg = not inreg
        continue
    if inreg:
        buf.append(op)
for i, r in enumerate(regions):
    c = collections.Counter(r)
    alu = sum(v for k, v in c.items() if k in ALU); fma = sum(v for k, v in c.items() if k in FMA)
    print("region %d: %d instrs, ALU-pipe %d, FMA-pipe %d, other %d" % (i, len(r), alu, fma, len(r) - alu - fma))
    print("   ", ", ".join("%s %d" % kv for kv in c.most_common(18)))
