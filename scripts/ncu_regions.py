#!/usr/bin/env python
"""Segment the SASS of an ncu source-page CSV into runs of equal execution count (basic-block groups)
and print each run's share of executed warp-instructions, stall samples and shared wavefronts."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
h = rows[1]; d = rows[2:]
ie = h.index('Instructions Executed'); ism = h.index('# Samples'); iw = h.index('L1 Wavefronts Shared')
tot = sum(int(r[ie]) for r in d); wf = sum(int(r[iw]) for r in d if r[iw].isdigit()); sm = sum(int(r[ism]) for r in d)
print('total inst', tot, 'shared wavefronts', wf, 'samples', sm)
i = 0
while i < len(d):
    e = int(d[i][ie]); j = i; s = 0; q = 0; w = 0
    while j < len(d) and abs(int(d[j][ie]) - e) <= max(1, 0.02 * e):
        s += int(d[j][ie]); q += int(d[j][ism]); w += int(d[j][iw]) if d[j][iw].isdigit() else 0; j += 1
    if s > thr * tot:
        print('lines %4d-%4d (%3d) exec~%9d  sum %10d (%4.1f%%) samples %6d (%4.1f%%) wavefronts %9d' % (i, j - 1, j - i, e, s, 100 * s / tot, q, 100.0 * q / sm, w))
    i = j
