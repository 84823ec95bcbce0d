#!/usr/bin/env python
"""Per-SASS-instruction listing (program order) from an `ncu --page source --csv --print-source sass` export:
executed warp-instructions (millions), stall samples, instruction text.  usage: ncu_sass.py export.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] in ("Address", "#") or (hdr is None and "Source" in r and "Instructions Executed" in r):
        hdr = r
        continue
    if hdr is None:
        continue
    try:
        ie = int(r[hdr.index("Instructions Executed")])
        ss = int(r[hdr.index("Warp Stall Sampling (All Samples)")])
    except (ValueError, IndexError):
        continue
    print("%9.2f %7d  %s" % (ie / 1e6, ss, r[hdr.index("Source")].strip()[:110]))
