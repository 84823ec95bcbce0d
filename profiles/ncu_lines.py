#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line.
usage: ncu_lines.py export.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
lines = []
hdr = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] == 'Function Name': continue
    if hdr and r[0].isdigit():
        try:
            ie = int(r[hdr.index('Instructions Executed')])
            ss = int(r[hdr.index('Warp Stall Sampling (All Samples)')])
        except ValueError:
            continue
        lines.append((cur_file, int(r[0]), r[1].strip(), ie, ss))
ti = sum(l[3] for l in lines); ts = sum(l[4] for l in lines)
print("total inst %d  total samples %d" % (ti, ts))
print("--- by instructions executed")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    print("%5.1f%% inst %5.1f%% stall  %s:%d  %s" % (100.0*l[3]/ti, 100.0*l[4]/max(ts,1), l[0], l[1], l[2][:90]))
print("--- by stall samples")
for l in sorted(lines, key=lambda l: -l[4])[:top]:
    print("%5.1f%% stall %5.1f%% inst  %s:%d  %s" % (100.0*l[4]/max(ts,1), 100.0*l[3]/ti, l[0], l[1], l[2][:90]))
