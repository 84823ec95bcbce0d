#!/bin/bash
# ncu --set full of one kernel on one data class: counters, per-line listing, per-SASS listing.
# usage: gpu_ncu_k1.sh <tag> <mode> [kernel-regex]   (MIB=... overrides the 1 GiB default)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/$1; MODE=$2; K=${3:-k_compress_fast}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"$K" -c ${NCU_COUNT:-1} -o $O -f python tools/ncu_target.py --mib ${MIB:-1024} --mode $MODE > $O.ncu.log 2>&1
ncu -i $O.ncu-rep --page raw --csv > $O.raw.csv 2>> $O.ncu.log
ncu -i $O.ncu-rep --page source --csv --print-source cuda,sass > $O.source.csv 2>> $O.ncu.log
ncu -i $O.ncu-rep --page source --csv --print-source sass > $O.sass.csv 2>> $O.ncu.log
python tools/ncu_summary.py $O.raw.csv > $O.summary.jsonl; python tools/ncu_summary.py $O.raw.csv "" pipes > $O.pipes.jsonl
python profiles/ncu_lines.py $O.source.csv 70 > $O.lines.txt 2>&1
python tools/ncu_sass.py $O.sass.csv > $O.sass.txt 2>&1
rm -f $O.ncu-rep $O.source.csv $O.raw.csv $O.sass.csv
cat $O.summary.jsonl
