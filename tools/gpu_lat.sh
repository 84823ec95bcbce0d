#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/$1
: > $O.lat.jsonl
for mib in 592 1184 2368 4736 8288; do
  timeout 600 python tools/class_probe.py --mib $mib --block-id 7 --reps 2 --classes text >> $O.lat.jsonl 2>> $O.lat.err
done
cat $O.lat.jsonl; tail -3 $O.lat.err
