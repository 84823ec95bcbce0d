#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/$1
: > $O.lat.jsonl
for mib in 592 2368 4144; do
  for t in "" "k1_variant=4"; do
    timeout 600 python tools/class_probe.py --mib $mib --block-id 7 --reps 2 --classes text --tune "$t" >> $O.lat.jsonl 2>> $O.lat.err
  done
done
cat $O.lat.jsonl; tail -3 $O.lat.err
