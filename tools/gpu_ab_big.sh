#!/bin/bash
# A/B of 4 MiB-block runs.  usage: gpu_ab_big.sh <tag> <mib> <classes> <tune...>
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/$1; MIB=$2; CL=$3; shift; shift; shift
: > $O.big.jsonl
for t in "$@"; do
  [ "$t" = "-" ] && t=""
  timeout 600 python tools/class_probe.py --mib $MIB --block-id 7 --reps 2 --classes $CL --tune "$t" >> $O.big.jsonl 2>> $O.big.err
done
cat $O.big.jsonl; tail -3 $O.big.err
