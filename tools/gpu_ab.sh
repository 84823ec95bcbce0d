#!/bin/bash
# A/B class timing with tune strings.  usage: gpu_ab.sh <tag> <classes> <tune1> [tune2 ...]   ("-" = no tune)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/$1; CL=$2; shift; shift
: > $O.ab.jsonl
for t in "$@"; do
  [ "$t" = "-" ] && t=""
  timeout 300 python tools/class_probe.py --mib ${MIB:-1024} --reps 3 --classes $CL --tune "$t" >> $O.ab.jsonl 2>> $O.ab.err
done
cat $O.ab.jsonl; tail -3 $O.ab.err
