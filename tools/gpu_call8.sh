#!/bin/bash
# GPU call 8 (round 2): K3 adaptive jump mode + 16-byte candidate evaluation: parity and threshold sweep.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/r2c8
timeout 900 python -m pytest tests/test_gpu_hc.py tests/test_second_source.py tests/test_golden.py tests/test_gpu_frame.py -m gpu -x -q > $O.pytest.log 2>&1; echo "pytest exit $?" >> $O.pytest.log
tail -5 $O.pytest.log
for t in "spare5=8" "spare5=16" "" "spare5=48" "spare5=1000000" "k3_variant=16"; do
  for mode in 0 1; do
    timeout 300 python tools/hc_probe.py --mib 1024 --mode $mode --reps 2 --tune "$t" >> $O.k3.jsonl 2>> $O.k3.err
  done
done
for t in "" "k3_variant=16"; do
  timeout 300 python tools/hc_probe.py --mib 1024 --mode 0 --level 6 --reps 2 --tune "$t" >> $O.k3.jsonl 2>> $O.k3.err
  timeout 300 python tools/hc_probe.py --mib 1024 --mode 0 --level 8 --reps 2 --tune "$t" >> $O.k3.jsonl 2>> $O.k3.err
done
cat $O.k3.jsonl; tail -3 $O.k3.err
