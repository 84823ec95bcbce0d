#!/bin/bash
# GPU call 9 (round 2): K3 jump mode gated by hop distance: parity + timing.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/r2c11
timeout 900 python -m pytest tests/test_gpu_hc.py tests/test_second_source.py tests/test_golden.py tests/test_gpu_frame.py tests/test_gpu_fullsize.py -m gpu -x -q > $O.pytest.log 2>&1; echo "pytest exit $?" >> $O.pytest.log
tail -3 $O.pytest.log
for t in "" "k3_variant=16"; do
  for mode in 0 1 4; do
    timeout 300 python tools/hc_probe.py --mib 1024 --mode $mode --reps 2 --tune "$t" >> $O.k3.jsonl 2>> $O.k3.err
  done
done
timeout 300 python tools/hc_probe.py --mib 4096 --mode 0 --reps 2 >> $O.k3.jsonl 2>> $O.k3.err
timeout 300 python tools/hc_probe.py --mib 4096 --mode 0 --reps 2 --tune "k3_variant=16" >> $O.k3.jsonl 2>> $O.k3.err
cat $O.k3.jsonl; tail -3 $O.k3.err
