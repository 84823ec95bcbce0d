#!/bin/bash
# quick K1/K2 class timing of the current build (+ optional block parity).  usage: gpu_probe_class.sh <tag> [pytest|nopytest] [extra class_probe args]
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/$1; shift
if [ "$1" = "pytest" ]; then
  timeout 900 python -m pytest tests/test_gpu_block.py tests/test_second_source.py tests/test_golden.py tests/test_gpu_frame.py -m gpu -x -q > $O.pytest.log 2>&1; echo "pytest exit $?" >> $O.pytest.log
  tail -3 $O.pytest.log
fi
shift
timeout 300 python tools/class_probe.py --mib 1024 --reps 3 "$@" > $O.class.jsonl 2> $O.class.err
cat $O.class.jsonl; tail -3 $O.class.err
