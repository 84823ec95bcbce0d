#!/bin/bash
# GPU call 2 (round 2): K2 speculative front end + K1 TMA ring variant: parity suites, A/B timings, full bench line.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/r2c2
timeout 900 python -m pytest tests -m gpu -x -q > $O.pytest.log 2>&1; echo "pytest exit $?" >> $O.pytest.log
tail -4 $O.pytest.log
B2_TEST_TUNE=k1_variant=2 timeout 900 python -m pytest tests/test_gpu_block.py tests/test_gpu_frame.py tests/test_golden.py tests/test_second_source.py tests/test_gpu_fullsize.py -m gpu -x -q > $O.pytest_ring.log 2>&1; echo "pytest ring exit $?" >> $O.pytest_ring.log
tail -4 $O.pytest_ring.log
for t in "" "k1_variant=2" "k2_occ=10" "k2_variant=1"; do
  timeout 300 python tools/class_probe.py --mib 1024 --reps 3 --tune "$t" >> $O.class.jsonl 2>> $O.class.err
done
for t in "" "k1_variant=2"; do
  timeout 300 python tools/class_probe.py --mib 1024 --reps 2 --block-id 7 --classes text,binary,mixed --tune "$t" >> $O.class4m.jsonl 2>> $O.class.err
done
cat $O.class.jsonl $O.class4m.jsonl
tail -5 $O.class.err
timeout 1200 python bench.py --steps 5 --warmup 3 > $O.bench.json 2> $O.bench.err; echo "bench exit $?"; tail -c 3000 $O.bench.json; tail -c 1500 $O.bench.err
