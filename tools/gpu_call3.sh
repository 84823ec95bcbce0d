#!/bin/bash
# GPU call 3 (round 2): K1 ring fix, K2 adaptive fallback + occupancy, bench stream fix.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/r2c3
B2_TEST_TUNE=k1_variant=2 timeout 900 python -m pytest tests/test_gpu_block.py tests/test_gpu_frame.py tests/test_golden.py tests/test_second_source.py tests/test_gpu_fullsize.py -m gpu -x -q > $O.pytest_ring.log 2>&1; echo "pytest ring exit $?" >> $O.pytest_ring.log
tail -6 $O.pytest_ring.log
timeout 900 python -m pytest tests -m gpu -x -q > $O.pytest.log 2>&1; echo "pytest exit $?" >> $O.pytest.log
tail -4 $O.pytest.log
for t in "" "k1_variant=2" "k2_occ=9"; do
  timeout 300 python tools/class_probe.py --mib 1024 --reps 3 --tune "$t" >> $O.class.jsonl 2>> $O.class.err
done
for t in "" "k1_variant=2"; do
  timeout 300 python tools/class_probe.py --mib 1024 --reps 2 --block-id 7 --classes text,binary,mixed --tune "$t" >> $O.class4m.jsonl 2>> $O.class.err
done
cat $O.class.jsonl $O.class4m.jsonl
tail -5 $O.class.err
timeout 1200 python bench.py --steps 2 --warmup 3 --no-e2e --config3-gib 2 --config4-gib 1 > $O.bench.json 2> $O.bench.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c3.bench.json').read().strip().split('\n')[-1])
print(d['value'], d['compress_gbs'], d['decompress_gbs'])
print(json.dumps(d.get('config5',{}).get('records_4KiB')))
print(json.dumps(d.get('config3',{}).get('block_checksums')))
PY
tail -c 800 $O.bench.err
