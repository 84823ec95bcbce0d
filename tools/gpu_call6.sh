#!/bin/bash
# GPU call 6 (round 2): K1 micro-trims, K2 heavy-first order: parity + A/B.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/r2c6
timeout 900 python -m pytest tests -m gpu -x -q > $O.pytest.log 2>&1; echo "pytest exit $?" >> $O.pytest.log
tail -4 $O.pytest.log
for t in "" "spare0=1"; do
  timeout 300 python tools/class_probe.py --mib 1024 --reps 4 --classes text,mixed --tune "$t" >> $O.class.jsonl 2>> $O.class.err
done
cat $O.class.jsonl; tail -3 $O.class.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline --no-e2e > $O.bench.json 2> $O.bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c6.bench.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['compress_gbs'], d['decompress_gbs'], d['roofline']['kernel_ms'], d['roofline_decompress']['kernel_ms'], d['roofline_decompress']['index_walk_ms'])
PY
