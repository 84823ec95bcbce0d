#!/bin/bash
# GPU call 4 (round 2): K2 per-block front-end choice; pageable-buffer mover (e2e); quick ring profile.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/r2c4
timeout 900 python -m pytest tests -m gpu -x -q > $O.pytest.log 2>&1; echo "pytest exit $?" >> $O.pytest.log
tail -4 $O.pytest.log
timeout 300 python tools/class_probe.py --mib 1024 --reps 3 >> $O.class.jsonl 2>> $O.class.err
cat $O.class.jsonl; tail -3 $O.class.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > $O.bench.json 2> $O.bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c4.bench.json').read().strip().split('\n')[-1])
print(d['value'], d['compress_gbs'], d['decompress_gbs']); print(json.dumps(d['e2e']))
PY
tail -c 600 $O.bench.err
# pageable round trip correctness at odd sizes through the Python API (numpy buffers are pageable)
timeout 300 python - <<'PY' > $O.pageable.log 2>&1
import sys, time, numpy as np
sys.path.insert(0, '.')
import zig_lz4_b200 as z
from zig_lz4_b200 import datagen
ctx = z.Context(0)
for n in ((700 << 20) + 12345, (300 << 20) + 1, 5 << 20):
    for kw in (dict(blockSizeID=4, blockMode=1), dict(blockSizeID=4, blockMode=1, blockChecksumFlag=1, contentChecksumFlag=1), dict(blockSizeID=7, blockMode=1)):
        zp = z.lz4f.Preferences(**kw)
        src = datagen.generate(n, mode=4)
        cap = z.lz4f.compressFrameBound(n, zp)
        dst = np.empty(cap, dtype=np.uint8); back = np.empty(n, dtype=np.uint8)
        t0 = time.perf_counter(); cs = ctx.compress_frame(src, zp, dst=dst); t1 = time.perf_counter()
        m = ctx.decompress_frame(dst[:cs], dst=back); t2 = time.perf_counter()
        assert m == n and (back == src).all(), (n, kw)
        print(n, kw, "compress %.1f ms decompress %.1f ms -> %.2f GB/s" % ((t1-t0)*1e3, (t2-t1)*1e3, n/(t2-t0)/1e9), flush=True)
print("pageable ok")
PY
tail -12 $O.pageable.log
# one ncu pass over the ring variant of K1 (text, 256 MiB) for the record
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_compress_fast" -c 1 -o $O.k1ring -f python tools/ncu_target.py --mib 256 --mode 0 --tune k1_variant=2 > $O.ncu.log 2>&1
ncu -i $O.k1ring.ncu-rep --page source --csv --print-source cuda,sass > $O.k1ring_source.csv 2>> $O.ncu.log
ncu -i $O.k1ring.ncu-rep --page raw --csv > $O.k1ring_raw.csv 2>> $O.ncu.log
python profiles/ncu_lines.py $O.k1ring_source.csv 40 > $O.k1ring_lines.txt 2>&1
head -30 $O.k1ring_lines.txt
rm -f $O.k1ring.ncu-rep
# K3 occupancy curve (CTAs of 4 warps per SM capped): does keeping the tables inside L2 pay?
for t in "k3_variant=1" "k3_variant=2" "k3_variant=4" "k3_variant=8" ""; do
  timeout 300 python tools/hc_probe.py --mib 1024 --mode 0 --reps 2 --tune "$t" >> $O.k3curve.jsonl 2>> $O.k3.err
done
cat $O.k3curve.jsonl
