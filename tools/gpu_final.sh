#!/bin/bash
# Final measurements of a build (round 2): tests, the two bench arms, the ncu launch list of the bench command, one
# `ncu --set full` capture of K1/K2 on the bench workload (-> profiles/traffic.json), per-line listings, a K3 summary.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/${1:-r2final}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem --format=csv > $O.smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O.pytest.log 2>&1; echo "pytest exit $?" >> $O.pytest.log
tail -3 $O.pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O.smoke.log 2>&1; tail -1 $O.smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > $O.ref.json 2> $O.ref.err; echo "ref exit $?"
timeout 1500 python bench.py --steps 5 --warmup 3 > $O.bench.json 2> $O.bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('$O.bench.json').read().strip().split('\n')[-1])
r=json.loads(open('$O.ref.json').read().strip().split('\n')[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'c', d['compress_gbs'], 'd', d['decompress_gbs'], 'e2e', d['e2e']['value'], 'pageable', d['e2e']['pageable']['value'], 'ref', r['value'], 'same_config', d['config']==r['config'])
print('roofline', d['roofline']['frac'], d['roofline_decompress']['frac'], 'traffic', d['roofline']['traffic'], d['roofline']['traffic_note'])
PY
# launch list of the bench command (cold-cache, serialised per-launch times: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O.launches.csv \
   python bench.py --steps 2 --warmup 1 --no-extra --no-e2e --no-cpu-baseline > $O.ncu_launches.log 2>&1
# full capture of the two codec kernels on the bench workload
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_compress_fast|k_decompress" -c 3 -o $O.k1k2 -f \
   python bench.py --steps 1 --warmup 1 --no-extra --no-e2e --no-cpu-baseline > $O.ncu_full.log 2>&1
ncu -i $O.k1k2.ncu-rep --page raw --csv > $O.k1k2_raw.csv 2>> $O.ncu_full.log
ncu -i $O.k1k2.ncu-rep --page source --csv --print-source cuda,sass -k regex:k_compress_fast > $O.k1_source.csv 2>> $O.ncu_full.log
ncu -i $O.k1k2.ncu-rep --page source --csv --print-source cuda,sass -k regex:k_decompress > $O.k2_source.csv 2>> $O.ncu_full.log
python profiles/ncu_lines.py $O.k1_source.csv 50 > $O.k1_mixed_lines.txt 2>&1
python profiles/ncu_lines.py $O.k2_source.csv 50 > $O.k2_mixed_lines.txt 2>&1
python tools/make_traffic.py $O.k1k2_raw.csv $O.traffic.json
rm -f $O.k1k2.ncu-rep $O.k1_source.csv $O.k2_source.csv
# K3: DRAM traffic of HC-9 on 256 MiB of text
timeout 900 ncu --set full --clock-control none -k regex:k_compress_hc -c 1 -o $O.k3 -f python tools/hc_probe.py --mib 256 --mode 0 > $O.ncu_k3.log 2>&1
ncu -i $O.k3.ncu-rep --page raw --csv > $O.k3_raw.csv 2>> $O.ncu_k3.log
rm -f $O.k3.ncu-rep
ls -la gpurun_out | grep ${1:-r2final}
