#!/bin/bash
# GPU call 1 (round 2): full GPU test suite on the new K2 front end, A/B timings, per-line ncu listings of K1 and K2.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/r2c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O.smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O.pytest.log 2>&1; echo "pytest exit $?" >> $O.pytest.log
tail -5 $O.pytest.log
for t in "k2_variant=1" "k2_variant=2" "k2_variant=2,k2_occ=8" "k2_variant=2,k2_occ=12"; do
  timeout 300 python tools/class_probe.py --mib 1024 --reps 3 --tune "$t" >> $O.class.jsonl 2>> $O.class.err
done
timeout 300 python tools/class_probe.py --mib 1024 --reps 2 --block-id 7 --classes text,binary,mixed --tune "k2_variant=1" >> $O.class4m.jsonl 2>> $O.class.err
timeout 300 python tools/class_probe.py --mib 1024 --reps 2 --block-id 7 --classes text,binary,mixed --tune "k2_variant=2" >> $O.class4m.jsonl 2>> $O.class.err
cat $O.class.jsonl $O.class4m.jsonl
timeout 600 python bench.py --steps 5 --warmup 3 > $O.bench.json 2> $O.bench.err; tail -c 1500 $O.bench.json
# per-line listings (text class, 256 MiB): SourceCounters + stall sampling
for mode in 0; do
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_compress_fast|k_decompress" -c 2 \
     -o $O.text_k1k2 -f python tools/ncu_target.py --mib 256 --mode $mode > $O.ncu.log 2>&1
  ncu -i $O.text_k1k2.ncu-rep --page source --csv --print-source cuda,sass -k regex:k_compress_fast > $O.k1_source.csv 2>> $O.ncu.log
  ncu -i $O.text_k1k2.ncu-rep --page source --csv --print-source cuda,sass -k regex:k_decompress > $O.k2_source.csv 2>> $O.ncu.log
  ncu -i $O.text_k1k2.ncu-rep --page raw --csv > $O.k1k2_raw.csv 2>> $O.ncu.log
  python profiles/ncu_lines.py $O.k1_source.csv 60 > $O.k1_lines.txt 2>&1
  python profiles/ncu_lines.py $O.k2_source.csv 60 > $O.k2_lines.txt 2>&1
done
ls -la gpurun_out | tail -20
