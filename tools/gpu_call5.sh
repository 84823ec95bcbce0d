#!/bin/bash
# GPU call 5 (round 2): mover flush fix, K2 wide copy + L1 prefetch; K2 profile.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/r2c5a
timeout 900 python -m pytest tests -m gpu -x -q > $O.pytest.log 2>&1; echo "pytest exit $?" >> $O.pytest.log
tail -4 $O.pytest.log
timeout 300 python tools/class_probe.py --mib 1024 --reps 3 >> $O.class.jsonl 2>> $O.class.err
timeout 300 python tools/class_probe.py --mib 1024 --reps 3 --tune k2_variant=1 >> $O.class.jsonl 2>> $O.class.err
cat $O.class.jsonl; tail -3 $O.class.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_decompress" -c 1 -o $O.k2 -f python tools/ncu_target.py --mib 256 --mode 0 > $O.ncu.log 2>&1
ncu -i $O.k2.ncu-rep --page source --csv --print-source cuda,sass > $O.k2_source.csv 2>> $O.ncu.log
ncu -i $O.k2.ncu-rep --page raw --csv > $O.k2_raw.csv 2>> $O.ncu.log
ncu -i $O.k2.ncu-rep --page details --csv > $O.k2_details.csv 2>> $O.ncu.log
python profiles/ncu_lines.py $O.k2_source.csv 50 > $O.k2_lines.txt 2>&1
rm -f $O.k2.ncu-rep
head -5 $O.k2_lines.txt
