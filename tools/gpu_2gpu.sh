#!/bin/bash
# 2 GPUs: the NCCL two-rank product-path test and bench.py under torchrun at N=2 (usage: gpu_2gpu.sh <tag>)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/${1:-n2}
nvidia-smi -L > $O.smi.txt 2>&1
timeout 900 python -m pytest tests/test_sharded.py -m gpu -x -q -rs > $O.pytest.log 2>&1; echo "pytest exit $?" >> $O.pytest.log
tail -4 $O.pytest.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 3 --warmup 3 --config3-gib 4 --config4-gib 2 > $O.bench_n2.json 2> $O.bench_n2.err; echo "bench n2 exit $?"
python - <<PY
import json
d=json.loads(open('$O.bench_n2.json').read().strip().split('\n')[-1])
print('N=2 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'c3', {k:(v['compress_gbs'],v['decompress_gbs']) for k,v in d['config3'].items() if isinstance(v,dict) and 'compress_gbs' in v})
PY
tail -c 600 $O.bench_n2.err
