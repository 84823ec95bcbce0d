#!/bin/bash
# bench.py under torchrun at N GPUs (headline only).  usage: gpu_ngpu.sh <tag> <N>
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/$1; N=$2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --steps 5 --warmup 3 --no-extra > $O.bench_n$N.json 2> $O.bench_n$N.err; echo "bench n$N exit $?"
python - <<PY
import json
d=json.loads(open('$O.bench_n$N.json').read().strip().split('\n')[-1])
print('N=$N value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('pcie_measured_gbs'))
PY
tail -c 400 $O.bench_n$N.err
