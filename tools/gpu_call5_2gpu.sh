#!/bin/bash
# GPU call 5 (round 2, 2 GPUs): the NCCL two-rank product-path test, and bench.py under torchrun at N=2.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/r2c5
nvidia-smi -L > $O.smi.txt 2>&1
timeout 900 python -m pytest tests/test_sharded.py -m gpu -x -q -rs > $O.pytest.log 2>&1; echo "pytest exit $?" >> $O.pytest.log
tail -6 $O.pytest.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 3 --warmup 3 --config3-gib 4 --config4-gib 2 > $O.bench_n2.json 2> $O.bench_n2.err; echo "bench n2 exit $?"
tail -c 4000 $O.bench_n2.json; tail -c 1500 $O.bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $O.ref_n2.json 2> $O.ref_n2.err; echo "ref n2 exit $?"; tail -c 600 $O.ref_n2.json
