#!/bin/bash
# 2-GPU box: whole GPU suite (the NCCL tests run here), bench at N=1 and N=2 (graphs with captured NCCL, and eager).
TAG=${1:-r1e}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q ) > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err
LADDER_DP_GRAPHS=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --celeba-batch 0 > $OUT/${TAG}_bench_n2_eager.json 2> $OUT/${TAG}_bench_n2_eager.err
ls -la $OUT | grep $TAG
