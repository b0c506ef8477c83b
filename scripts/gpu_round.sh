#!/bin/bash
# One gpurun call on a B200 box.  Usage (repo root):  bash scripts/gpu_round.sh <tag> [quick|full|dp]
#   quick : whole GPU test suite, smoke, bench (both CelebA legs inside), CUPTI per-kernel times          (~4 min)
#   full  : quick + ncu launch list of the timed region + ncu --set full of the step's leading kernels     (~15 min)
#   dp    : on a 2-GPU box -- GPU suite (NCCL tests run), bench at N=1 and N=2 (captured and eager NCCL)   (~7 min)
TAG=${1:-rX}
MODE=${2:-quick}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q ) > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
if [ "$MODE" = "dp" ]; then
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err
  LADDER_DP_GRAPHS=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --celeba-batch 0 > $OUT/${TAG}_bench_n2_eager.json 2> $OUT/${TAG}_bench_n2_eager.err
else
  timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
  timeout 200 python scripts/kernel_times.py mnist_fashion 1024 4 > $OUT/${TAG}_kernel_times_fashion.txt 2>&1
  timeout 200 python scripts/kernel_times.py celeba 64 4 > $OUT/${TAG}_kernel_times_celeba.txt 2>&1
fi
if [ "$MODE" = "full" ]; then
  # launch list of the timed region only (1 iteration; cold-cache and serialised: compare SHARES)
  LADDER_BENCH_PROFILE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
    --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 > $OUT/${TAG}_ncu_list.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:tma_kernel -s 3 -c 1 -f \
    -o $OUT/${TAG}_dominant_fprop python scripts/roofline_kernel.py > $OUT/${TAG}_ncu_dominant.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:tma_kernel|thin_k|tap_' -s 12 -c 8 -f \
    -o $OUT/${TAG}_targets python scripts/ncu_targets.py > $OUT/${TAG}_ncu_targets.log 2>&1
  timeout 300 ncu --set full --clock-control none -k regex:mix -c 4 -f -o $OUT/${TAG}_mix_kernel \
    python scripts/mix_microbench.py > $OUT/${TAG}_ncu_mix.log 2>&1
fi
ls -la $OUT | grep ${TAG}
