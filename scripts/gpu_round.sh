#!/bin/bash
# One gpurun call: GPU parity tests, bench lines (both arms), ncu launch list + full capture of the top kernel.
# Usage (from the repo root on the GPU box): bash scripts/gpu_round.sh [tag]
TAG=${1:-r1b}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench_fashion.json 2> $OUT/${TAG}_bench_fashion.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
timeout 400 python bench.py --workload celeba --batch 64 --steps 10 --warmup 3 --cpu-sample 4 > $OUT/${TAG}_bench_celeba.json 2> $OUT/${TAG}_bench_celeba.err
# launch list of the timed region only (2 iterations; cold-cache and serialised: compare SHARES)
LADDER_BENCH_PROFILE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
  --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 > $OUT/${TAG}_ncu_list.log 2>&1
LADDER_BENCH_PROFILE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
  --csv --log-file $OUT/${TAG}_launches_celeba.csv python bench.py --workload celeba --batch 64 --steps 1 --warmup 3 > $OUT/${TAG}_ncu_list_celeba.log 2>&1
# full capture of the dominant kernel family (3 launches of the TMA-fed tcgen05 conv kernel inside the timed region)
LADDER_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:tma_kernel -c 6 -f -o $OUT/${TAG}_tma_kernel python bench.py --steps 1 --warmup 3 > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT
timeout 200 python scripts/tc_microbench.py bf16 > $OUT/${TAG}_tc_microbench.jsonl 2>&1
timeout 300 python scripts/mix_microbench.py > $OUT/${TAG}_mix_microbench.json 2>&1
LADDER_BENCH_PROFILE=1 timeout 400 ncu --set full --clock-control none --import-source on \
  -k regex:mix -c 4 -f -o $OUT/${TAG}_mix_kernel python scripts/mix_microbench.py > $OUT/${TAG}_ncu_mix.log 2>&1
ls -la $OUT
