#!/bin/bash
TAG=${1:-r1d}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -q ) > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
timeout 500 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 200 python scripts/kernel_times.py mnist_fashion 1024 4 > $OUT/${TAG}_kernel_times_fashion.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tma_kernel -s 3 -c 1 -f -o $OUT/${TAG}_dominant_fprop python scripts/roofline_kernel.py > $OUT/${TAG}_ncu_dominant.log 2>&1
ls -la $OUT | grep $TAG
