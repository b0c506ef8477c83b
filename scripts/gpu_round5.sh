#!/bin/bash
# Final 1-GPU call of the round: demo-path tests, whole suite, smoke, bench, then (time permitting) the ncu launch list.
TAG=${1:-r1f}
OUT=gpurun_out
mkdir -p $OUT
( timeout 300 python -m pytest tests/test_gpu_demo.py -m gpu -q ) > $OUT/${TAG}_pytest_demo.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 200 python scripts/kernel_times.py mnist_fashion 1024 4 > $OUT/${TAG}_kernel_times_fashion.txt 2>&1
LADDER_BENCH_PROFILE=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
  --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 > $OUT/${TAG}_ncu_list.log 2>&1
ls -la $OUT | grep $TAG
