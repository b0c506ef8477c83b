#!/bin/bash
# Second GPU call of the round: new parity tests first, then the whole GPU suite, bench (both workloads in one line),
# and one ncu --set full capture of the dominant kernel launched alone.
TAG=${1:-r1c}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 600 python -m pytest tests/test_gpu_mixture.py tests/test_gpu_layers.py tests/test_gpu_layers_tc.py tests/test_gpu_engine.py -m gpu -q ) > $OUT/${TAG}_pytest_new.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_new.log
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
timeout 500 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tma_kernel -s 3 -c 1 -f -o $OUT/${TAG}_dominant_fprop python scripts/roofline_kernel.py > $OUT/${TAG}_ncu_dominant.log 2>&1
timeout 200 python scripts/tc_microbench.py bf16 > $OUT/${TAG}_tc_microbench.jsonl 2>&1
ls -la $OUT
