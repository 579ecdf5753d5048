#!/bin/bash
# gpurun call 1: raster parity + microbench + tcgen05 probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/raster_microbench.py > gpurun_out/raster_bench.log 2>&1; tail -2 gpurun_out/raster_bench.log
timeout 300 python tools/raster_microbench.py --frames 96 > gpurun_out/raster_bench96.log 2>&1; tail -1 gpurun_out/raster_bench96.log
: > gpurun_out/probe.log
for t in 0 1 2 3 4 5; do for v in 0 1 2 3 4; do
  timeout 20 gvfdiffusion_b200/csrc/probe/tc_probe $t $v >> gpurun_out/probe.log 2>&1 || echo "test $t variant $v: exit $?" >> gpurun_out/probe.log
done; done
cat gpurun_out/probe.log
