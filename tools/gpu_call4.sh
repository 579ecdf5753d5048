#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py -q --timeout 300 -x 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench_r1_a.json; tail -5 gpurun_out/bench_r1_a.err
