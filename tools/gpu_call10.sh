#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q --timeout 300 -k "gemm" 2>&1 | tail -8
timeout 300 python tools/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench2.txt
