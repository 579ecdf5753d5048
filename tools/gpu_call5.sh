#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -6
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err; echo "bench exit $?"
tail -c 3500 gpurun_out/bench_r1_b.json; tail -5 gpurun_out/bench_r1_b.err
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:attn_fwd -s 8 -c 4 -f -o gpurun_out/prof_attn python tools/profile_kernels.py attn > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log
timeout 600 $NCU --set full --import-source on -k regex:gemm_kernel -s 4 -c 2 -f -o gpurun_out/prof_gemm python tools/profile_kernels.py gemm > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
timeout 600 $NCU --set full --import-source on -k regex:"sort_blend|preprocess|scatter" -s 3 -c 3 -f -o gpurun_out/prof_raster python tools/profile_kernels.py raster > gpurun_out/ncu_raster.log 2>&1; tail -2 gpurun_out/ncu_raster.log
timeout 900 $NCU --metrics gpu__time_duration.sum -s 40 -c 700 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches_r1.csv
ls -la gpurun_out
