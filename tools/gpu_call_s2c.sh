#!/bin/bash
# round-1 profile set: per-NFE launch list + full captures of the three headline kernels
ncu --profile-from-start off --cache-control none --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/nfe.csv python tools/nfe_breakdown.py > gpurun_out/nfe.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_fwd6 -c 1 -o gpurun_out/attn6_full -f python tools/profile_kernels.py attn > gpurun_out/p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_ws -s 4 -c 2 -o gpurun_out/gemm_ws_full -f python tools/profile_kernels.py gemm > gpurun_out/p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sort_blend|scatter|preprocess" -s 4 -c 3 -o gpurun_out/raster_full -f python tools/profile_kernels.py raster > gpurun_out/p3.log 2>&1
ncu --set full --clock-control none -k regex:"attn_small_mma|ln_mod" -s 2 -c 2 -o gpurun_out/small_full -f python tools/nfe_breakdown.py nfe > gpurun_out/p4.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/p*.log
