#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:"sort_blend" -s 2 -c 1 -o gpurun_out/blend2_full -f python tools/profile_kernels.py raster > gpurun_out/blend2.log 2>&1
tail -3 gpurun_out/blend2.log
