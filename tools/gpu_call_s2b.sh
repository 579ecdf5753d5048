#!/bin/bash
timeout 200 python -m pytest tests/test_raster_gpu.py tests/test_raster_backward_gpu.py tests/test_api_gpu.py -x -q 2>&1 | tail -3
timeout 100 python tools/raster_microbench.py
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv -k regex:"preprocess|scan|scatter|sort_blend" -s 12 -c 8 --log-file gpurun_out/raster_list.csv python tools/raster_microbench.py --iters 2 > /dev/null 2>&1
grep -E "^\"[0-9]" gpurun_out/raster_list.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | tail -8
