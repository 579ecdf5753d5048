#!/bin/bash
timeout 300 python -m pytest tests/test_raster_gpu.py tests/test_api_gpu.py -x -q 2>&1 | tail -4
timeout 200 python tools/raster_microbench.py 2>&1 | tail -2
