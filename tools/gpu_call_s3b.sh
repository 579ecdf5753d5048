#!/bin/bash
timeout 300 python -m pytest tests/test_losses_gpu.py tests/test_raster_backward_gpu.py tests/test_api_gpu.py -x -q 2>&1 | tail -15
timeout 100 python tools/loss_bench.py 2>&1 | tail -4
