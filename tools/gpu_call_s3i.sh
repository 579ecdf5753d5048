#!/bin/bash
timeout 300 python -m pytest tests/test_raster_backward_gpu.py -x -q --tb=short > gpurun_out/s3i.log 2>&1
tail -3 gpurun_out/s3i.log
