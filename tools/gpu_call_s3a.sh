#!/bin/bash
timeout 300 python -m pytest tests/test_losses_gpu.py -x -q 2>&1 | tail -15
timeout 200 python -m pytest tests/test_ops_gpu.py -x -q -k "generations" 2>&1 | tail -3
timeout 100 python tools/loss_bench.py 2>&1 | tail -4
