#!/bin/bash
timeout 300 python -m pytest tests/test_ops_gpu.py -q --timeout 120 -k "attn or attention" 2>&1 | tail -8
timeout 100 python tools/attn_experiments.py 2>&1 | tail -28
