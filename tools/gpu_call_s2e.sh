#!/bin/bash
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/s2e_bench.json 2> gpurun_out/s2e_bench.err; echo "bench exit $?"
ncu --profile-from-start off --cache-control none --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/nfe.csv python tools/nfe_breakdown.py > gpurun_out/nfe.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"attn_fwd6|attn_merge" -c 2 -o gpurun_out/attn6_full -f python tools/profile_kernels.py attn > gpurun_out/p1.log 2>&1
timeout 100 python tools/gemm_bench.py > gpurun_out/gemm_bench.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
