#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py -q --timeout 300 2>&1 | tail -8
timeout 300 python tools/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_e.json 2> gpurun_out/bench_r1_e.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_e.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms_eager')}); print(d['e2e'])
for k,v in d['roofline_detail'].items(): print(k, round(v['avg_ms'],4), round(v['tflops'],1))
PY
tail -5 gpurun_out/bench_r1_e.err
timeout 600 ncu --clock-control none --set full --import-source on -k regex:attn_fwd -s 8 -c 1 -f -o gpurun_out/prof_attn_v2 python tools/profile_kernels.py attn > gpurun_out/ncu_attn_v2.log 2>&1; tail -2 gpurun_out/ncu_attn_v2.log
