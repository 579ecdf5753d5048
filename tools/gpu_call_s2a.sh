#!/bin/bash
# full GPU suite + fc1 timing + bench after the generation-2 GEMM / temporal MMA / attention trimming
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 100 python tools/gemm_bench.py 2>&1 | grep -E "fc1|qkv"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s2a_bench.json 2> gpurun_out/s2a_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2a_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms_eager')}); print(d['e2e'])
for k,v in d['roofline_detail'].items(): print(k, round(v['avg_ms'],4), round(v['tflops'],1))
PY
tail -3 gpurun_out/s2a_bench.err
