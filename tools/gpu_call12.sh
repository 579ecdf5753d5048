#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -5
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_i.json 2> gpurun_out/bench_r1_i.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_i.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms_eager')}); print(d['e2e'])
for k,v in d['roofline_detail'].items(): print(k, round(v['avg_ms'],4), round(v['tflops'],1))
PY
tail -5 gpurun_out/bench_r1_i.err
