#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -6
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_c.json 2> gpurun_out/bench_r1_c.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_c.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms_eager')}); print(d['e2e']); 
for k,v in d['roofline_detail'].items(): print(k, round(v['avg_ms'],4), round(v['tflops'],1))
PY
tail -5 gpurun_out/bench_r1_c.err
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -s 40 -c 500 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
