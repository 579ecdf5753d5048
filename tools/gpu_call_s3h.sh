#!/bin/bash
timeout 300 python -m pytest tests/test_api_gpu.py -x -q 2>&1 | tail -3
for f in "" "--no-prefetch"; do
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $f > gpurun_out/s3h_bench$f.json 2> gpurun_out/s3h_bench$f.err; echo "bench exit $?"
python - "$f" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/s3h_bench{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1] or "prefetch", d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"].get("max_abs_diff_vs_resident_rgba"))
PY
done
