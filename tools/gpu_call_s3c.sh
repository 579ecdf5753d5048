#!/bin/bash
timeout 300 python -m pytest tests/test_raster_gpu.py -x -q 2>&1 | tail -5
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/s3c_bench.json 2> gpurun_out/s3c_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/s3c_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["stage_ms_eager"], d["roofline"]["frac"])
PY
tail -3 gpurun_out/s3c_bench.err
