#!/bin/bash
# round-1 closing data set: full GPU suite, bench line, per-NFE ncu launch list, rasteriser ncu --set full, smoke
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/s3f_bench.json 2> gpurun_out/s3f_bench.err; echo "bench exit $?"
ncu --profile-from-start off --cache-control none --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/nfe.csv python tools/nfe_breakdown.py > gpurun_out/nfe.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sort_blend|scatter|preprocess" -s 4 -c 3 -o gpurun_out/raster_full -f python tools/profile_kernels.py raster > gpurun_out/p3.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/s3f_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["stage_ms_eager"], d["roofline"]["frac"], d["roofline_raster"]["frac"], d["clocks"])
PY
