#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02ah_bench.json 2> gpurun_out/r02ah_bench.err; echo rc=$?; tail -3 gpurun_out/r02ah_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02ah_bench.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "graph", d.get("graph_replay"))
PY
