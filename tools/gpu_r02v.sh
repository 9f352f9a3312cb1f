#!/usr/bin/env bash
mkdir -p gpurun_out
for occ in 12 14; do
  B3GS_BWD_OCC=$occ timeout 600 python bench.py --config dtu --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02v.json 2> gpurun_out/r02v.err
  python - "dtu packed occ=$occ" <<'PY'
import json,sys
d=json.load(open("gpurun_out/r02v.json")); k=d["kernels"]
print(sys.argv[1], "| step", d["ms_per_step"], d["value"], "bwd", k["composite_backward"]["ms"])
PY
done
