#!/usr/bin/env bash
mkdir -p gpurun_out
for shape in 0 1 2 3 4; do
for cfg in dtu lego; do
  B3GS_PREBWD_SHAPE=$shape timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02q.json 2> gpurun_out/r02q.err
  python - "$cfg shape=$shape" <<'PY'
import json,sys
d=json.load(open("gpurun_out/r02q.json")); k=d["kernels"]
print(sys.argv[1], "| step", d["ms_per_step"], "K8+K9", k["preprocess_backward"]["ms"], k["preprocess_backward"]["GBps"])
PY
done
done
