#!/usr/bin/env bash
# K8+K9: launch shapes with 64-thread blocks (and the staged-output variant, B3GS_PREBWD_STAGE=1)
mkdir -p gpurun_out
for sh in 6 8 9; do
for cfg in dtu lego; do
  B3GS_PREBWD_SHAPE=$sh timeout 600 python bench.py --config $cfg --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02ai_bench.json 2> gpurun_out/r02ai_bench.err || tail -3 gpurun_out/r02ai_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02ai_bench.json"))
print("shape=$sh $cfg", d["ms_per_step"], d["value"], {k:(v["ms"], v["GBps"]) for k,v in d.get("kernels",{}).items() if k=="preprocess_backward"}, flush=True)
PY
done
done
