#!/usr/bin/env bash
# scatter occupancy against staging capacity: 4 (or 5) blocks per SM with fewer staging slots
mkdir -p gpurun_out
run() {  # cfg batch band stage
  B3GS_BIN_BATCH=$2 B3GS_BIN_BAND=$3 B3GS_BIN_STAGE=$4 timeout 600 python bench.py --config $1 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02ag_bench.json 2> gpurun_out/r02ag_bench.err || tail -3 gpurun_out/r02ag_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02ag_bench.json"))
print("$1 batch=$2 band=$3 stage=$4", d["ms_per_step"], {k:v["ms"] for k,v in d.get("kernels",{}).items() if k in ("depth_sort","binning")}, flush=True)
PY
}
for c in "0 0 0" "768 0 5984" "640 0 6368" "512 0 6752" "512 0 4089" "768 0 8000"; do run dtu $c; done
