#!/usr/bin/env bash
mkdir -p gpurun_out
run() { # env..., label
  env "$@" timeout 600 python bench.py --config dtu --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02p.json 2> gpurun_out/r02p.err
  python - "$*" <<'PY'
import json,sys
d=json.load(open("gpurun_out/r02p.json")); k=d["kernels"]
print(sys.argv[1], "| step", d["ms_per_step"], "fwd", k["composite_forward"]["ms"], "bwd", k["composite_backward"]["ms"])
PY
}
run B3GS_BWD_PIX=4 B3GS_BWD_OCC=8
run B3GS_BWD_PIX=4 B3GS_BWD_OCC=10
run B3GS_BWD_PIX=4 B3GS_BWD_OCC=12
run B3GS_BWD_PIX=4 B3GS_BWD_OCC=14
run B3GS_BWD_PIX=2 B3GS_BWD_OCC=6
run B3GS_BWD_PIX=2 B3GS_BWD_OCC=7
run B3GS_BWD_PIX=2 B3GS_BWD_OCC=8
run B3GS_FWD_OCC=4
run B3GS_FWD_OCC=6
run B3GS_FWD_OCC=3
