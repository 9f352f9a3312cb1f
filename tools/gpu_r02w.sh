#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stock_reference.py -m gpu -q -x 2>&1 | tail -3
run() { env "$@" timeout 600 python bench.py --config $CFG --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02w.json 2> gpurun_out/r02w.err
  python - "$CFG $*" <<'PY'
import json,sys
d=json.load(open("gpurun_out/r02w.json")); k=d["kernels"]
print(sys.argv[1], "| step", d["ms_per_step"], d["value"], "bwd", k["composite_backward"]["ms"])
PY
}
for CFG in lego fern_pair; do
run B3GS_BWD_PACKED=0
run B3GS_BWD_PACKED=1 B3GS_BWD_OCC=6
run B3GS_BWD_PACKED=1 B3GS_BWD_OCC=7
run B3GS_BWD_PACKED=1 B3GS_BWD_OCC=8
done
