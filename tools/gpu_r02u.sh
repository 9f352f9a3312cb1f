#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stock_reference.py -m gpu -q -x 2>&1 | tail -4
for packed in 1 0; do
for cfg in dtu lego fern_pair; do
  B3GS_BWD_PACKED=$packed timeout 600 python bench.py --config $cfg --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02u.json 2> gpurun_out/r02u.err
  python - "$cfg packed=$packed" <<'PY'
import json,sys
d=json.load(open("gpurun_out/r02u.json")); k=d["kernels"]
print(sys.argv[1], "| step", d["ms_per_step"], d["value"], "bwd", k["composite_backward"]["ms"], "fwd", k["composite_forward"]["ms"])
PY
done
done
