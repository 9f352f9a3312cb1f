#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stock_reference.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02ac_pytest.log
tail -3 gpurun_out/r02ac_pytest.log
run() {  # cfg batch band stage
  B3GS_BIN_BATCH=$2 B3GS_BIN_BAND=$3 B3GS_BIN_STAGE=$4 timeout 600 python bench.py --config $1 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02ac_bench.json 2> gpurun_out/r02ac_bench.err || tail -3 gpurun_out/r02ac_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02ac_bench.json"))
print("$1 batch=$2 band=$3 stage=$4", d["ms_per_step"], {k:v["ms"] for k,v in d.get("kernels",{}).items() if k in ("depth_sort","binning")}, flush=True)
PY
}
for c in "0 0 0" "1024 0 0"; do run dtu $c; done
for c in "0 0 0" "384 0 0"; do run lego $c; done
for c in "0 0 0"; do run fern_pair $c; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_bins -s 2 -c 2 -f -o gpurun_out/r02ac_bins python tools/gpu_step.py native dtu 3 > gpurun_out/r02ac_ncu.log 2>&1
