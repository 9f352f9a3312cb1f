#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stock_reference.py -m gpu -q -x 2>&1 | tail -3
for mode in 0 1; do
for cfg in dtu lego fern_pair; do
B3GS_BIN_STAGED=$mode timeout 600 python bench.py --config $cfg --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02n_bench.json"))
print("$cfg staged=$mode", d["ms_per_step"], d["value"], {k:v["ms"] for k,v in d.get("kernels",{}).items() if k in ("binning","depth_sort")})
PY
done
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tile_bins -s 8 -c 4 --csv --log-file gpurun_out/r02n_bins.csv python tools/gpu_step.py native dtu 6 > /dev/null 2>&1
grep tile_bins gpurun_out/r02n_bins.csv | awk -F'","' '{print $5, $NF}'
