#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stock_reference.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02d_pytest.log
tail -3 gpurun_out/r02d_pytest.log
for cfg in dtu lego fern_pair; do
  for mode in bins:256 bins:512 bins:1024 default:0; do
    B3GS_BINNING=${mode%%:*} B3GS_BIN_BATCH=${mode##*:} timeout 600 python bench.py --config $cfg --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
    python - <<PY
import json
d=json.load(open("gpurun_out/r02d_bench.json"))
print("${cfg} ${mode}", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], {k:v["ms"] for k,v in d.get("kernels",{}).items() if k in ("depth_sort","binning")})
PY
  done
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 24 -c 12 --csv --log-file gpurun_out/r02d_launches_dtu_bins.csv python tools/gpu_step.py native dtu 6 > gpurun_out/r02d_step.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02d_launches_dtu_bins.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); mi=hdr.index('Metric Name')
for r in rows[1:]: print(r[ki][:50], r[mi], r[vi])
PY
