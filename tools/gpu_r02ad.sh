#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stock_reference.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02ad_pytest.log
tail -3 gpurun_out/r02ad_pytest.log
for cfg in dtu lego fern_pair; do
  timeout 600 python bench.py --config $cfg --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02ad_bench_$cfg.json 2> gpurun_out/r02ad_bench.err || tail -3 gpurun_out/r02ad_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02ad_bench_$cfg.json"))
print("$cfg", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], {k:v["ms"] for k,v in d.get("kernels",{}).items()}, flush=True)
PY
done
