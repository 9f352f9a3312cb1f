#!/usr/bin/env bash
# round 2, call b: direct tile binning — parity, sanitizer, A/B against the radix path
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02b_pytest.log
tail -4 gpurun_out/r02b_pytest.log
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r02b_sanitize_$tool.log 2>&1; echo "sanitize $tool rc=$?"; tail -2 gpurun_out/r02b_sanitize_$tool.log
done
for cfg in dtu lego fern_pair; do
  for mode in bins radix; do
    B3GS_BINNING=$mode timeout 600 python bench.py --config $cfg --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02b_bench_${cfg}_${mode}.json 2> gpurun_out/r02b_bench_${cfg}_${mode}.err
    python - <<PY
import json
d=json.load(open("gpurun_out/r02b_bench_${cfg}_${mode}.json"))
print("${cfg} ${mode}", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], {k:v["ms"] for k,v in d.get("kernels",{}).items()})
PY
  done
done
