#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_dp_gloo.py 2>&1 | tail -25 > gpurun_out/r02k_pytest.log
tail -3 gpurun_out/r02k_pytest.log
for occ in 3; do
for cfg in dtu lego; do
B3GS_PREBWD_OCC=$occ timeout 600 python bench.py --config $cfg --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02k_bench.json"))
print("$cfg occ=$occ", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], {k:v["ms"] for k,v in d.get("kernels",{}).items()}, d["roofline"].get("issue_slot_frac"))
PY
done
done
