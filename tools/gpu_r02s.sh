#!/usr/bin/env bash
# final validation of the round-2 tree on one B200: tests, smoke, sanitizer, both bench arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02s_pytest.log; tail -2 gpurun_out/r02s_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r02s_sanitize_$tool.log 2>&1; echo "sanitize $tool rc=$?"; tail -1 gpurun_out/r02s_sanitize_$tool.log
done
timeout 900 python bench.py --impl reference > gpurun_out/r02s_bench_reference.json 2> gpurun_out/r02s_bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r02s_bench_native.json 2> gpurun_out/r02s_bench_native.err; echo "nat rc=$?"; tail -2 gpurun_out/r02s_bench_native.err
python - <<'PY'
import json
for f in ("r02s_bench_reference","r02s_bench_native"):
    d=json.load(open("gpurun_out/%s.json"%f))
    print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"], "launches", d.get("gpu_launches"))
    print("  configs", {k:(v.get("value"),v.get("ms_per_step")) for k,v in (d.get("configs") or {}).items()})
    if "kernels" in d: print("  kernels", {k:(v["ms"],v["GBps"],v.get("issue_slot_frac")) for k,v in d["kernels"].items()}); print("  roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["traffic"])
    print("  growth", (d.get("growth") or {}).get("views_per_s"), (d.get("growth") or {}).get("worst_step_over_phase_median"), (d.get("growth") or {}).get("steps_over_1p5x_median"))
    print("  cpu", d.get("cpu_baseline")); nr=d.get("next_rows") or {}
    print("  train_iteration", json.dumps(nr.get("train_iteration"))[:700])
PY
