#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02h_bench_reference.json 2> gpurun_out/r02h_bench_reference.err
echo "ref rc=$?"; tail -2 gpurun_out/r02h_bench_reference.err
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r02h_bench_native.json 2> gpurun_out/r02h_bench_native.err
echo "nat rc=$?"; tail -2 gpurun_out/r02h_bench_native.err
python - <<'PY'
import json
for f in ("r02h_bench_reference","r02h_bench_native"):
    d=json.load(open("gpurun_out/%s.json"%f))
    print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], json.dumps(d.get("configs"))[:300])
    print("  next_rows", json.dumps(d.get("next_rows"))[:2500])
    print("  growth", json.dumps(d.get("growth"))[:600])
PY
