#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_dp_gloo.py 2>&1 | tail -25 > gpurun_out/r02f_pytest.log
tail -3 gpurun_out/r02f_pytest.log
for a in 1 0; do
B3GS_ASYNC=$a timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02f_bench_async$a.json 2> gpurun_out/r02f_bench_async$a.err
echo "rc=$?"; tail -2 gpurun_out/r02f_bench_async$a.err
B3GS_ASYNC=$a timeout 900 python bench.py --config lego --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02f_bench_lego_async$a.json 2> gpurun_out/r02f_bench_lego_async$a.err
B3GS_ASYNC=$a timeout 900 python bench.py --config fern_pair --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02f_bench_fern_async$a.json 2> gpurun_out/r02f_bench_fern_async$a.err
done
python - <<'PY'
import json
for f in ("r02f_bench_async1","r02f_bench_async0","r02f_bench_lego_async1","r02f_bench_lego_async0","r02f_bench_fern_async1","r02f_bench_fern_async0"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "ratio %.3f"%(d["e2e"]["value"]/d["value"]), json.dumps(d.get("growth"))[:700])
    except Exception as ex: print(f, "ERR", ex)
PY
