#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02e_bench_native_n2.json 2> gpurun_out/r02e_bench_native_n2.err
echo "n2 rc=$?"; tail -3 gpurun_out/r02e_bench_native_n2.err
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r02e_bench_native_n1.json 2> gpurun_out/r02e_bench_native_n1.err
echo "n1 rc=$?"; tail -3 gpurun_out/r02e_bench_native_n1.err
timeout 600 python -m pytest tests/test_dp_gloo.py -m gpu -q 2>&1 | tail -3
python - <<'PY'
import json
for f in ("r02e_bench_native_n2","r02e_bench_native_n1"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("dp_check"), json.dumps(d.get("configs"))[:400], json.dumps(d.get("growth"))[:900])
    except Exception as ex: print(f, "ERR", ex)
PY
