#!/usr/bin/env bash
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N > gpurun_out/r02y_bench_native_n$N.json 2> gpurun_out/r02y_bench_native_n$N.err
echo "bench rc=$?"; tail -2 gpurun_out/r02y_bench_native_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02y_bench_native_n$N.json"))
print("N=$N", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], (d.get("dp_check") or {}).get("ok"), json.dumps({k:(v.get("value"),v.get("ms_per_step")) for k,v in d.get("configs",{}).items()}))
PY
