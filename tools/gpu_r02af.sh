#!/usr/bin/env bash
# diagnostic: where does the end-to-end loop lose time at N = 8?  the same run without the per-step upload
N=${1:-8}
mkdir -p gpurun_out
B3GS_BENCH_E2E_NO_UPLOAD=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --no-extra > gpurun_out/r02af_n$N.json 2> gpurun_out/r02af_n$N.err
echo "bench rc=$?"; tail -2 gpurun_out/r02af_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02af_n$N.json"))
print("N=$N no-upload", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"])
PY
