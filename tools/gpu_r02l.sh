#!/usr/bin/env bash
# N GPUs: exchange microbench + the bench at N (default workload + configs)
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/peer_check.py > gpurun_out/r02l_peer_check_n$N.log 2>&1
echo "peer_check rc=$?"; grep "world=" gpurun_out/r02l_peer_check_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r02l_bench_native_n$N.json 2> gpurun_out/r02l_bench_native_n$N.err
echo "bench rc=$?"; tail -2 gpurun_out/r02l_bench_native_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02l_bench_native_n$N.json"))
print("N=$N", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("dp_check"), json.dumps({k:(v.get("value"),v.get("ms_per_step")) for k,v in d.get("configs",{}).items()}), d["config"]["parallelism"])
PY
