#!/usr/bin/env bash
# N GPUs: exchange check + bench at N with both barrier protocols
N=${1:-2}
mkdir -p gpurun_out
true > gpurun_out/r02i_peer_check_n$N.log 2>&1
echo "peer_check rc=$?"; grep "world=" gpurun_out/r02i_peer_check_n$N.log; tail -3 gpurun_out/r02i_peer_check_n$N.log | grep -i "error\|Traceback" 
for mode in 1:2 1:4 1:8 0:1; do
for cfg in dtu fern_pair; do
B3GS_DP_OVERLAP=${mode%%:*} B3GS_DP_CHUNKS=${mode##*:} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --config $cfg --steps 100 --warmup 10 --no-extra > gpurun_out/r02i_bench_n${N}_${cfg}_$mode.json 2> gpurun_out/r02i_bench_n${N}_${cfg}_$mode.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02i_bench_n${N}_${cfg}_$mode.json"))
    print("N=$N $cfg overlap=$mode", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("dp_check",{}).get("ok"))
except Exception as ex: print("N=$N $cfg $mode ERR", ex)
PY
done
done
