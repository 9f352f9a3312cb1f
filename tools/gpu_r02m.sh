#!/usr/bin/env bash
N=${1:-8}
mkdir -p gpurun_out
B3GS_AR_SWEEP=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/peer_check.py > gpurun_out/r02m_peer_sweep_n$N.log 2>&1
echo "rc=$?"; grep "world=\|sweep" gpurun_out/r02m_peer_sweep_n$N.log; tail -3 gpurun_out/r02m_peer_sweep_n$N.log | grep -i "error"
