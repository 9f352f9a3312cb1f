#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --config lego --steps 10 --warmup 3 --no-extra > gpurun_out/r02z.json 2> gpurun_out/r02z.err
echo rc=$?; tail -4 gpurun_out/r02z.err; python -c "
import json; d=json.load(open('gpurun_out/r02z.json')); print(d['value'], d['dp_check'])"
