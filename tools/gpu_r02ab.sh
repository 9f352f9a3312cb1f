#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_bins -s 2 -c 2 -f -o gpurun_out/r02ab_bins python tools/gpu_step.py native dtu 3 > gpurun_out/r02ab_ncu.log 2>&1
tail -3 gpurun_out/r02ab_ncu.log; ls -la gpurun_out/r02ab*
