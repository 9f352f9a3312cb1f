#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tile_bins" -s 4 -c 2 -o gpurun_out/r02o_bins python tools/gpu_step.py native dtu 4 > gpurun_out/r02o_ncu.log 2>&1
ls -la gpurun_out/r02o_bins.ncu-rep
