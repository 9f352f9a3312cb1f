#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 24 --csv --log-file gpurun_out/r02c_launches_dtu_bins.csv python tools/gpu_step.py native dtu 6 > gpurun_out/r02c_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tile_bins|bin_apply|bin_chunk|bin_tile" -s 10 -c 5 -o gpurun_out/r02c_bins python tools/gpu_step.py native dtu 4 > gpurun_out/r02c_ncu.log 2>&1
ls -la gpurun_out | tail -5
