#!/usr/bin/env bash
# final tree: full validation (tests, smoke, sanitizer, both bench arms) + ncu launch list + ncu --set full of a dtu step
bash tools/gpu_r02s.sh
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 33 -c 33 --csv --log-file gpurun_out/r02t_launches_native_dtu.csv python tools/gpu_step.py native dtu 6 > gpurun_out/r02t_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 22 -c 11 -f -o gpurun_out/r02t_dtu_full python tools/gpu_step.py native dtu 3 > gpurun_out/r02t_ncu.log 2>&1
ls -la gpurun_out/r02t*
