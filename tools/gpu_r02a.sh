#!/usr/bin/env bash
# round 2, call a: parity against the stock reference + both bench arms on the new default (DTU)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_dp_gloo.py 2>&1 | tail -40 > gpurun_out/r02a_pytest.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02a_bench_reference.json 2> gpurun_out/r02a_bench_reference.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_native.json 2> gpurun_out/r02a_bench_native.err
tail -5 gpurun_out/r02a_pytest.log; cat gpurun_out/r02a_bench_reference.json | cut -c1-600; cat gpurun_out/r02a_bench_native.json | cut -c1-600
