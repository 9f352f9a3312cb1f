#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_dp_gloo.py 2>&1 | tail -25 > gpurun_out/r02g_pytest.log
tail -3 gpurun_out/r02g_pytest.log
python tools/activation_probe.py 2>&1 | tail -7
