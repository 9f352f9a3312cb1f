#!/usr/bin/env bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cuda_graph" 2>&1 | tail -25
