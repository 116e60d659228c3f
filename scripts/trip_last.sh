#!/usr/bin/env bash
# last check of the shipped build: smoke, the CLI byte-for-byte tests, the firmware tests
set -u
mkdir -p gpurun_out
timeout 60 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 130 python -m pytest tests/test_cli.py tests/test_gpu_firmware.py -m gpu -x -q 2>&1 | tail -2
