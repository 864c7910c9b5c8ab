#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_encode.py -x -q -m gpu > gpurun_out/t22.log 2>&1; echo "tests rc=$?"; grep -n "^E  " gpurun_out/t22.log | head -12 | cut -c1-400; tail -3 gpurun_out/t22.log
