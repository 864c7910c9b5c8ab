#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encode.py -x -q -m gpu > gpurun_out/t25.log 2>&1; echo "tests rc=$?"; grep -n "^E  " gpurun_out/t25.log | head -8 | cut -c1-400; tail -2 gpurun_out/t25.log
