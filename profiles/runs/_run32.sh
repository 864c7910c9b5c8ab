#!/bin/bash
mkdir -p gpurun_out
E2FM_HITS_TRACE=2 E2FM_TRACE_MULT=2 timeout 600 python profiles/exp/hits_trace.py mid50 > gpurun_out/trace2_mid50.log 2>&1; echo "rc=$?"
E2FM_HITS_TRACE=2 E2FM_TRACE_MULT=2 E2FM_HITS_SERIAL=1 timeout 600 python profiles/exp/hits_trace.py mid50 > gpurun_out/trace2_mid50_serial.log 2>&1; echo "rc=$?"
tail -28 gpurun_out/trace2_mid50.log; tail -28 gpurun_out/trace2_mid50_serial.log
