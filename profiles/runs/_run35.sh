#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "search_hits or seed_search" > gpurun_out/t21.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t21.log
export E2FM_TRACE_MULT=2 E2FM_TRACE_TIGHT=1
E2FM_HITS_TRACE=1 timeout 600 python profiles/exp/hits_trace.py mid50 > gpurun_out/trace1_overlap.log 2>&1; echo "rc=$?"
E2FM_HITS_TRACE=0 timeout 600 python profiles/exp/hits_trace.py mid50 > gpurun_out/trace0_overlap.log 2>&1; echo "rc=$?"
E2FM_HITS_TRACE=0 E2FM_HITS_OVERLAP=0 timeout 600 python profiles/exp/hits_trace.py mid50 > gpurun_out/trace0_nooverlap.log 2>&1; echo "rc=$?"
E2FM_HITS_TRACE=0 E2FM_HITS_GRAPH=0 timeout 600 python profiles/exp/hits_trace.py mid50 > gpurun_out/trace0_overlap_nograph.log 2>&1; echo "rc=$?"
tail -12 gpurun_out/trace1_overlap.log; grep wall gpurun_out/trace0_overlap.log gpurun_out/trace0_nooverlap.log gpurun_out/trace0_overlap_nograph.log
