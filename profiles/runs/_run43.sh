#!/bin/bash
# human (configs[3]): stage timeline of e2fm_search_hits with the final pipeline; wall times with / without the second compute stream and the graph launches
mkdir -p gpurun_out
export E2FM_TRACE_TIGHT=1
E2FM_HITS_TRACE=1 timeout 900 python profiles/exp/hits_trace.py human > gpurun_out/trace_human_final.log 2>&1; echo "trace rc=$?"
E2FM_HITS_TRACE=0 timeout 600 python profiles/exp/hits_trace.py human > gpurun_out/wall_human_final.log 2>&1; grep wall gpurun_out/wall_human_final.log | tr '\n' ' '; echo
E2FM_HITS_TRACE=0 E2FM_HITS_OVERLAP=0 timeout 600 python profiles/exp/hits_trace.py human > gpurun_out/wall_human_nooverlap.log 2>&1; grep wall gpurun_out/wall_human_nooverlap.log | tr '\n' ' '; echo
E2FM_HITS_TRACE=0 E2FM_HITS_GRAPH=0 timeout 600 python profiles/exp/hits_trace.py human > gpurun_out/wall_human_nograph.log 2>&1; grep wall gpurun_out/wall_human_nograph.log | tr '\n' ' '; echo
E2FM_HITS_TRACE=0 E2FM_HITS_GRAPH=0 E2FM_HITS_OVERLAP=0 timeout 600 python profiles/exp/hits_trace.py human > gpurun_out/wall_human_plain.log 2>&1; grep wall gpurun_out/wall_human_plain.log | tr '\n' ' '; echo
tail -17 gpurun_out/trace_human_final.log
