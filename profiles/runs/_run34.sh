#!/bin/bash
# human (configs[3]): the bench line with the tight packing + graph-launched stages, then the stage timeline for two stage sizes
mkdir -p gpurun_out
timeout 1500 python bench.py --workload human --steps 10 --warmup 3 > gpurun_out/bench_human_r2b.json 2> gpurun_out/bench_human_r2b.err; echo "bench rc=$?"
for st in 1048576 2097152 524288; do
  E2FM_HITS_TRACE=1 E2FM_TRACE_TIGHT=1 timeout 600 python profiles/exp/hits_trace.py human $st > gpurun_out/trace_human_tight_$st.log 2>&1; echo "trace $st rc=$?"
  grep wall gpurun_out/trace_human_tight_$st.log
done
E2FM_HITS_TRACE=0 E2FM_TRACE_TIGHT=1 E2FM_HITS_GRAPH=0 timeout 600 python profiles/exp/hits_trace.py human > gpurun_out/trace_human_tight_nograph.log 2>&1; grep wall gpurun_out/trace_human_tight_nograph.log
E2FM_HITS_TRACE=0 E2FM_TRACE_TIGHT=1 timeout 600 python profiles/exp/hits_trace.py human > gpurun_out/trace_human_tight_graph.log 2>&1; grep wall gpurun_out/trace_human_tight_graph.log
python - <<'P'
import json
d=json.load(open("gpurun_out/bench_human_r2b.json")); print(d["value"]/1e9, d["e2e"], d["parity"], d.get("cpu_baseline",{}).get("value"))
P
tail -14 gpurun_out/trace_human_tight_1048576.log
