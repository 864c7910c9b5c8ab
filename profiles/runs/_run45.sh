#!/bin/bash
# small batches (a GPU's share at N = 8 / N = 4): which launch form is best
mkdir -p gpurun_out
export E2FM_TRACE_TIGHT=1 E2FM_HITS_TRACE=0 E2FM_TRACE_REPS=8
python -c "import bench; bench.prepare('mid50', 0)" > /dev/null 2>&1
for n in 500000 1250000 2500000; do
  for v in "final:" "plain:E2FM_HITS_GRAPH=0 E2FM_HITS_OVERLAP=0" "graph_only:E2FM_HITS_OVERLAP=0" "overlap_only:E2FM_HITS_GRAPH=0"; do
    name=${v%%:*}; envs=${v#*:}
    echo -n "n=$n $name: "; env $envs E2FM_TRACE_N=$n timeout 300 python profiles/exp/hits_trace.py mid50 2>&1 | grep wall | tail -5 | awk '{printf "%s ", $3}'; echo
  done
done
