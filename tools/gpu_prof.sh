#!/bin/bash
# Profiling visit: ncu --set full captures of the two product kernels + compute-sanitizer passes.
# usage: tools/gpu_prof.sh <tag>
tag=${1:-p}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong-scaling"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:node_mp_tc2_kernel -s 12 -c 2 \
  -o gpurun_out/${tag}_node_mp_tc2_full -f $B > gpurun_out/${tag}_ncu_node.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:edge_mp_tc2_kernel -s 15 -c 2 \
  -o gpurun_out/${tag}_edge_mp_tc2_full -f $B > gpurun_out/${tag}_ncu_edge.log 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python tools/sanitize.py > gpurun_out/${tag}_sanitize_$tool.log 2>&1
  echo "$tool exit $?" | tee -a gpurun_out/${tag}_sanitize_$tool.log
  tail -4 gpurun_out/${tag}_sanitize_$tool.log
done
