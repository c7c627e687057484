#!/bin/bash
tag=${1:-n}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong-scaling"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:node_mp_tc2_kernel -s 12 -c 2 \
  -o gpurun_out/${tag}_node_mp_tc2_full -f $B > gpurun_out/${tag}_ncu_node.log 2>&1
tail -2 gpurun_out/${tag}_ncu_node.log
