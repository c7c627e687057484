#!/bin/bash
# A/B visit: L2 eviction hints (LB200_L2HINT = message kernel mask, LB200_L2HINT_NODE = node kernel mask).
# usage: tools/gpu_l2.sh <tag> [edge_mode:node_mode ...]
tag=${1:-l2}; shift
modes=${@:-0:0 5:0 4:0 13:0 5:1 5:2 5:3 13:3 0:0 5:0}
mkdir -p gpurun_out
for m in $modes; do
  LB200_L2HINT=${m%%:*} LB200_L2HINT_NODE=${m##*:} timeout 600 python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline --no-strong-scaling \
    > gpurun_out/${tag}_m$m.json 2> gpurun_out/${tag}_m$m.err
  python - "$m" gpurun_out/${tag}_m$m.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    r = d["roofline"]
    print(f"mode {sys.argv[1]:3s} {d['ms_per_step']*1e3:8.1f} us/step  edge {r['avg_launch_ms']*1e3:7.1f} us ({r['frac']:.3f})  "
          f"node share {r['node_kernel_share_of_step']:.3f}  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as exc:
    print(sys.argv[1], "FAILED", exc)
PY
done
