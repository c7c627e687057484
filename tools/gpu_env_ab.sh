#!/bin/bash
# A/B visit over environment switches.   usage: tools/gpu_env_ab.sh <tag> "ENV=.. ENV=.." "ENV=.." ...   (STEPS=, WORKLOAD=)
tag=${1:-ab}; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 600 python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline --no-strong-scaling ${WORKLOAD:+--workload $WORKLOAD} \
    > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  python - "$envs" gpurun_out/${tag}_$i.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    r = d["roofline"]
    print(f"{sys.argv[1]:36s} {d['ms_per_step']*1e3:8.1f} us/step  edge {r['avg_launch_ms']*1e3:7.1f} us ({r['frac']:.3f})  "
          f"node {r['node_kernel_share_of_step'] * d['config']['ms_per_step_kernel_leg_eager_with_events'] * 1e3 / d['config']['mp_steps']:6.1f} us  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as exc:
    print(sys.argv[1], "FAILED", exc)
PY
done
