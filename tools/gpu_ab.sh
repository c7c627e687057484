#!/bin/bash
# A/B visit: short bench lines under kernel-variant environment switches.   usage: tools/gpu_ab.sh <tag>
tag=${1:-ab}
mkdir -p gpurun_out
run() {  # name, env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong-scaling "$@" \
    > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - "$name" gpurun_out/${tag}_${name}.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    r = d["roofline"]
    print(f"{sys.argv[1]:28s} {d['ms_per_step']*1e3:8.1f} us/step  edge {r['avg_launch_ms']*1e3:7.1f} us ({r['frac']:.3f})  "
          f"node share {r['node_kernel_share_of_step']:.3f}  e2e {d['e2e']['value']:.3e}  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as exc:
    print(sys.argv[1], "FAILED", exc)
PY
}
run pdl_on LB200_PDL=1 --
run pdl_off LB200_PDL=0 --
run pdl_on2 LB200_PDL=1 --
run pdl_off2 LB200_PDL=0 --
run tgv2d X=1 -- --workload tgv2d
run tgv2d_nopdl LB200_PDL=0 -- --workload tgv2d
run rpf2d X=1 -- --workload rpf2d
run dam2d X=1 -- --workload dam2d
run ldc3d_8k X=1 -- --workload ldc3d
