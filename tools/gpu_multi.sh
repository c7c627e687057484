#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): decomposition check against the single-GPU rollout, then the
# driver's own bench command line at N ranks.   usage: tools/gpu_multi.sh <tag> <N> [steps]
tag=${1:-m}; N=${2:-2}; K=${3:-20}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 tools/dist_check.py --case rpf3d_125k --steps 4 > gpurun_out/${tag}_distcheck.log 2>&1
tail -3 gpurun_out/${tag}_distcheck.log
timeout 300 $TR --master-port 29542 tools/dist_check.py --case rpf3d_8k --steps 6 --spread 0.12 --mp 3 > gpurun_out/${tag}_distcheck_spread.log 2>&1
tail -2 gpurun_out/${tag}_distcheck_spread.log
timeout 900 $TR --master-port 29543 bench.py --gpus $N --steps $K --warmup 3 > gpurun_out/${tag}_bench_g$N.json 2> gpurun_out/${tag}_bench_g$N.err
tail -c 3000 gpurun_out/${tag}_bench_g$N.json
tail -5 gpurun_out/${tag}_bench_g$N.err
