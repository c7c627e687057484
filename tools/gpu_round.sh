#!/bin/bash
# One GPU-box visit: parity tests, the default bench line, a per-launch device-time list.
# usage: tools/gpu_round.sh <tag> [pytest-args...]     (outputs under gpurun_out/<tag>_*)
tag=${1:-run}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${tag}_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_tests.log
tail -5 gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 1500 gpurun_out/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong-scaling > gpurun_out/${tag}_launches.log 2>&1
python profiles/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches.summary.txt 2>&1
head -30 gpurun_out/${tag}_launches.summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
