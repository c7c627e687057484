#!/bin/bash
# Cross-check visit: build with the superseded first-generation kernels and run the comparison tests, then
# restore the product library.   usage: tools/gpu_crosscheck.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
LB200_BUILD_CROSSCHECK=1 python -m lagrangebench_b200.build > gpurun_out/${tag}_build.log 2>&1
LB200_BUILD_CROSSCHECK=1 timeout 900 python -m pytest tests/test_gns_gpu.py -q -m gpu -s > gpurun_out/${tag}_crosscheck.log 2>&1
tail -4 gpurun_out/${tag}_crosscheck.log; grep -E "v2 vs v1" gpurun_out/${tag}_crosscheck.log
python -m lagrangebench_b200.build > /dev/null 2>&1
