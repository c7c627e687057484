#!/bin/bash
# A/B visit over builds of the library: ab/<variant>.so (default: old new) are swapped in turn into the package (the
# build stamp stays, so the loader accepts them).   usage: tools/gpu_so_ab.sh <tag> [rounds]   (VARIANTS=, STEPS=, WORKLOAD=)
tag=${1:-so}; rounds=${2:-3}
cp lagrangebench_b200/_lb200.so /tmp/_lb200_keep.so
for r in $(seq 1 $rounds); do
  for v in ${VARIANTS:-old new}; do
    cp ab/$v.so lagrangebench_b200/_lb200.so
    tools/gpu_env_ab.sh ${tag}_${v}$r "LB200_AB=$v"
  done
done
cp /tmp/_lb200_keep.so lagrangebench_b200/_lb200.so
