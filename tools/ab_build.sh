#!/bin/bash
# A/B builds of libmrf_b200.so with extra -D flags (developer tool).  usage: tools/ab_build.sh name "-DFOO=1 -DBAR" [name2 "flags2" ...]
# Outputs build_ab/<name>.so (git-ignored, travels with gpurun); tools/ab_rollout.py times them on the GPU.
set -e
cd "$(dirname "$0")/.."
mkdir -p build_ab
pids=()
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $flags \
      -Xptxas -v -o build_ab/$name.so multi-robot-fabrics_b200/csrc/mrf_b200.cu > build_ab/$name.log 2>&1 \
      && echo "built $name" || echo "FAILED $name (build_ab/$name.log)" ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
