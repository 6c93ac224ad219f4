#!/bin/bash
# usage: tools/ncu_k2_variant.sh <tag> <lib.so>   -- ncu --set full of two steady-state k2_hash launches of a library variant
tag=$1; lib=$2
AURICLASS_CUDA_LIB=$PWD/$lib ncu --set full --clock-control none --import-source on -k regex:k2_hash -s 16 -c 2 \
    -o gpurun_out/prof_k2_${tag} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dist > gpurun_out/ncu_k2_${tag}.log 2>&1
echo "ncu $tag rc=$?"
