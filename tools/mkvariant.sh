#!/bin/bash
# Tuning experiments: build a library variant (k = 27 only, seconds) into gpurun_variants/<name>.so; tools/variants.sh
# then times bench.py's kernel breakdown for every variant in ONE gpurun call (AURICLASS_CUDA_LIB selects the library).
# usage: tools/mkvariant.sh <name> <srcdir> [extra nvcc flags, e.g. -DAC_K2_BRANCHY]
#   <srcdir> = auriclass_b200/csrc, or a copy of it (git archive of another commit) for an A/B against history
name=$1; src=$(realpath $2); shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/gpurun_variants; mkdir -p $out
bd=$(mktemp -d)/build_$name; mkdir -p $bd/auriclass_b200/csrc $bd/include
cp $src/*.cu $src/*.cuh $bd/auriclass_b200/csrc/; cp $root/include/*.h $bd/include/
cd $bd/auriclass_b200/csrc
for f in ac_api ac_parse ac_hash ac_select ac_dist ac_xchg; do
  fl="-DAC_ONLY_K27"
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $fl "$@" -c $f.cu -o $f.o &
done; wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -Xlinker --no-undefined -o $out/$name.so ac_api.o ac_parse.o ac_hash.o ac_select.o ac_dist.o ac_xchg.o -cudart shared -Xlinker -rpath=/usr/local/cuda/lib64 -ldl -lpthread && ls -la $out/$name.so
