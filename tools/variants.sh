#!/bin/bash
# times bench.py's kernel breakdown for every library variant under gpurun_variants/ (tuning experiments)
for lib in gpurun_variants/*.so; do
  AURICLASS_CUDA_LIB=$PWD/$lib timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dist 2>gpurun_out/variant_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', round(d['value'],1), {k: round(v,3) for k,v in d['kernel_ms'].items()})"
done
