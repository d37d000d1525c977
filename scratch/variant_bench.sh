#!/bin/bash
# usage: scratch/variant_bench.sh NAME...   (each NAME is variants_NAME.so in the repo root; "base" = the in-tree library)
for v in "$@"; do
  lib=$PWD/variants_$v.so; [ "$v" = base ] && lib=$PWD/faqcs_b200/libfaqcs_b200.so
  FAQCS_B200_LIB=$lib timeout 300 python bench.py --steps 4 --warmup 3 --batches-per-step 10 --no-cpu-baseline --e2e-steps 0 --contexts-per-gpu 1 2>gpurun_out/variant_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['roofline']['ms_per_batch'], {k: round(x['ms'],4) for k,x in d['roofline']['kernels'].items()}, round(d['roofline']['frac'],4))"
done
