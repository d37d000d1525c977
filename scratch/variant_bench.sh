#!/bin/bash
# usage: scratch/variant_bench.sh NAME...   (each NAME is variants_NAME.so in the repo root; "base" = the in-tree library)
for v in "$@"; do
  lib=$PWD/variants_$v.so; [ "$v" = base ] && lib=$PWD/faqcs_b200/libfaqcs_b200.so
  FAQCS_B200_LIB=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], d['roofline']['segments_ms'])"
done
