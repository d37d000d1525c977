#!/bin/bash
for v in "$@"; do
  lib=$PWD/variants_$v.so; [ "$v" = base ] && lib=$PWD/faqcs_b200/libfaqcs_b200.so
  FAQCS_B200_LIB=$lib timeout 400 python bench.py --workload c3 --block-pairs 50000 --batch-pairs 200000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c3 $v', round(d['value']/1e6,1), d['roofline']['segments_ms'])"
done
