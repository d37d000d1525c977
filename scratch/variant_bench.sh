#!/bin/bash
for v in E4 E5 E6 E8; do
  FAQCS_B200_LIB=$PWD/variants_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e6), d['roofline']['segments_ms'])"
done
