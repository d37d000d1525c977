#!/bin/bash
for v in NOATOM P1 P1NOATOM; do
  FAQCS_B200_LIB=$PWD/variants_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['roofline']['segments_ms'])"
done
