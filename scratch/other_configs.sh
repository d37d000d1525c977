#!/bin/bash
for wl in c4 c5 c3; do
  extra=""; if [ $wl = c3 ]; then extra="--block-pairs 50000 --batch-pairs 200000 --steps 2 --warmup 3"; else extra="--steps 5 --warmup 3"; fi
  timeout 500 python bench.py --workload $wl --no-cpu-baseline --e2e-steps 0 $extra 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$wl', round(d['value']/1e6,1), 'Mreads/s', round(d['config']['gbases_per_s'],1), 'Gb/s', d['roofline']['segments_ms'])"
done
