#!/bin/bash
# Un-profiled bench lines of the other BASELINE configs (C3, C4, C5) with the in-tree library -> gpurun_out/other_<wl>.json
for wl in c3 c4 c5; do
  if [ $wl = c3 ]; then extra="--block-pairs 50000 --batch-pairs 400000 --batches-per-step 2 --steps 3 --warmup 3"; else extra="--batches-per-step 10 --steps 3 --warmup 3"; fi
  timeout 600 python bench.py --workload $wl $extra --no-cpu-baseline --e2e-steps 0 > gpurun_out/other_$wl.json 2> gpurun_out/other_$wl.err
  python -c "
import json,sys; d=json.loads(open('gpurun_out/other_$wl.json').read()); print('$wl', round(d['value']/1e6,1), 'M reads/s', {k: round(x['ms'],4) for k,x in d['roofline']['kernels'].items()}, round(d['roofline']['frac'],4))"
done
