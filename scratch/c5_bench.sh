#!/bin/bash
timeout 500 python bench.py --workload c5 --no-cpu-baseline --e2e-steps 0 --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c5', round(d['value']/1e6,1), 'Mreads/s', d['roofline']['segments_ms'])"
