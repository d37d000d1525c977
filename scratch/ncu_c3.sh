#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:k_adapter -s 2 -c 1 -o gpurun_out/$1 -f \
    python bench.py --workload c3 --steps 1 --warmup 3 --batches-per-step 1 --block-pairs 50000 --batch-pairs 50000 --no-cpu-baseline --e2e-steps 0 --contexts-per-gpu 1 > gpurun_out/$1.log 2>&1
