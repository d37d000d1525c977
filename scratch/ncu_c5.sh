#!/bin/bash
# usage: scratch/ncu_c5.sh REGEX NAME [SKIP]: ncu --set full capture of one kernel on the C5 workload (1 M reads per batch)
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-3} -c 1 -o gpurun_out/$2 -f \
    python bench.py --workload c5 --steps 1 --warmup 3 --batches-per-step 1 --block-pairs 250000 --batch-pairs 1000000 --no-cpu-baseline --e2e-steps 0 --contexts-per-gpu 1 > gpurun_out/$2.log 2>&1
