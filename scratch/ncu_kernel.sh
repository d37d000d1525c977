#!/bin/bash
# usage: scratch/ncu_kernel.sh REGEX NAME : ncu --set full capture of one kernel (1 M pairs per batch) -> gpurun_out/NAME.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-3} -c 1 -o gpurun_out/$2 -f \
    python bench.py --steps 1 --warmup 3 --batches-per-step 1 --block-pairs 250000 --batch-pairs 1000000 --no-cpu-baseline --e2e-steps 0 --contexts-per-gpu 1 > gpurun_out/$2.log 2>&1
