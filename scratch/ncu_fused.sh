#!/bin/bash
# ncu --set full capture of the fused kernel (1 M pairs per batch), written to gpurun_out/$1.ncu-rep
name=${1:-fused}
ncu --set full --clock-control none --import-source on -k regex:k_trim_emit -s 3 -c 1 -o gpurun_out/$name -f \
    python bench.py --steps 1 --warmup 3 --batches-per-step 1 --block-pairs 250000 --batch-pairs 1000000 --no-cpu-baseline --e2e-steps 0 > gpurun_out/$name.log 2>&1
