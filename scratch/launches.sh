#!/bin/bash
# ncu launch list (device time + DRAM bytes per launch) of one bench step: gpurun_out/$1.csv
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/$1.csv \
    python bench.py --steps 1 --warmup 3 --batches-per-step 1 --block-pairs 250000 --batch-pairs 1000000 --no-cpu-baseline --e2e-steps 0 --contexts-per-gpu 1 > gpurun_out/$1.log 2>&1
