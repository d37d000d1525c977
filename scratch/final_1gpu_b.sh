#!/bin/bash
# Second evidence run of round 2 (after the k-mer, two-context and adapter-sweep work): bench line, other configs, launch list,
# ncu of k_adapter, memcheck of the new kernels.
set -x
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
bash scratch/other_configs.sh > gpurun_out/r2_other_final.txt 2>&1
bash scratch/launches.sh r2_launches_final
bash scratch/ncu_c3.sh r2_k_adapter
compute-sanitizer --tool memcheck python -m pytest tests/test_kmer.py tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x -k "naive or unaligned or two_contexts or micro_adapter or c3_adapters or polya" 2>&1 | tail -5 > gpurun_out/r2_sanitizer_memcheck_b.txt
cat gpurun_out/r2_sanitizer_memcheck_b.txt; cat gpurun_out/r2_other_final.txt | tail -4
