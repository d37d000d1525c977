#!/bin/bash
# Round-2 evidence run on one GPU: tests, bench lines (both arms), other configs, launch list, ncu captures, sanitizer.
set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_final_tests.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
bash scratch/other_configs.sh > gpurun_out/r2_other_final.txt 2>&1
python scratch/pcie_probe_n.py > gpurun_out/r2_pcie_n1.json 2>/dev/null
bash scratch/launches.sh r2_launches_final
bash scratch/ncu_kernel.sh k_trim r2_k_trim
bash scratch/ncu_kernel.sh "k_emit" r2_k_emit 3
bash scratch/ncu_kernel.sh "k_frame_lines" r2_k_frame 6
bash scratch/ncu_c3.sh r2_k_adapter
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "micro or crlf or routing or third or lone or pieces" 2>&1 | tail -5 > gpurun_out/r2_sanitizer_memcheck.txt
