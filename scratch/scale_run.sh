#!/bin/bash
# Scaling evidence on one multi-GPU box: bench.py and the PCIe probe at N = 8, 4, 2 (as many as are visible), multi-GPU tests.
NG=$(python -c "import torch; print(torch.cuda.device_count())")
echo "gpus visible: $NG"
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_multigpu_tests_n$NG.log
for N in 8 4 2; do
  [ $N -le $NG ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 \
      > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N scratch/pcie_probe_n.py \
      > gpurun_out/r2_pcie_n$N.json 2> gpurun_out/r2_pcie_n$N.err
done
