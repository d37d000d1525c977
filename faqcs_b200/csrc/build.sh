#!/bin/bash
# Build libfaqcs_b200.so (C ABI + sm_100a kernels) in-tree: faqcs_b200/libfaqcs_b200.so
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${FQ_OUT:-$HERE/../libfaqcs_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC,-Wall,-Wno-unused-function -shared ${FQ_NVCC_EXTRA:-} \
    -ccbin /usr/bin/g++ \
    "$HERE/fq_api.cu" -o "$OUT" -lcudart -ldl
echo "built $OUT"
