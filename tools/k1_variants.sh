#!/bin/bash
# build gj_kernels.cu variants on the GPU box and time them (experiment helper, not part of the product)
set -e
cd gravity2_b200/csrc
for v in "128 2 0" "128 1 0" "128 4 0" "256 2 0" "512 2 0" "128 8 0" "256 2 1"; do
  set -- $v
  d=/tmp/var_$1_$2_$3; mkdir -p $d
  extra="-DGJ_SLAB=$1 -DGJ_KG_UNROLL=$2"; [ "$3" = "1" ] && extra="$extra -DGJ_W_TREE"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $extra -c gj_kernels.cu -o $d/gj.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $d/lib.so api.o $d/gj.o gom_kernels.o popc_kernels.o philox.o boot_mma.o -cudart static
  echo "== slab=$1 unroll=$2 wtree=$3"
  (cd ../.. && GV2_LIB=$d/lib.so python tools/k1bench.py)
done
