#!/bin/bash
set -e
cd gravity2_b200/csrc
for v in 512 256; do
  d=/tmp/mvar_$v; mkdir -p $d
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DBM_PRODUCERS=$v -c boot_mma.cu -o $d/bm.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $d/lib.so api.o gj_kernels.o gom_kernels.o popc_kernels.o philox.o $d/bm.o -cudart static
  echo "== producers=$v"
  (cd ../.. && GV2_LIB=$d/lib.so python tools/boot_bench.py 4096 30000 128 && GV2_LIB=$d/lib.so python tools/boot_bench.py 1000 5000 100 && GV2_LIB=$d/lib.so timeout 100 python tools/mma_check.py 200 700 40 0.3 | tail -1)
done
