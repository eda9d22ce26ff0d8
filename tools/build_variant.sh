#!/bin/bash
# usage: tools/build_variant.sh NAME [nvcc -D flags...]  ->  mvregfus_b200/lib/var_NAME.so (kernel-variant experiments)
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -Xptxas=-v "$@" -I include \
  -o mvregfus_b200/lib/var_$name.so mvregfus_b200/csrc/*.cu -lcufft -lcudart 2>&1 | grep -A2 "${GREP:-fuse_blend_kernelIfLb0}" | grep -E "Used|spill"
