#!/bin/bash
# tools/build_variant.sh <name> [-DMACRO ...]: build/variants/lib_<name>.so with extra macros (A/B runs: APEX_B200_LIB=...)
name=$1; shift
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared --threads 4 "$@" \
  -o build/variants/lib_$name.so apex_b200/csrc/*.cu
