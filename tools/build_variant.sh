#!/bin/bash
# tools/build_variant.sh NAME -DFLAG=... : builds cobaya_b200/lib/variants/NAME.so (kernel experiments;
# select with COBAYA_B200_LIB=...).  Same flags as cobaya_b200/build.py.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p cobaya_b200/lib/variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared \
  -Xptxas -v --expt-relaxed-constexpr "$@" -o cobaya_b200/lib/variants/$name.so cobaya_b200/csrc/engine.cu \
  > cobaya_b200/lib/variants/$name.log 2>&1
echo built $name
