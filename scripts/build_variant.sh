#!/bin/bash
# Build one more copy of the library with extra nvcc flags into build/variants/<name>.so (travels to the GPU box;
# selected there with WAFER_B200_LIB): bash scripts/build_variant.sh <name> "<extra nvcc flags>"
set -eu
NAME=$1; FLAGS=${2:-}
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a $FLAGS -O3 -std=c++17 -lineinfo -fmad=false \
  -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr -Xptxas -v -shared \
  -o build/variants/$NAME.so wafer_b200/csrc/wafer_b200.cu -ldl 2> build/variants/$NAME.ptxas.log
grep -A2 "sweep_tb2_kernelILb0" build/variants/$NAME.ptxas.log | grep -E "Used|spill" | tr '\n' ' '; echo " <- $NAME"
