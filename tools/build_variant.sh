#!/bin/bash
# Build an experiment variant of the library: tools/build_variant.sh NAME [sm100-source] [extra nvcc flags]
# -> 1-stage-wseg_b200/variants/NAME.so (git-ignored, travels with gpurun).  On the GPU box:
#    cp 1-stage-wseg_b200/variants/NAME.so 1-stage-wseg_b200/libpamr_b200.so
set -e
unset CC CXX
NAME=$1; SRC=${2:-pamr_propagate_sm100.cu}; shift; shift || true
cd "$(dirname "$0")/../1-stage-wseg_b200/csrc"
mkdir -p ../variants build
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include -I."
$NV "$@" -c "$SRC" -o build/variant_$NAME.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/$NAME.so build/pamr_capi.o build/pamr_affinity.o build/pamr_propagate.o build/variant_$NAME.o build/pamr_resident.o build/pamr_epilogue.o build/pamr_loss.o -cudart static
echo "built variants/$NAME.so"
