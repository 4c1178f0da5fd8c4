#!/bin/bash
# Experiment variant of the library with pamr_resident.cu compiled with extra flags:
#   tools/build_variant_res.sh NAME [extra nvcc flags]  ->  1-stage-wseg_b200/variants/NAME.so  (use with PAMR_B200_LIB=...)
set -e
unset CC CXX
NAME=$1; shift
cd "$(dirname "$0")/../1-stage-wseg_b200/csrc"
mkdir -p ../variants build
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include -I."
$NV "$@" -c pamr_resident.cu -o build/variant_$NAME.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/$NAME.so build/pamr_capi.o build/pamr_affinity.o build/pamr_propagate.o build/pamr_propagate_sm100.o build/variant_$NAME.o build/pamr_epilogue.o build/pamr_loss.o -cudart static
echo "built variants/$NAME.so"
