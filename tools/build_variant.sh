#!/bin/bash
# Kernel-variant experiments: rebuild only the translation units that see the loop-kernel macros (the C ABI file and the
# two loop TUs) with extra -D flags and link them with the default build's other objects.
# usage: tools/build_variant.sh <tag> "<extra nvcc flags>"   ->  multidronesim_b200/csrc/libmds_<tag>.so  (select with MDS_B200_LIB)
set -e
tag=$1; extra=$2
cd "$(dirname "$0")/../multidronesim_b200/csrc"
make -j8 >/dev/null
mkdir -p build_$tag
FLAGS="-O3 -std=c++17 -lineinfo -use_fast_math -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr $extra"
nvcc $FLAGS -c -o build_$tag/mds_kernels.o mds_kernels.cu &
nvcc $FLAGS -DMDS_TU_REAL=float -DMDS_TU_KIND=0 -c -o build_$tag/rollout_float_0.o mds_rollout_tu.cu &
nvcc $FLAGS -DMDS_TU_REAL=double -DMDS_TU_KIND=0 -c -o build_$tag/rollout_double_0.o mds_rollout_tu.cu &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libmds_$tag.so build_$tag/mds_kernels.o build_$tag/rollout_float_0.o build_$tag/rollout_double_0.o \
  build/rollout_float_1.o build/rollout_float_2.o build/rollout_float_3.o build/rollout_double_1.o build/rollout_double_2.o build/rollout_double_3.o
echo built libmds_$tag.so
