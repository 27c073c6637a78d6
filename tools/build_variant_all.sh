#!/bin/bash
# Kernel-variant experiments over EVERY translation unit: full rebuild with extra -D flags into its own object directory.
# usage: tools/build_variant_all.sh <tag> "<extra nvcc flags>"   ->  multidronesim_b200/csrc/libmds_<tag>.so  (select with MDS_B200_LIB)
set -e
tag=$1; extra=$2
cd "$(dirname "$0")/../multidronesim_b200/csrc"
make -j8 OBJDIR=build_$tag TARGET=libmds_$tag.so EXTRA="$extra" > /dev/null
echo built libmds_$tag.so
