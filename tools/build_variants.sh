#!/bin/bash
# Builds tuning variants of libmcl_cuda.so HERE (nvcc cross-compiles) into botlab_b200/variants/, so a single gpurun call
# can bench them all with MCL_LIB=<path> without spending GPU-minutes on compilation.
#   tools/build_variants.sh name1="-DFLAG=1 ..." name2="..."
set -e
cd "$(dirname "$0")/../botlab_b200/csrc"
mkdir -p ../variants
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  ( make -s -B OUT=../variants/libmcl_$name.so EXTRA="$flags" >/dev/null 2>&1 && cp build.log ../variants/$name.log \
    && echo "$name: $(grep -A2 'score_kernelILi1ELb1ELb1ELb0' ../variants/$name.log | grep -o 'Used [0-9]* registers') [$flags]" ) || echo "$name: BUILD FAILED"
done
