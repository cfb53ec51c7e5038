#!/bin/bash
# Compile-time parameter sweep of the production resampler: builds profiles/exp/libss2_<tag>.so per variant
# (only tps.cu is recompiled; the other objects come from stabstitch2_b200/build/).
#   usage: profiles/build_variants.sh tag1:"-DL3_MINB=3 -DL3_NCELL=16" tag2:"..."
set -e
cd "$(dirname "$0")/.."
mkdir -p profiles/exp
for spec in "$@"; do
  tag="${spec%%:*}"; flags="${spec#*:}"
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $flags \
      -c stabstitch2_b200/csrc/tps.cu -o profiles/exp/tps_$tag.o 2>/dev/null
    objs=$(ls stabstitch2_b200/build/*.o | grep -v '/tps.o')
    nvcc -shared -o profiles/exp/libss2_$tag.so $objs profiles/exp/tps_$tag.o -lcudart -lcuda
    echo "built $tag ($flags)" ) &
done
wait
