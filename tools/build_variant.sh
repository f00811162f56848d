#!/bin/bash
# A/B builds of the same ABI: tools/build_variant.sh <tag> <file.cu> [extra nvcc flags...] -> build/libwwb200_<tag>.so
# (one source recompiled with the extra flags, the other objects taken from the in-tree build; use with WWB200_LIB=...)
set -e
tag=$1; src=$2; shift 2
cd "$(dirname "$0")/.."
python -m wakeword_detection_b200.build > /dev/null
mkdir -p build
C=wakeword_detection_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c $C/$src -o build/${src%.cu}_$tag.o
objs=""
for o in $C/*.o; do
  b=$(basename $o .o)
  [ -f $C/$b.cu ] || continue
  if [ "$b.cu" == "$src" ]; then objs="$objs build/${b}_$tag.o"; else objs="$objs $o"; fi
done
nvcc -shared -o build/libwwb200_$tag.so $objs -lcudart
echo build/libwwb200_$tag.so
