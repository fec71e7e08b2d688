#!/bin/bash
# usage: tools/build_variant.sh NAME FILE.cu "-DFOO=1 ..."  -> grokimagecompression_b200/libgrok_b200_NAME.so
# rebuilds one translation unit with extra defines and links it with the other, already built, objects
set -e
cd "$(dirname "$0")/../grokimagecompression_b200/csrc"
name=$1; file=$2; defs=$3
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --fmad=false $defs -c $file -o /tmp/var_${name}.o
objs=""
for f in api mct dwt t1_enc t1_dec rd ht; do
  if [ "$f.cu" == "$file" ]; then objs="$objs /tmp/var_${name}.o"; else objs="$objs $f.o"; fi
done
nvcc $ARCH -shared -cudart static -o ../libgrok_b200_${name}.so $objs
echo built libgrok_b200_${name}.so
