#!/bin/bash
# Variant build of the library (development tool): tools/build_variant.sh <tag> <file.cu> [nvcc flags...]
#   -> mofem_b200/csrc/build/libvar_<tag>.so  (<file.cu> recompiled with the extra flags, the other objects reused)
set -e
tag=$1; src=$2; shift 2
B=mofem_b200/csrc/build
extra=""
case $src in integrate.cu|sample.cu|rho.cu|overlay.cu) extra="-fmad=false";; esac
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $extra "$@" \
  -Xptxas -v -c mofem_b200/csrc/$src -o $B/var_${tag}_${src%.cu}.o 2> $B/var_${tag}.ptxas
objs=""
for o in $B/integrate.o $B/sample.o $B/rho.o $B/overlay.o $B/csr.o $B/scan.o $B/pcg.o $B/api.o; do
  if [ "$o" == "$B/${src%.cu}.o" ]; then objs="$objs $B/var_${tag}_${src%.cu}.o"; else objs="$objs $o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $B/libvar_$tag.so $objs -lcudart -ldl
echo $B/libvar_$tag.so
