#!/bin/bash
# usage: scripts/build_variant.sh <name> <extra nvcc flags...>: scratch_libs/<name>.so = the library with builtin_expreg.cu and
# api.cu recompiled with the extra flags (the other translation units are taken from csrc/_build)
set -e
name=$1; shift
B=mcmcf90_b200/csrc/_build; V=/tmp/variant_$name; mkdir -p $V scratch_libs
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -Imcmcf90_b200/csrc"
nvcc $F "$@" -c mcmcf90_b200/csrc/builtin_expreg.cu -o $V/builtin_expreg.o &
nvcc $F "$@" -c mcmcf90_b200/csrc/api.cu -o $V/api.o &
wait
objs=$(ls $B/*.o | grep -v -e builtin_expreg.o -e api.o)
nvcc -shared -gencode arch=compute_100a,code=sm_100a $objs $V/builtin_expreg.o $V/api.o -Xlinker -soname=libmcmcb200.so -ldl -o scratch_libs/$name.so
echo built scratch_libs/$name.so
