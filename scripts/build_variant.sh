#!/bin/bash
# scripts/build_variant.sh NAME [nvcc flags...] — a libssb_core variant under build/variants/NAME.so: search.cu recompiled with
# the given flags (e.g. -DSV_TIMING=1), the other objects taken from the regular build (run build.py first)
set -e
name=$1; shift
mkdir -p build/variants build/vobj
nvcc -std=c++17 -O3 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -gencode arch=compute_100a,code=sm_100a "$@" \
  -c stringsearchlib_b200/csrc/search.cu -o build/vobj/search_$name.o
nvcc -std=c++17 -O3 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -gencode arch=compute_100a,code=sm_100a "$@" \
  -c stringsearchlib_b200/csrc/sieve.cu -o build/vobj/sieve_$name.o
nvcc -shared -cudart static -o build/variants/$name.so build/obj/primitives.o build/obj/build.o build/vobj/sieve_$name.o build/vobj/search_$name.o \
  build/obj/merge.o build/obj/digest.o build/obj/engine.o -gencode arch=compute_100a,code=sm_100a
echo built build/variants/$name.so
