#!/bin/bash
# Experiment build: tools/build_variant.sh <file.cu> <name> [nvcc flags...]  ->  csrc/libmv_b200_<name>.so with that one
# source recompiled with the extra flags (e.g. -DMV_SN_TRACE), every other object reused from the last build.py run.
# Load it from the probe tools with MV_ALT_LIB=libmv_b200_<name>.so.  The product library is not touched.
set -e
cd "$(dirname "$0")/../myrtle-vision_b200/csrc"
src=$1; name=$2; shift 2
base=${src%.cu}
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $src -o /tmp/${base}_${name}.o
objs=$(ls *.o | grep -v "^${base}.o$")
nvcc -shared -o libmv_b200_${name}.so $objs /tmp/${base}_${name}.o -gencode arch=compute_100a,code=sm_100a
echo built libmv_b200_${name}.so
