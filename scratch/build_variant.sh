#!/bin/bash
# scratch/build_variant.sh NAME -DFLAG...: libkeynet_b200 with pgtile_tc.cu recompiled under extra flags -> scratch/libkeynet_b200_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden "$@" -c -o /tmp/pgtile_tc_$name.o keynet_b200/csrc/pgtile_tc.cu
objs=$(ls keynet_b200/lib/obj/*.o | grep -v pgtile_tc.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o scratch/libkeynet_b200_$name.so $objs /tmp/pgtile_tc_$name.o
