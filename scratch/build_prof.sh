#!/bin/bash
# library with per-role clock64 accounting in pg_tile_tc_kernel (scratch/tile_prof.py): only pgtile_tc.cu is recompiled.
# usage: scratch/build_prof.sh MASK   (bit s of MASK records the regions 10 s .. 10 s + 9: 2 = issuers, 4 = splitters, 8 = gather warps,
# 16 = epilogue; 1 = phase stamps only) -> scratch/libkeynet_b200_prof<MASK>.so, loaded with KEYNET_B200_LIB=...
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -DKN_TILE_PROF=${1:-2} -c -o /tmp/pgtile_tc_prof.o keynet_b200/csrc/pgtile_tc.cu
objs=$(ls keynet_b200/lib/obj/*.o | grep -v pgtile_tc.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o scratch/libkeynet_b200_prof${1:-2}.so $objs /tmp/pgtile_tc_prof.o
