#!/bin/bash
# Development aid: builds dgfem-acoustic_b200/lib/variants/libdgb_<name>.so with extra nvcc flags for stage_ws.cu
# (e.g. -DDGB_WS_GAP=2); select it at run time with DGB_LIB=<path>. Usage: profiles/build_variant.sh <name> [flags...]
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG=$ROOT/dgfem-acoustic_b200
mkdir -p $PKG/lib/variants $PKG/build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I$ROOT/include "$@" -c $PKG/csrc/stage_ws.cu -o $PKG/build/stage_ws_$NAME.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a $PKG/build/dgb_api.cu.o $PKG/build/stage_generic.cu.o $PKG/build/stage_tiled.cu.o $PKG/build/stage_ws_$NAME.o $PKG/build/partition.cpp.o -o $PKG/lib/variants/libdgb_$NAME.so -ldl -lgomp
echo built $PKG/lib/variants/libdgb_$NAME.so
