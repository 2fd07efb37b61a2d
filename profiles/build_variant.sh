#!/bin/bash
# Development aid: builds dgfem-acoustic_b200/lib/variants/libdgb_<name>.so with extra nvcc flags for ONE kernel source
# (default stage_ws.cu; e.g. `build_variant.sh te16 stage_bb.cu -DDGB_BB_TE=16`, `build_variant.sh gap2 -DDGB_WS_GAP=2`);
# select it at run time with DGB_LIB=<path>. Usage: profiles/build_variant.sh <name> [source.cu] [flags...]
# Needs the regular build first (python -c "import __graft_entry__ as g; g.build()"): the other objects are reused.
set -e
NAME=$1; shift
SRC=stage_ws.cu
case "$1" in *.cu) SRC=$1; shift;; esac
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG=$ROOT/dgfem-acoustic_b200
mkdir -p $PKG/lib/variants $PKG/build
OBJ=$PKG/build/${SRC%.cu}_$NAME.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fopenmp -I$ROOT/include "$@" -c $PKG/csrc/$SRC -o $OBJ
OTHERS=$(ls $PKG/build/*.cu.o $PKG/build/partition.cpp.o | grep -v "/$SRC.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a $OTHERS $OBJ -o $PKG/lib/variants/libdgb_$NAME.so -ldl -lgomp
echo built $PKG/lib/variants/libdgb_$NAME.so
