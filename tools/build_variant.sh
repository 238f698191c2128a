#!/bin/bash
# Build the engine as it is at a git revision into variants/NAME.so (for A/B runs with OM_LIB_PATH):
#   tools/build_variant.sh REV NAME [extra nvcc flags]
REV=$1; NAME=$2; shift 2
D=$(mktemp -d)
mkdir -p $D/openmines_b200/csrc $D/include variants
for f in openmines_b200/csrc/om_engine.cu openmines_b200/csrc/om_engine.cuh openmines_b200/csrc/om_rng.cuh include/openmines_b200.h; do git show $REV:$f > $D/$f; done
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -shared -Xcompiler -fPIC "$@" -I $D/include -o variants/$NAME.so $D/openmines_b200/csrc/om_engine.cu && echo built variants/$NAME.so
rm -rf $D
