#!/bin/bash
# tools/build_variant.sh <name> [-D...]: builds sprl_b200/lib/variants/<name>.so = the library with extra flags for evalnet.cu
# (timing experiments; select it with SPRL_B200_LIB=<path>).  The other objects come from the regular build.
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p sprl_b200/lib/variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" -c sprl_b200/csrc/evalnet.cu -o sprl_b200/lib/variants/$name.o &&
nvcc -shared -o sprl_b200/lib/variants/$name.so sprl_b200/lib/env.o sprl_b200/lib/search.o sprl_b200/lib/engine.o sprl_b200/lib/variants/$name.o -gencode arch=compute_100a,code=sm_100a &&
rm sprl_b200/lib/variants/$name.o && echo built sprl_b200/lib/variants/$name.so
