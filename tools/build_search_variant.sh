#!/bin/bash
# tools/build_search_variant.sh <name> [-D...]: builds sprl_b200/lib/variants/<name>.so = the library with extra flags for
# search.cu (timing experiments, e.g. -DSPRL_SEARCH_WARPS_PER_BLOCK=2 -DSPRL_SEARCH_BLOCKS_PER_SM=20); select it with
# SPRL_B200_LIB=<path>.  The other objects come from the regular build.
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p sprl_b200/lib/variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -fmad=false -prec-div=true -prec-sqrt=true -ftz=false "$@" -Xptxas -v -c sprl_b200/csrc/search.cu -o sprl_b200/lib/variants/$name.o 2> sprl_b200/lib/variants/$name.ptxas.log &&
nvcc -shared -o sprl_b200/lib/variants/$name.so sprl_b200/lib/env.o sprl_b200/lib/variants/$name.o sprl_b200/lib/engine.o sprl_b200/lib/evalnet.o -gencode arch=compute_100a,code=sm_100a &&
rm sprl_b200/lib/variants/$name.o && grep -A2 "k_roundINS_7Othello" sprl_b200/lib/variants/$name.ptxas.log | grep -E "registers|spill" && echo built sprl_b200/lib/variants/$name.so
