#!/bin/bash
# builds libsprl_b200.so with extra -D flags for search.cu (timing experiments): tools/build_search_variant.sh -DSPRL_SEARCH_BLOCKS_PER_SM=24
cd "$(dirname "$0")/.."
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -fmad=false -prec-div=true -prec-sqrt=true -ftz=false "$@" -c sprl_b200/csrc/search.cu -o sprl_b200/lib/search.o &&
nvcc -shared -o sprl_b200/lib/libsprl_b200.so sprl_b200/lib/env.o sprl_b200/lib/search.o sprl_b200/lib/engine.o sprl_b200/lib/evalnet.o -gencode arch=compute_100a,code=sm_100a
