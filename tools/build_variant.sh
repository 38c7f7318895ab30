#!/bin/bash
# builds libsprl_b200.so with extra -D flags for evalnet.cu (timing experiments): tools/build_variant.sh -DSPRL_EVALNET_CLUSTER=2 ...
cd "$(dirname "$0")/.."
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" -c sprl_b200/csrc/evalnet.cu -o sprl_b200/lib/evalnet.o &&
nvcc -shared -o sprl_b200/lib/libsprl_b200.so sprl_b200/lib/env.o sprl_b200/lib/search.o sprl_b200/lib/engine.o sprl_b200/lib/evalnet.o -gencode arch=compute_100a,code=sm_100a
