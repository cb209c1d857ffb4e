#!/bin/bash
# Variant of the library with extra nvcc flags on the TMEM engine's translation unit:
#   tools/build_variant_tm.sh NAME "-DFOO=1 ..."  ->  anime4kcpp_b200/lib/libvariant_NAME.so   (run with ACB200_LIB=<that path>; git-ignored scratch)
set -e
cd "$(dirname "$0")/../anime4kcpp_b200/csrc"
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -ccbin g++ -I../../include $2 -c acb200_seg_tm.cu -o build/variants/acb200_seg_tm_$1.o
OBJS=$(ls build/*.o | grep -v "acb200_seg_tm")
nvcc -gencode arch=compute_100a,code=sm_100a -shared --cudart static -ccbin g++ -o ../lib/libvariant_$1.so build/variants/acb200_seg_tm_$1.o $OBJS -lpthread -ldl -lrt
