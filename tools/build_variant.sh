#!/bin/bash
# Build a variant of the library with extra nvcc flags:  tools/build_variant.sh NAME "-DFOO ..."  -> anime4kcpp_b200/lib/libvariant_NAME.so
# (run it with ACB200_LIB=$PWD/anime4kcpp_b200/lib/libvariant_NAME.so; variants are git-ignored scratch)
set -e
cd "$(dirname "$0")/../anime4kcpp_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -ccbin g++ -I../../include $2 -c acb200.cu -o build/acb200_$1.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared --cudart static -ccbin g++ -o ../lib/libvariant_$1.so build/acb200_$1.o build/CBinding.o build/Image.o build/ImageOps.o build/Model.o build/Processor.o build/Stream.o build/weights.o -lpthread -ldl -lrt
