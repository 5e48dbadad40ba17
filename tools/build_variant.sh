#!/bin/bash
# development aid: build variants/lib_<name>.so with extra nvcc flags (A/B runs via NBLDPC_LIB=variants/lib_<name>.so)
name=$1; shift
cd "$(dirname "$0")/.."; mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -cudart static \
     -Iinclude "$@" -o variants/lib_$name.so nbldpc_b200/csrc/decoder.cu nbldpc_b200/csrc/encoder.cu nbldpc_b200/csrc/code.cpp nbldpc_b200/csrc/comm.cpp -ldl
