#!/bin/bash
# builds tuning variants of the library: tools/tune/lib_<name>.so  (name threads slots cells ctas)
set -e
cd "$(dirname "$0")/../ochre_b200/csrc"
build() {
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false --extended-lambda -std=c++17 \
    -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -shared \
    -DOC_PK_THREADS=$2 -DOC_PK_SLOTS=$3 -DOC_PK_CELLS=$4 -DOC_PK_CTAS=$5 $6 \
    -o ../../tools/tune/lib_$1.so pipeline.cu host_path.cpp &
}
build a 256 136 2048 2
build b 256 80 1024 3
build c 128 64 1024 4
build d 128 40 512 6
build e 512 320 4096 1
build f 256 56 768 4
build g 128 96 1024 3
wait
ls -la ../../tools/tune/
