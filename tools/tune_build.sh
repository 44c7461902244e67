#!/bin/bash
# builds tuning variants of the library: tools/tune/lib_<name>.so  (name threads slots cells ctas [extra flags])
set -e
cd "$(dirname "$0")/../ochre_b200/csrc"
mkdir -p ../../tools/tune
build() {
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false --extended-lambda -std=c++17 \
    -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -shared \
    -DOC_PK_THREADS=$2 -DOC_PK_SLOTS=$3 -DOC_PK_CELLS=$4 -DOC_PK_CTAS=$5 $6 \
    -o ../../tools/tune/lib_$1.so pipeline.cu host_path.cpp &
}
while read -r name thr slots cells ctas extra; do
  [ -z "$name" ] && continue
  build "$name" "$thr" "$slots" "$cells" "$ctas" "$extra"
done
wait
ls -la ../../tools/tune/
