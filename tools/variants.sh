#!/bin/bash
# Build experimental variants of the library that differ in lz77_kernel compile-time constants, for A/B runs on the GPU box:
#   tools/variants.sh NAME "-DZK_LZ_THREADS=768 ..."   ->  zlib.kotlin_b200/build/libzlibk_NAME.so   (use with ZLIBK_LIB=...)
set -e
cd "$(dirname "$0")/.."
NAME=$1; FLAGS=$2
B=zlib.kotlin_b200/build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v $FLAGS \
  -c zlib.kotlin_b200/csrc/zk_lz77.cu -o $B/zk_lz77_$NAME.o 2> $B/zk_lz77_$NAME.o.log
grep -h "registers\|spill" $B/zk_lz77_$NAME.o.log | tail -2
OBJS=""
for f in zk_adler zk_crc32 zk_huffman zk_pack zk_inflate zk_api; do OBJS="$OBJS $B/$f.o"; done
nvcc -shared -o $B/libzlibk_$NAME.so $OBJS $B/zk_lz77_$NAME.o -Xcompiler -fPIC
echo $B/libzlibk_$NAME.so
