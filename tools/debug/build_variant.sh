#!/bin/bash
# Development aid: build libjpegb200 with extra -D flags for k_tokens.cu into tools/debug/variants/<name>.so
# usage: tools/debug/build_variant.sh <name> [-DTK_WARPS_PER_CTA=12 -DTK_CTAS_PER_SM=1 ...]
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
C=jpeg-encoder-decoder_b200/csrc
mkdir -p tools/debug/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off "$@" -c $C/k_tokens.cu -o tools/debug/variants/$name.o 2>&1 | grep -E "error" || true
objs=$(ls $C/_obj/*.o | grep -v k_tokens.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/debug/variants/$name.so tools/debug/variants/$name.o $objs
rm tools/debug/variants/$name.o
echo built tools/debug/variants/$name.so
