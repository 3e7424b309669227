#!/bin/bash
# Development aid: build libjpegb200 with extra -D flags for the small pass-2 files (k_huffman, k_pack_runs, k_entropy, k_dct's
# cold kernels are left alone) into tools/debug/variants/<name>.so; the other objects come from csrc/_obj (run make first).
# usage: [FILES="k_dct k_pack_runs"] tools/debug/build_variant2.sh <name> [-DJB_FUSE_SCAN=0 ...]
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
C=jpeg-encoder-decoder_b200/csrc
V=tools/debug/variants
mkdir -p $V
objs=""
FILES=${FILES:-"k_huffman k_pack_runs k_entropy"}
for f in $FILES; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off "$@" -c $C/$f.cu -o $V/$name.$f.o 2>&1 | grep -E "error" || true
  objs="$objs $V/$name.$f.o"
done
pat=$(echo $FILES | sed -E "s/ +/.o|/g").o
rest=$(ls $C/_obj/*.o | grep -v -E "$pat")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $V/$name.so $objs $rest
rm $objs
echo built $V/$name.so
