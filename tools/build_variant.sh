#!/bin/bash
# Builds a kernel variant of libpfilter_b200.so under build_variants/<name>/ (git-ignored, travels to the GPU box):
#   tools/build_variant.sh <name> "<extra nvcc flags>" file.cu [file.cu ...]
# Only the named sources are recompiled with the extra flags; the other objects are taken from the main build.
# Select it at run time with PFILTER_B200_LIB=build_variants/<name>/libpfilter_b200.so.
set -e
name=$1; extra=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/pfilter-noetic_b200/csrc
out=$root/build_variants/$name
mkdir -p "$out"
flags="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-Wall,-Wno-unused-function -I$root/include"
objs=""
for f in "$src"/*.cu; do
  b=$(basename "$f" .cu)
  if [[ " $* " == *" $b.cu "* ]]; then
    /usr/local/cuda/bin/nvcc $flags $extra -c "$f" -o "$out/$b.o"
    objs="$objs $out/$b.o"
  else
    objs="$objs $src/$b.o"
  fi
done
/usr/local/cuda/bin/nvcc -shared -o "$out/libpfilter_b200.so" $objs -lcudart
echo "$out/libpfilter_b200.so"
