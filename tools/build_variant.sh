#!/bin/bash
# usage: tools/build_variant.sh <name> <file.cu> <extra nvcc flags...>   -> hetmogp_b200/lib/var_<name>.so (HMOGP_LIB=...)
set -e
name=$1; file=$2; shift 2
cd "$(dirname "$0")/../hetmogp_b200/csrc"
mkdir -p ../../build/var_$name
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $file -o ../../build/var_$name/${file%.cu}.o
objs=""
for f in engine mm_algebra lik_kernels proj_simt gram_simt tc_fwd tc_gram tc_gram2 tc_bwd optim; do
  if [ "$f.cu" == "$file" ]; then objs="$objs ../../build/var_$name/$f.o"; else objs="$objs ../../build/obj/$f.o"; fi
done
nvcc $ARCH -shared -o ../lib/var_$name.so $objs -lcudart -lcuda
echo built hetmogp_b200/lib/var_$name.so
