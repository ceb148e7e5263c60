#!/bin/bash
# Build a variant of the library with extra -D flags on ONE source file (experiments only; the product library is built by
# diffews_b200/_build.py):  scripts/build_variant.sh igemm.cu out.so -DDFW_GNIN_DBG=1
set -e
src=$1; out=$2; shift 2
here=$(cd "$(dirname "$0")/.." && pwd)
obj=/tmp/variant_$$.o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DDFW_BUILD "$@" -c $here/diffews_b200/csrc/$src -o $obj
objs=$(ls $here/diffews_b200/build/*.o | grep -v "/${src%.cu}.o")
nvcc -shared -o $out $obj $objs -gencode arch=compute_100a,code=sm_100a -cudart static
rm -f $obj
