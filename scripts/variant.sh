#!/bin/bash
# dev tool: build a variant of one .cu with extra -D flags into nimrud_b200/lib/variants/<name>.so
# usage: scripts/variant.sh <name> <file.cu> <nvcc flags...>
set -e
name=$1; file=$2; shift 2
HERE=nimrud_b200/csrc
mkdir -p nimrud_b200/lib/variants /tmp/variant_obj
base=$(basename ${file%.cu})
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 --fmad=true "$@" -c $HERE/$file -o /tmp/variant_obj/${base}_$name.o
objs=""
for o in $HERE/obj/*.o; do
  if [ "$(basename $o)" != "$base.o" ]; then objs="$objs $o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o nimrud_b200/lib/variants/$name.so $objs /tmp/variant_obj/${base}_$name.o
echo built $name
