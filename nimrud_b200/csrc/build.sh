#!/bin/bash
# builds nimrud_b200/lib/libnimrud_b200.so for sm_100a (nvcc cross-compiles without a GPU)
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/obj"
NVCC="${NVCC:-nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 --fmad=true ${NBR_EXTRA_NVCC_FLAGS}"
pids=()
for f in "$HERE"/*.cu; do
    o="$HERE/obj/$(basename "${f%.cu}").o"
    if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find "$HERE" -name '*.cuh' -newer "$o")" ] || [ "$HERE/../../include/nimrud_b200.h" -nt "$o" ]; then
        $NVCC $FLAGS -c "$f" -o "$o" &
        pids+=($!)
    fi
done
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libnimrud_b200.so" "$HERE"/obj/*.o
echo "built $OUT/libnimrud_b200.so"
