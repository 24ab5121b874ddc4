#!/bin/sh
# Builds nutils_b200/libb200fem.so for sm_100a (in-tree; the .so travels to the GPU box).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$HERE/../libb200fem.so"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -cudart static"
mkdir -p "$HERE/_obj"
SRCS="api api_elemset pattern pattern_elemset pattern_general assemble_generic assemble_elemset assemble_fast assemble_rows spmv"
OBJS=""
PIDS=""
for f in $SRCS; do
  src="$HERE/$f.cu"; obj="$HERE/_obj/$f.o"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ "$HERE/common.cuh" -nt "$obj" ] || [ "$HERE/../../include/b200fem.h" -nt "$obj" ]; then
    # translation units are independent: compile them side by side
    "$NVCC" $FLAGS ${B2_PTXAS_V:+-Xptxas -v} ${B2_EXPERIMENT:+-DB2_EXPERIMENT} -c "$src" -o "$obj" &
    PIDS="$PIDS $!"
  fi
  OBJS="$OBJS $obj"
done
for pid in $PIDS; do
  wait "$pid"
done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o "$OUT" $OBJS
echo "built $OUT"
