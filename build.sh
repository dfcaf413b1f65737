#!/usr/bin/env bash
# Builds libtagg.so (sm_100a only) in-tree and the CPU oracle.  Used by __graft_entry__.build().
set -euo pipefail
cd "$(dirname "$0")"
SRC=tantivy_aggregations_b200/csrc
OUT=tantivy_aggregations_b200/libtagg.so
OBJ=build/obj
mkdir -p "$OBJ"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall -Iinclude"
pids=()
SRCS="api exec generic stream stream_inst_rt_a stream_inst_rt_b stream_inst_ct_terms stream_inst_ct_hist stream_inst_ct_root mterms pct columns result compact topk comm docset"
for f in $SRCS; do
  if [ ! -f "$OBJ/$f.o" ] || [ "$SRC/$f.cu" -nt "$OBJ/$f.o" ] || [ -n "$(find $SRC include -name '*.h' -newer "$OBJ/$f.o" -o -name '*.cuh' -newer "$OBJ/$f.o" | head -1)" ]; then
    ( $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c "$SRC/$f.cu" -o "$OBJ/$f.o" ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" $(for f in $SRCS; do echo $OBJ/$f.o; done) -lcudart -ldl
make -s -C oracle
echo "built $OUT"
# typed C++ host facade: compile its test driver (host-only code over the C ABI)
g++ -O2 -std=c++17 -Wall -Iinclude -o tests/cpp/test_reference.bin tests/cpp/test_reference.cpp -Ltantivy_aggregations_b200 -ltagg -Wl,-rpath,'$ORIGIN/../../tantivy_aggregations_b200' -L/usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64
echo "built tests/cpp/test_reference.bin"
