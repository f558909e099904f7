#!/usr/bin/env bash
# Builds libsplade_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v"
mkdir -p build
pids=()
for f in api head_fwd head_bwd head_prologue loss loss_bwd_tc scores_tc topk emit trunk_gemm trunk trunk_bwd peer; do
  ( $NVCC $FLAGS -c $f.cu -o build/$f.o > build/$f.log 2>&1 || { cat build/$f.log; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o ../libsplade_b200.so build/api.o build/head_fwd.o build/head_bwd.o build/head_prologue.o build/loss.o build/loss_bwd_tc.o build/scores_tc.o build/topk.o build/emit.o build/trunk_gemm.o build/trunk.o build/trunk_bwd.o build/peer.o
echo "built $(cd .. && pwd)/libsplade_b200.so"
