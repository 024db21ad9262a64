#!/bin/bash
# Builds libcrnerf_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${CRNERF_OUT:-$HERE/../crnerf_b200/libcrnerf_b200.so}"   # CRNERF_OUT / CRNERF_DEFS: A/B builds (tools/ab_kernel.sh)
OBJ="${CRNERF_OBJ:-$HERE/_obj}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       -Xcompiler -fvisibility=hidden --use_fast_math=false)
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC)
mkdir -p "$OBJ"
pids=()
for f in api nerf_mlp sample crossray style_backward backward backward_gemm gram_tc loss encoder optim; do
  "$NVCC" "${FLAGS[@]}" ${CRNERF_DEFS:-} -c "$HERE/$f.cu" -o "$OBJ/$f.o" ${CRNERF_PTXAS_V:+-Xptxas -v} &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -o "$OUT" "$OBJ"/{api,nerf_mlp,sample,crossray,style_backward,backward,backward_gemm,gram_tc,loss,encoder,optim}.o -lcudart
echo "built $OUT"
