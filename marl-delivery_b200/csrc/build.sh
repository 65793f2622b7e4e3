#!/bin/bash
# Builds libmdx.so (sm_100a only) next to this script.  Used by __graft_entry__.build().
# The kernel-mode translation units compile in parallel; objects live in ./obj (git-ignored).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-O2,-Wall -Xptxas -v ${MDX_NVCC_EXTRA:-}"
mkdir -p obj
: > build.log
pids=()
for src in mdx_api.cu mdx_launch1.cu mdx_launch2.cu mdx_launch3.cu mdx_launch5.cu mdx_launch1t.cu mdx_launch2t.cu mdx_launch3t.cu mdx_extra.cu sampler.cpp mdx_jit.cpp; do
  [ -f "$src" ] || continue
  ( $NVCC $FLAGS -c -o "obj/${src%.*}.o" "$src" > "obj/${src%.*}.log" 2>&1 ) &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
cat obj/*.log >> build.log
if [ $rc -ne 0 ]; then cat build.log; exit 1; fi
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o libmdx.so obj/*.o -lpthread -ldl >> build.log 2>&1 || { cat build.log; exit 1; }
grep -E "error|warning|spill" build.log | grep -v " 0 bytes spill stores, 0 bytes spill loads" | head -40 || true
