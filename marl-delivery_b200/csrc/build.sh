#!/bin/bash
# Builds libmdx.so (sm_100a only) next to this script.  Used by __graft_entry__.build().
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
  -Xcompiler -fPIC,-ffp-contract=off,-O2,-Wall -Xptxas -v ${MDX_NVCC_EXTRA:-} \
  -shared -o libmdx.so mdx_api.cu sampler.cpp -lpthread 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning|registers|spill" build.log | grep -v "^$" | head -60 || true
