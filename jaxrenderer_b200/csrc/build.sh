#!/usr/bin/env bash
# Build libjr_b200.so for sm_100a, in-tree (the .so travels to the GPU box).
# -fmad=false: the kernels reproduce the oracle's scalar fp32 op order bit for
# bit, so no multiply-add contraction (see jr_device.cuh).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../lib"
mkdir -p "$out"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
srcs=("$here"/jr_common.cu "$here"/jr_forward.cu "$here"/jr_camera.cu)
[ -f "$here/jr_backward.cu" ] && srcs+=("$here/jr_backward.cu")
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false \
  -std=c++17 -diag-suppress 128 -Xcompiler -fPIC -shared ${JR_NVCC_EXTRA:-} \
  -o "$out/libjr_b200.so" "${srcs[@]}"
echo "built $out/libjr_b200.so"
