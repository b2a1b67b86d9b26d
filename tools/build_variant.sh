#!/usr/bin/env bash
# build an A/B variant of the library: tools/build_variant.sh NAME "-DJR_V3_K32_CTAS=4 ..."
# -> jaxrenderer_b200/lib/alt_NAME.so (select with JR_B200_LIB=...)
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
csrc="$here/jaxrenderer_b200/csrc"
out="$here/jaxrenderer_b200/lib/alt_$1.so"
srcs=("$csrc"/jr_common.cu "$csrc"/jr_forward.cu "$csrc"/jr_camera.cu "$csrc"/jr_backward.cu)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
  -diag-suppress 128 -Xcompiler -fPIC -shared $2 -o "$out" "${srcs[@]}"
echo "built $out"
