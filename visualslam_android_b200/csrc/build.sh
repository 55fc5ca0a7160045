#!/usr/bin/env bash
# Build visualslam_android_b200/libvslam_b200.so (CUDA kernels + C-ABI) for sm_100a, in-tree.
# -fmad=false: the FP64 geometry must not contract a*b+c, so that it reproduces the reference's x86/ARM arithmetic.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libvslam_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
newest=$(ls -t "$HERE"/*.cu "$HERE"/*.cuh "$HERE/../../include/vslam_b200.h" "$HERE/build.sh" | head -1)
if [ -f "$OUT" ] && [ "$OUT" -nt "$newest" ]; then echo "build: libvslam_b200.so up to date"; exit 0; fi
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -I$HERE/../../include -I$HERE ${VSLAM_NVCC_EXTRA:-}"
mkdir -p "$HERE/_build"
pids=()
for f in pyramid_fast track pose_fast search_fast patchfinder_ops keyframe_rest sbi mapfile api; do
  $NVCC $FLAGS -c "$HERE/$f.cu" -o "$HERE/_build/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$HERE/_build/pyramid_fast.o" "$HERE/_build/track.o" "$HERE/_build/pose_fast.o" "$HERE/_build/search_fast.o" "$HERE/_build/patchfinder_ops.o" "$HERE/_build/keyframe_rest.o" "$HERE/_build/sbi.o" "$HERE/_build/mapfile.o" "$HERE/_build/api.o"
echo "build: wrote $OUT"
