#!/usr/bin/env bash
# Build oracle/libvslam_oracle.so from oracle/vslam_oracle.cc (the CPU restatement; test infrastructure only).
# -ffp-contract=off: keep the reference's non-fused double arithmetic.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
if [ "$HERE/libvslam_oracle.so" -nt "$HERE/vslam_oracle.cc" ] && [ "$HERE/libvslam_oracle.so" -nt "$HERE/build_oracle.sh" ] && [ "$HERE/libvslam_oracle.so" -nt "$HERE/shim/cv_resize_linear_u8.h" ]; then
  echo "build_oracle: up to date"; exit 0
fi
${VSLAM_CXX:-/usr/bin/g++} -std=gnu++11 -O3 -fPIC -ffp-contract=off -fno-fast-math -Wall -Wno-unused-function -shared \
  "$HERE/vslam_oracle.cc" -o "$HERE/libvslam_oracle.so"
echo "build_oracle: wrote $HERE/libvslam_oracle.so"
