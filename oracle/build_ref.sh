#!/usr/bin/env bash
# Build oracle/_ref/libvslam_ref.so = the reference's OWN tracking front-end sources
# (/root/reference/jni, read where they lie) + oracle/ref_harness.cc (a C-ABI around them).
#
# TEST INFRASTRUCTURE ONLY.  The reference hard-codes absolute Eigen include paths
# ("/Users/ahcorde/Downloads/eigen/Eigen/Dense", "/opt/local/include/eigen3/Eigen/Dense") and
# needs OpenCV 2.4 + Eigen headers that this image does not have, so the recipe
#   1. streams each needed reference file through `sed` (ONLY the #include lines that name
#      those absolute paths are rewritten to <Eigen/Dense>) into a throw-away directory under
#      $TMPDIR -- nothing from /root/reference is ever written into the repository;
#   2. compiles them against oracle/shim/{Eigen,opencv2,opencv} (value-semantics stand-ins);
#   3. links with oracle/ref_harness.cc, which exports ref_* entry points (MapMaker.cc, Bundle.cc and HomographyInit.cc are the
#      reference's; only the never-started map-maker THREAD is stood in for);
#   4. writes ONLY the shared object into oracle/_ref/ (git-ignored, travels with gpurun).
# Without /root/reference (e.g. on the GPU box) the script is a no-op that keeps a prebuilt .so.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${VSLAM_REFERENCE_DIR:-/root/reference}/jni"
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ ! -d "$REF" ]; then
  echo "build_ref: $REF not present; keeping prebuilt $OUT/libvslam_ref.so (if any)"; exit 0
fi
if [ "$OUT/libvslam_ref.so" -nt "$HERE/ref_harness.cc" ] && [ "$OUT/libvslam_ref.so" -nt "$HERE/shim/Eigen/Dense" ] \
   && [ "$OUT/libvslam_ref.so" -nt "$HERE/shim/opencv2/core/core.hpp" ] && [ "$OUT/libvslam_ref.so" -nt "$HERE/build_ref.sh" ] \
   && [ "$OUT/libvslam_ref.so" -nt "$HERE/shim/cv_resize_linear_u8.h" ]; then
  echo "build_ref: up to date"; exit 0
fi
TMP="$(mktemp -d "${TMPDIR:-/tmp}/vslam_ref.XXXXXX")"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/vision"
FILES="ATANCamera.h ATANCamera.cc Bundle.h Bundle.cc HomographyInit.h HomographyInit.cc KeyFrame.h KeyFrame.cc LevelHelpers.h MEstimator.h Map.h Map.cc MapMaker.h MapMaker.cc MapPoint.h MapPoint.cc
MiniPatch.h MiniPatch.cc PatchFinder.h PatchFinder.cc RT.h Relocaliser.h Relocaliser.cc SmallBlurryImage.h SmallBlurryImage.cc
Tracker.h Tracker.cc TrackerData.h myWLS.h vision/ImageHandler.h vision/ImageHandler.cpp vision/cvfast.h vision/cvfast.cpp"
for f in $FILES; do
  sed -E 's@#include "(/Users/ahcorde/Downloads/eigen|/opt/local/include/eigen3)/Eigen/Dense"@#include <Eigen/Dense>@' "$REF/$f" > "$TMP/$f"
done
CXX="${VSLAM_CXX:-/usr/bin/g++}"   # (the image exports CXX=/opt/gcc/bin/g++, which links libstdc++ statically: iostreams of a dlopen-ed library then crash)
# -O3 as in jni/Application.mk; -ffp-contract=off: no FMA contraction (the reference's ARMv7/x86 builds have none);
# gnu++11 keeps std::random_shuffle / std::binary_function; -Wno-narrowing for cvfast.cpp's size_t->int ring offsets.
NDBG="-DNDEBUG"; [ -n "${VSLAM_REF_DEBUG:-}" ] && NDBG=""   # NDK release builds define NDEBUG
FLAGS="-std=gnu++11 -O3 -fPIC -ffp-contract=off -fno-fast-math $NDBG -w -Wno-narrowing -I$HERE/shim -I$TMP"
OBJS=""
for f in $FILES; do
  case "$f" in *.cc|*.cpp)
    o="$TMP/$(echo "$f" | tr '/.' '__').o"
    $CXX $FLAGS -c "$TMP/$f" -o "$o" &
    OBJS="$OBJS $o";;
  esac
done
$CXX $FLAGS -c "$HERE/ref_harness.cc" -o "$TMP/ref_harness.o" &
FAIL=0; for p in $(jobs -p); do wait "$p" || FAIL=1; done; [ "$FAIL" = 0 ] || { echo "build_ref: compile failed"; exit 1; }
$CXX -shared -o "$OUT/libvslam_ref.so" $OBJS "$TMP/ref_harness.o" -lpthread
echo "build_ref: wrote $OUT/libvslam_ref.so"
