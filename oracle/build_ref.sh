#!/usr/bin/env bash
# Test infrastructure only (never on the product path).
#
# Compiles the UNMODIFIED reference (steve-the-bayesian/BOOM) from the sources
# where they lie under $BOOM_REF (default /root/reference) into a static
# archive build/boomref/libboom_ref.a.  Nothing is copied into this repo; the
# objects live under build/ (git-ignored AND gpurun-ignored), the linked
# artefacts that tests/bench use go to oracle/_ref/ (git-ignored, travels to
# the GPU box).  Recipe follows SURVEY.md §8(c1): 468 TUs, g++ -O2, no bazel.
set -euo pipefail
REF=${BOOM_REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/.." && pwd)
OBJ=$ROOT/build/boomref/obj
LIB=$ROOT/build/boomref/libboom_ref.a
JOBS=${JOBS:-$(nproc)}
if [ ! -d "$REF" ]; then
  echo "build_ref.sh: $REF not present (GPU box?) - using prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OBJ"
cd "$REF"
ls Bmath/*.cpp LinAlg/*.cpp distributions/*.cpp cpputil/*.cpp math/*.cpp \
   math/cephes/*.cpp numopt/*.cpp stats/*.cpp Samplers/*.cpp \
   Samplers/Gilks/arms.cpp TargetFun/*.cpp Models/*.cpp Models/Policies/*.cpp \
   Models/PosteriorSamplers/*.cpp Models/Glm/*.cpp \
   Models/Glm/PosteriorSamplers/*.cpp test_utils/*.cpp > "$OBJ/../tus.txt"
CXXFLAGS="-std=c++17 -O2 -fPIC -w -I$REF -I$REF/Bmath -I$REF/math/cephes -DADD_ -DNDEBUG"
compile_one() {
  src=$1
  obj="$OBJ/$(echo "$src" | tr '/' '_' | sed 's/\.cpp$/.o/')"
  if [ ! -f "$obj" ] || [ "$REF/$src" -nt "$obj" ]; then
    g++ $CXXFLAGS -c "$REF/$src" -o "$obj" || { echo "FAILED $src" >&2; exit 1; }
  fi
}
export -f compile_one
export OBJ REF CXXFLAGS
xargs -P "$JOBS" -I{} bash -c 'compile_one {}' < "$OBJ/../tus.txt"
rm -f "$LIB"
ar rcs "$LIB" "$OBJ"/*.o
echo "built $LIB ($(ls "$OBJ"/*.o | wc -l) objects)"
