#!/bin/sh
# Compiles oracle/_ref/libhop_ref.so from the reference tree's sources WHERE THEY LIE (read-only).
# The OpenGR fork has one defect that makes it unusable as compiled (KdTree::operator= is declared bool but has
# no return statement: src/OpenGR_4pcs/src/gr/accelerators/kdtree.h:148-156 -> g++ falls through / traps).
# A one-line-patched copy of THAT header is generated into a throw-away directory that shadows it on the
# include path for this compilation only; it is deleted afterwards and never enters the repository.
set -e
REF="${1:-/root/reference}"
CXX="${2:-g++}"
HERE="$(cd "$(dirname "$0")" && pwd)"
EIGEN="$REF/src/OpenGR_4pcs/3rdparty/Eigen"
GR="$REF/src/OpenGR_4pcs/src"
mkdir -p "$HERE/_ref"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
SRCS="$HERE/ref_eigen_lm.cpp"
INCS="-I$EIGEN"
if [ -f "$HERE/ref_opengr.cpp" ]; then
  mkdir -p "$TMP/gr/accelerators"
  # insert "return true;" before the closing brace of operator= (first "}" at 2-space indent after the signature)
  awk 'BEGIN{s=0} /bool operator= *\(/{s=1} {if(s==1 && $0 ~ /^[ \t]*}[ \t]*$/){print "    return true;"; s=2} print}' \
      "$GR/gr/accelerators/kdtree.h" > "$TMP/gr/accelerators/kdtree.h"
  SRCS="$SRCS $HERE/ref_opengr.cpp"
  INCS="-I$TMP -I$GR $INCS"
fi
if [ -f "$HERE/ref_cluster.cpp" ]; then SRCS="$SRCS $HERE/ref_cluster.cpp"; fi
if [ -f "$HERE/ref_umeyama.cpp" ]; then SRCS="$SRCS $HERE/ref_umeyama.cpp"; fi
# the reference's vendored libigl (header-only use): signed distance to a mesh, what SDFchecker calls
if [ -f "$HERE/ref_sdf.cpp" ] && [ -f "$REF/src/perception/include/igl/signed_distance.h" ]; then
  SRCS="$SRCS $HERE/ref_sdf.cpp"
  INCS="$INCS -I$REF/src/perception/include"
fi
$CXX -std=c++17 -O2 -fopenmp -fPIC -shared -w $INCS -o "$HERE/_ref/libhop_ref.so" $SRCS
echo "oracle: built $HERE/_ref/libhop_ref.so"
