#!/usr/bin/env bash
# Build the walexi/gnn.cpp reference (CPU, single thread) into oracle/_ref/.
#
# TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path.
#
# The reference HEAD does not compile.  This recipe never
# copies reference sources into the repo: it stages a scratch copy under a
# temp dir, applies six mechanical, semantics-preserving compile fixes with
# sed, compiles them together with OUR driver (oracle/ref_driver.cpp) and
# writes only the resulting binary to oracle/_ref/ (git-ignored, ships to the
# GPU box with the snapshot).
#
# usage: oracle/build_ref.sh [/root/reference]
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/include" ]; then
  echo "build_ref: reference not present at $REF (expected on the GPU box); keeping prebuilt $OUT" >&2
  exit 0
fi
TMP="$(mktemp -d /tmp/gnnref.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF/include" "$REF/src" "$TMP/"
cd "$TMP"
# (1) wrong header name                      include/operation.h:9, src/nn.cpp:6
sed -i 's/#include "util.h"/#include "utils.h"/' include/operation.h src/nn.cpp
# (2) missing standard headers               include/utils.h:3
sed -i '0,/#include <valarray>/s//#include <memory>\n#include <climits>\n#include <ctime>\n#include <tuple>\n#include <stdexcept>\n#include <valarray>/' include/utils.h
# (3) generate_random<T> is called as a template but declared as a plain function   include/utils.h:94
sed -i 's/^float generate_random(const float &low, const float &high);/float generate_random(const float \&low, const float \&high);\ntemplate <class T> T generate_random(const float \&low, const float \&high) { return static_cast<T>(generate_random(low, high)); }/' include/utils.h
# (4) stale non-template randn contradicting tensor.h:864      src/tensor.cpp:15-24
sed -i '15,24d' src/tensor.cpp
# (5) nn::Embedding defined but never declared                 src/nn.cpp:455-461
sed -i '455,461d' src/nn.cpp
# (6) unqualified names / duplicate declaration                include/graph.h:132, src/graph.cpp:155
sed -i '132s/const tensor<int>/const cyg::tensor<int>/; 132s/const tptr<float>/const cyg::tptr<float>/g' include/graph.h
sed -i '155s/auto out = message/auto out_msg = message/' src/graph.cpp
mkdir -p "$OUT"
CXX="${ORACLE_CXX:-/usr/bin/g++}"
FLAGS="-std=c++20 -O2 -fpermissive -w -Iinclude -Isrc"
$CXX $FLAGS -c src/utils.cpp  -o utils.o
$CXX $FLAGS -c src/tensor.cpp -o tensor.o
$CXX $FLAGS -c src/graph.cpp  -o graph.o
# nn.cpp is included textually by the driver: cross_entropy_loss/tanh/sigmoid are
# `inline` there (src/nn.cpp:355,366,442) and not linkable from another TU.
$CXX $FLAGS -DREF_NN_CPP="\"$TMP/src/nn.cpp\"" "$HERE/ref_driver.cpp" utils.o tensor.o graph.o -o "$OUT/ref_gcn"
echo "built $OUT/ref_gcn"
