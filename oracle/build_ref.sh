#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Builds the reference's own CPU implementation of the MAS preconditioner
# (SeSchwarzPreconditioner.cpp + SeOmp.cpp, read from where they lie under
# $MAS_REFERENCE_DIR, default /root/reference) into oracle/_ref/:
#   libmas_ref.so        the reference as shipped
#   libmas_prev.so       SeSchwarzPreconditionerPreviousVersion.h (the older variant), z only
#   libmas_ref_q5fix.so  same as libmas_ref.so, with the four-line cull at cpp:991-994 removed
#                        (PrefixSumLx truncates its cross-block
#                        prefix once a level has >33,792 nodes; needed for the
#                        4.2M-vertex config only)
# No reference source is written into the repository: the reference is
# MSVC-only as shipped, so a scratch copy under mktemp gets six mechanical
# sed edits to SUPPORT headers only (alignment attribute, `static` on explicit
# specialisations, __m128::m128_f32, two anonymous-union members, one #pragma);
# the algorithm file is compiled untouched with -DWIN32 so that its real AVX2
# LDLtInverse512/SchwarzLocalXSym bodies are the ones that run.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref="${MAS_REFERENCE_DIR:-/root/reference}"
out="$here/_ref"
if [ ! -f "$ref/SeSchwarzPreconditioner.cpp" ]; then
	echo "build_ref.sh: reference not present at $ref; keeping prebuilt $out" >&2
	exit 0
fi
mkdir -p "$out"
tmp="$(mktemp -d)"
trap 'rm -rf "$tmp"' EXIT
cp "$ref"/*.h "$ref"/*.cpp "$tmp"/
chmod -R u+w "$tmp"
sed -i 's/__declspec(align(n))/alignas(n)/' "$tmp/SePreDefine.h"
sed -i -E 's/^static[[:space:]]+(SE_INLINE[[:space:]]+)?bool (IsContain|IsIntersect|IsOverlap)/\1bool \2/' "$tmp/SeAabb.h" "$tmp/SeAabbSimd.h"
sed -i 's/pack\.m128_f32\[i\]/((float*)\&pack)[i]/g' "$tmp/SeVectorSimd.h"
sed -i '/struct { SeVector3<float> xyz; };/d; /struct { SeVector4<float> xyzw; };/d' "$tmp/SeVectorSimd.h"
sed -i '/#pragma intrinsic(_BitScanForward)/d' "$tmp/SeIntrinsic.h"
mkdir -p "$tmp/shim"
: > "$tmp/shim/intrin.h"
cp "$here/ref_shim/msvc_shim.h" "$tmp/shim/"
cp "$here/ref_harness.cpp" "$here/prev_harness.cpp" "$tmp/"

flags=(-std=c++20 -O2 -DNDEBUG -fopenmp -mavx2 -mfma -mlzcnt -mpopcnt -fpermissive -w -DWIN32 -fPIC
	-I"$tmp/shim" -I"$tmp" -include msvc_shim.h)

g++ "${flags[@]}" -shared -o "$out/libmas_ref.so" \
	"$tmp/SeSchwarzPreconditioner.cpp" "$tmp/SeOmp.cpp" "$tmp/ref_harness.cpp"

# Q5-fixed variant: delete the cull `if (vid >= (levelNum + blockDim - 1) / blockDim * blockDim) { break; }`
# inside PrefixSumLx's cross-block prefix loop (cpp:991-994) in the scratch copy.
python3 - "$tmp/SeSchwarzPreconditioner.cpp" "$tmp/SeSchwarzPreconditioner_q5.cpp" <<'EOF'
import sys, re
src = open(sys.argv[1], encoding="latin-1").read()
pat = re.compile(r"if \(vid >= \(levelNum \+ blockDim - 1\) / blockDim \* blockDim\)\s*\{\s*break;\s*\}")
new, n = pat.subn("", src)
assert n == 1, f"expected exactly one Q5 cull, found {n}"
open(sys.argv[2], "w", encoding="latin-1").write(new)
EOF
g++ "${flags[@]}" -shared -o "$out/libmas_ref_q5fix.so" \
	"$tmp/SeSchwarzPreconditioner_q5.cpp" "$tmp/SeOmp.cpp" "$tmp/ref_harness.cpp"
# the older variant (header-only class SeSchwarzPreconditionerPreviousVersion) as a second cross-check of z
g++ "${flags[@]}" -shared -o "$out/libmas_prev.so" "$tmp/SeOmp.cpp" "$tmp/prev_harness.cpp"
echo "built $out/libmas_ref.so $out/libmas_ref_q5fix.so $out/libmas_prev.so"
