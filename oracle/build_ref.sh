#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.
#
# Builds the UNMODIFIED reference (jqly/VoxelRayTrace20190722) hot-path sources
# together with oracle/ref_harness.cc into oracle/_ref/libvrt_ref.so.
#
#   * never runs the reference's own build system (it is a VS2017 .vcxproj);
#   * never writes to /root/reference and never copies reference sources into
#     the repo: the two build-only accommodations of SURVEY.md 8(c) are applied
#     to a throw-away copy under $(mktemp -d), which is deleted afterwards;
#   * the only output is oracle/_ref/libvrt_ref.so (+ a BUILD_INFO text file).
#
# Accommodations (neither touches arithmetic):
#   1. graphics_math.h pastes `operator##Op` (MSVC-only token paste; g++ hard
#      error).  sed 's/operator##Op/operator Op/g' on the copy.
#   2. libstdc++ 13 has no std::sqrtf/tanf/...; a force-included shim header
#      injects `using ::sqrtf;` etc. into namespace std.
# Flags: -std=c++17 -O2 -ffp-contract=off == MSVC x64 /O2 /fp:precise without
# FMA contraction (VoxelRayTrace20190722.vcxproj:91-120).
set -euo pipefail

HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${VRT_REFERENCE_DIR:-/root/reference}/VoxelRayTrace20190722"
OUT="$HERE/_ref"

if [ ! -f "$REF/voxel_octree.cc" ]; then
        echo "build_ref.sh: reference sources not found at $REF (fine on the GPU box: prebuilt oracle/_ref is used)" >&2
        exit 3
fi

TMP="$(mktemp -d /tmp/vrt_ref_build.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT

mkdir -p "$TMP/src" "$OUT"
for f in camera.cc camera.h voxel_octree.cc voxel_octree.h tribox2.cc tribox2.h \
         raytri.cc raytri.h tiny_obj_loader.cc tiny_obj_loader.h graphics_math.h \
         util.h stb_image.h; do
        cp "$REF/$f" "$TMP/src/$f"
done
cp -r "$REF/thread_pool_cpp" "$TMP/src/thread_pool_cpp"

# accommodation 1
sed -i 's/operator##Op/operator Op/g' "$TMP/src/graphics_math.h"

# accommodation 2
cat > "$TMP/libm_shim.h" <<'EOF'
#ifdef __cplusplus
#include <cmath>
#include <math.h>
namespace std {
using ::sqrtf; using ::sinf; using ::cosf; using ::tanf;
using ::expf; using ::powf; using ::log2f;
}
#endif
EOF

CXX="${CXX:-g++}"
CXXFLAGS="-std=c++17 -O2 -ffp-contract=off -fPIC -w -include $TMP/libm_shim.h -I$TMP/src"

pids=()
for f in camera voxel_octree tribox2 raytri tiny_obj_loader; do
        $CXX $CXXFLAGS -c "$TMP/src/$f.cc" -o "$TMP/$f.o" &
        pids+=($!)
done
$CXX $CXXFLAGS -c "$HERE/ref_harness.cc" -o "$TMP/ref_harness.o" &
pids+=($!)
for p in "${pids[@]}"; do
        wait "$p"  # (set -e: a failed compile stops the build instead of linking a partial library)
done

$CXX -shared -o "$OUT/libvrt_ref.so" "$TMP"/*.o -lpthread

{
        echo "built: $(date -u +%Y-%m-%dT%H:%M:%SZ)"
        echo "compiler: $($CXX --version | head -1)"
        echo "flags: -std=c++17 -O2 -ffp-contract=off"
        echo "reference: $REF"
        (cd "$REF" && sha256sum camera.cc voxel_octree.cc tribox2.cc raytri.cc graphics_math.h camera.h voxel_octree.h)
} > "$OUT/BUILD_INFO.txt"
echo "built $OUT/libvrt_ref.so"
