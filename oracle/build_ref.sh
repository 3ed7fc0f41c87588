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
         util.h stb_image.h stb_image_write.h; do
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

# ---------------------------------------------------------------------------------------------
# DROP-IN link test (north star: "main.cc links against the GPU path as a drop-in").
# The reference's OWN main.cc + voxel_octree.h + camera.h + voxel_octree.cc + camera.cc, from the
# same throw-away copy, linked against libvrt.so through voxelraytrace20190722_b200/cpp/vrt_dropin.cc:
#   * tribox2.cc and raytri.cc are NOT compiled: triBoxOverlap / intersect_triangle3 come from the binding;
#   * voxel_octree.cc is compiled with -Dray_march_init=... -Dray_march=... (preprocessor renames, no edit),
#     so its CPU build/traversal keep other names and gi::ray_march_init / gi::ray_march resolve to the
#     binding (CUDA); obj2voxel, Triangle, texel_fetch, cone_trace stay the reference's own host code;
#   * camera.cc is compiled with -DCamera=RefCpuCamera: Film stays the reference's, Camera is the binding's;
#   * main.cc: ONLY the hard-coded Windows path of sponza.obj (main.cc:47) is replaced by ./dropin_scene.obj.
#     main_dropin_small additionally shrinks the two films (main.cc:34-35,75) so that the per-ray launches of
#     the test finish in seconds -- the geometry, cameras, depth and every call site are untouched.
# Outputs: oracle/_ref/main_dropin, oracle/_ref/main_dropin_small (git-ignored, shipped by gpurun).
# ---------------------------------------------------------------------------------------------
PKG="$HERE/../voxelraytrace20190722_b200"
if [ "${VRT_SKIP_DROPIN:-0}" != "1" ] && [ -f "$PKG/libvrt.so" ]; then
        for f in main.cc stb_image_write.h; do cp "$REF/$f" "$TMP/src/$f"; done
        grep -q 'sponza.obj"' "$TMP/src/main.cc"
        sed -i 's#^\(\s*\)"C:.*sponza\.obj",#\1"./dropin_scene.obj",#' "$TMP/src/main.cc"
        grep -q '"./dropin_scene.obj",' "$TMP/src/main.cc"
        sed -e 's/int W = 1024;/int W = 96;/' -e 's/int H = 1024;/int H = 96;/' \
            -e 's/Film sfilm(1, 1, 2048, 2048);/Film sfilm(1, 1, 128, 128);/' "$TMP/src/main.cc" > "$TMP/src/main_small.cc"
        grep -q 'int W = 96;' "$TMP/src/main_small.cc" && grep -q 'Film sfilm(1, 1, 128, 128);' "$TMP/src/main_small.cc"
        DFLAGS="-std=c++17 -O2 -ffp-contract=off -w -include $TMP/libm_shim.h -I$TMP/src -I$HERE/../include"
        $CXX $DFLAGS -Dray_march_init=ref_cpu_ray_march_init -Dray_march=ref_cpu_ray_march -c "$TMP/src/voxel_octree.cc" -o "$TMP/d_voxel_octree.o" &
        p1=$!
        $CXX $DFLAGS -DCamera=RefCpuCamera -c "$TMP/src/camera.cc" -o "$TMP/d_camera.o" &
        p2=$!
        $CXX $DFLAGS -c "$PKG/cpp/vrt_dropin.cc" -o "$TMP/d_binding.o" &
        p3=$!
        $CXX $DFLAGS -c "$TMP/src/main.cc" -o "$TMP/d_main.o" &
        p4=$!
        $CXX $DFLAGS -c "$TMP/src/main_small.cc" -o "$TMP/d_main_small.o" &
        p5=$!
        for p in $p1 $p2 $p3 $p4 $p5; do wait "$p"; done
        for m in main main_small; do
                $CXX -o "$OUT/${m/main/main_dropin}" "$TMP/d_$m.o" "$TMP/d_voxel_octree.o" "$TMP/d_camera.o" "$TMP/d_binding.o" \
                     "$TMP/tiny_obj_loader.o" -L"$PKG" -lvrt -lpthread -Wl,-rpath,'$ORIGIN/../../voxelraytrace20190722_b200'
        done
        # the same small main.cc linked against the reference's own objects only (CPU): the checker of the drop-in test
        $CXX -o "$OUT/main_refcpu_small" "$TMP/d_main_small.o" "$TMP/camera.o" "$TMP/voxel_octree.o" "$TMP/tribox2.o" \
             "$TMP/raytri.o" "$TMP/tiny_obj_loader.o" -lpthread
        # the hot symbols of the drop-in binaries must come from the binding, the reference's CPU versions keep their renamed symbols
        nm -C "$OUT/main_dropin" > "$TMP/syms.txt"
        grep -q ' T gi::ray_march(gi::VoxelOctree' "$TMP/syms.txt"
        grep -q 'gi::ref_cpu_ray_march(' "$TMP/syms.txt"
        grep -q ' T triBoxOverlap' "$TMP/syms.txt"
        grep -q ' T Camera::gen_rays4' "$TMP/syms.txt"
        echo "built $OUT/main_dropin and $OUT/main_dropin_small (reference main.cc linked against libvrt.so)"
fi
