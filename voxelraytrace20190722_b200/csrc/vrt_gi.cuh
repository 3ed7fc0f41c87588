// vrt_gi.cuh -- device side of the GI rows: VoxelOctree::compute_illum
// (voxel_octree.h:71-81), cone_trace / orthonormal_basis / HemiCones (voxel_octree.cc:216-303)
// on the flat node array.  Per-node GI state lives in a side array of kGiStride floats per
// node: six float4 {illum r, g, b, coverage}, one per axis lobe (96 B; a cone sample reads three of them).
// Arithmetic follows the reference expression by expression (vrt_exact.cuh rules: no FMA,
// IEEE division and sqrt).  One deviation, documented in DESIGN.md: int(std::log2f(x)) is
// evaluated for the correctly rounded log2f (gi_split_level) -- the host libm's log2f is not
// specified to the last bit, so the integer `split_level` can differ from a given libm when
// maxdist/diam is one of the few floats just below a power of two.
#pragma once

#include "vrt_exact.cuh"
#include "vrt_internal.h"

namespace vrt {


// Triangle::get_albedo (voxel_octree.cc:471-484): the material's diffuse colour, or -- for a
// textured material -- the nearest texel at the clamped barycentric interpolation of the vertex
// texture coordinates: jql::barycentric (graphics_math.h:1082-1100), unit_cycle + texel_fetch with
// the vertical flip (voxel_octree.cc:392-422).  kd_default is used when no materials are set.
__device__ __forceinline__ float gi_unit_cycle(float s)
{
        // while (s > 1) s -= 1; while (s < 0) s += 1;  -- the reference does not terminate for |s| >= 2^24
        // (s -+ 1 == s); such coordinates are cut off after 2^24 steps here
        for (int i = 0; i < (1 << 24) && s > 1.f; ++i)
                s = fsub(s, 1.f);
        for (int i = 0; i < (1 << 24) && s < 0.f; ++i)
                s = fadd(s, 1.f);
        return s;
}

__device__ __forceinline__ void gi_albedo(const TreeDev& tr, uint32_t tri, const float hit[3], const float kd_default[3],
                                          float out[3])
{
        if (!tr.mat_tri) {
                out[0] = kd_default[0];
                out[1] = kd_default[1];
                out[2] = kd_default[2];
                return;
        }
        const float4 m = __ldg(&tr.mat_kd[__ldg(&tr.mat_tri[tri])]);
        const int tx = __float_as_int(m.w);
        if (tx < 0) {
                out[0] = m.x;
                out[1] = m.y;
                out[2] = m.z;
                return;
        }
        const float4 a = __ldg(&tr.tri4[3ull * tri]), b = __ldg(&tr.tri4[3ull * tri + 1]), c = __ldg(&tr.tri4[3ull * tri + 2]);
        const float v0[3] = { fsub(b.x, a.x), fsub(b.y, a.y), fsub(b.z, a.z) };
        const float v1[3] = { fsub(c.x, a.x), fsub(c.y, a.y), fsub(c.z, a.z) };
        const float v2[3] = { fsub(hit[0], a.x), fsub(hit[1], a.y), fsub(hit[2], a.z) };
        const float d00 = dot3(v0[0], v0[1], v0[2], v0[0], v0[1], v0[2]);
        const float d01 = dot3(v0[0], v0[1], v0[2], v1[0], v1[1], v1[2]);
        const float d11 = dot3(v1[0], v1[1], v1[2], v1[0], v1[1], v1[2]);
        const float d20 = dot3(v2[0], v2[1], v2[2], v0[0], v0[1], v0[2]);
        const float d21 = dot3(v2[0], v2[1], v2[2], v1[0], v1[1], v1[2]);
        const float denom = fsub(fmul(d00, d11), fmul(d01, d01));
        float bc[3] = { 0.f, 0.f, 0.f };
        if (denom != 0.f) {
                bc[1] = fdiv(fsub(fmul(d11, d20), fmul(d01, d21)), denom);
                bc[2] = fdiv(fsub(fmul(d00, d21), fmul(d01, d20)), denom);
                bc[0] = fsub(fsub(1.0f, bc[1]), bc[2]);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
                bc[k] = clampf(bc[k], 0.f, 1.f);
        const float2 t0 = __ldg(&tr.mat_uv[3ull * tri]), t1 = __ldg(&tr.mat_uv[3ull * tri + 1]), t2 = __ldg(&tr.mat_uv[3ull * tri + 2]);
        const float tcx = fadd(fadd(fmul(bc[0], t0.x), fmul(bc[1], t1.x)), fmul(bc[2], t2.x));
        const float tcy = fadd(fadd(fmul(bc[0], t0.y), fmul(bc[1], t1.y)), fmul(bc[2], t2.y));
        const int4 td = __ldg(&tr.mat_tex[tx]);  // byte offset, width, height, channels
        int x = (int)fmul(gi_unit_cycle(tcx), (float)td.y);
        int y = (int)fmul(gi_unit_cycle(tcy), (float)td.z);
        x = x > td.y - 1 ? td.y - 1 : (x < 0 ? 0 : x);
        y = y > td.z - 1 ? td.z - 1 : (y < 0 ? 0 : y);
        y = td.z - 1 - y;
        const uint8_t* px = tr.mat_texels + (size_t)(uint32_t)td.x + ((size_t)y * td.y + x) * td.w;
        float pixel[3] = { 0.f, 0.f, 0.f };
#pragma unroll
        for (int k = 0; k < 3; ++k)
                if (k < td.w)
                        pixel[k] = (float)__ldg(px + k);
#pragma unroll
        for (int k = 0; k < 3; ++k)
                out[k] = fdiv(pixel[k], 255.f);
}

// VoxelOctree::compute_illum(d): sum_i clamp(dot(illum_d[i], d), 0, 1) * illum[i]
__device__ __forceinline__ void gi_compute_illum(const float* __restrict__ g, const float d[3], float out[3])
{
        const float4* g4 = reinterpret_cast<const float4*>(g);
        const float4 q0 = __ldg(g4), q1 = __ldg(g4 + 1), q2 = __ldg(g4 + 2), q3 = __ldg(g4 + 3), q4 = __ldg(g4 + 4),
                     q5 = __ldg(g4 + 5);
        const float il[18] = { q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, q2.x, q2.y, q2.z,
                               q3.x, q3.y, q3.z, q4.x, q4.y, q4.z, q5.x, q5.y, q5.z };
        out[0] = out[1] = out[2] = 0.f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
                // illum_d[] = +x +y +z -x -y -z (voxel_octree.cc:19-20); jql::dot keeps the zero terms
                const float s = (i < 3) ? 1.f : -1.f;
                const float ax = (i % 3 == 0) ? s : 0.f, ay = (i % 3 == 1) ? s : 0.f, az = (i % 3 == 2) ? s : 0.f;
                const float coeff = clampf(dot3(ax, ay, az, d[0], d[1], d[2]), 0.f, 1.f);
#pragma unroll
                for (int k = 0; k < 3; ++k)
                        out[k] = fadd(out[k], fmul(coeff, il[3 * i + k]));
        }
}

// int(std::log2f(x)) for a finite x >= 1 with log2f ROUNDED TO NEAREST: the exponent k of x,
// plus one when x is one of the last n floats below 2^(k+1) whose logarithm rounds up to k+1.
// log2(x) >= (k+1) - h with h = half the float spacing just below k+1 = 2^(j-24), j = floor(log2(k))
// (k >= 1), holds for the top floor(2^24 * h * ln 2) = floor(2^j * ln 2) = 0,1,2,5,11,22,.. floats of the binade
// (tests/test_abi_host.py checks this closed form against float(log2(double(x))) around every
// binade boundary).  Integer-only: the cone trace evaluates it once per step.
__device__ __forceinline__ int gi_split_level(float x)
{
        const uint32_t b = __float_as_uint(x);
        const int k = (int)(b >> 23) - 127;
        const int j = 31 - __clz(k | 1);
        const uint32_t n = (uint32_t)((0x2c160b05020100ull >> (8 * j)) & 0xffull);  // floor(2^j ln 2), j < 7
        return k + (((b & 0x7fffffu) + n >= 0x800000u) ? 1 : 0);
}

// ---------------------------------------------------------------------------
// Point location with a cached path.  The reference descends from the root at every step of a
// cone (voxel_octree.cc:261-271): child i = (p.x > c.x ? 4:0) + (p.y > c.y ? 2:0) + (p.z > c.z ? 1:0)
// with c the node's centre.  Consecutive steps move p by a tenth of the sampled cell, so the path
// rarely changes.  We keep ONE cached path per thread: for every level l of it the cumulative
// per-axis interval (LO_l, HI_l] of coordinates that reproduce the first l choices exactly --
// LO = max of the centres passed with "p > c" true, HI = min of the centres passed with it false
// -- and the node reached.  A query at depth s is then six compares against level min(s, len);
// only when they fail (or the path is too short) are nodes and table entries loaded again, from
// the deepest level that still holds.  The decisions are the reference's own comparisons, so the
// located node is identical.  Cache record (7 words) per level 1..L in shared memory,
// [level][word][thread].
// ---------------------------------------------------------------------------
constexpr int kGiPathWords = 7;
constexpr uint32_t kGiAbsent = 0xffffffffu;

struct GiPath {
        uint32_t x, y, z;  // heap coordinates of the deepest cached level
        int len;           // number of cached choices (levels 1..len are valid)
};

__device__ __forceinline__ void gi_path_reset(GiPath& gp)
{
        gp.x = gp.y = gp.z = 1u;
        gp.len = 0;
}

// Node of level s (1 <= s <= L) that the reference's descent reaches for p, or kGiAbsent when the
// descent ends earlier in one of the reference's empty leaves (an absent child).
__device__ __forceinline__ uint32_t gi_locate(const TreeDev& tr, float* __restrict__ sc, int stride, GiPath& gp,
                                              const float p[3], int s)
{
        const float inf = __int_as_float(0x7f800000);
        int l = min(s, gp.len);
        float lo[3] = { -inf, -inf, -inf }, hi[3] = { inf, inf, inf };
        uint32_t node = 0;
        while (l > 0) {
                const float* e = sc + (size_t)(kGiPathWords * (l - 1)) * stride;
                lo[0] = e[0];
                lo[1] = e[stride];
                lo[2] = e[2 * stride];
                hi[0] = e[3 * stride];
                hi[1] = e[4 * stride];
                hi[2] = e[5 * stride];
                if (p[0] > lo[0] && p[0] <= hi[0] && p[1] > lo[1] && p[1] <= hi[1] && p[2] > lo[2] && p[2] <= hi[2]) {
                        node = __float_as_uint(e[6 * stride]);
                        break;
                }
                --l;
        }
        if (l == s)
                return node;
        if (l == 0) {
                lo[0] = lo[1] = lo[2] = -inf;
                hi[0] = hi[1] = hi[2] = inf;
                node = 0;
        }
        if (node == kGiAbsent)
                return kGiAbsent;  // the valid part of the path already ends in an empty leaf
        // re-descend from level l with the reference's loads and comparisons, refreshing the cache
        uint32_t x = gp.x >> (gp.len - l), y = gp.y >> (gp.len - l), z = gp.z >> (gp.len - l);
        while (l < s) {
                const uint2 rec = __ldg(&tr.nodes[node]);
                const float2 bx = __ldg(&tr.tab2[0][x]), by = __ldg(&tr.tab2[1][y]), bz = __ldg(&tr.tab2[2][z]);
                const float c[3] = { fmul(fadd(bx.x, bx.y), .5f), fmul(fadd(by.x, by.y), .5f), fmul(fadd(bz.x, bz.y), .5f) };
                uint32_t i = 0;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                        const bool up = p[a] > c[a];
                        i = 2u * i + (up ? 1u : 0u);
                        lo[a] = up ? fmaxf(lo[a], c[a]) : lo[a];
                        hi[a] = up ? hi[a] : fminf(hi[a], c[a]);
                }
                x = 2u * x + (i >> 2);
                y = 2u * y + ((i >> 1) & 1u);
                z = 2u * z + (i & 1u);
                node = ((rec.y >> i) & 1u) ? rec.x + __popc(rec.y & ((1u << i) - 1u)) : kGiAbsent;
                float* e = sc + (size_t)(kGiPathWords * l) * stride;
                ++l;
                e[0] = lo[0];
                e[stride] = lo[1];
                e[2 * stride] = lo[2];
                e[3 * stride] = hi[0];
                e[4 * stride] = hi[1];
                e[5 * stride] = hi[2];
                e[6 * stride] = __uint_as_float(node);
                if (node == kGiAbsent)
                        break;
        }
        gp.x = x;
        gp.y = y;
        gp.z = z;
        gp.len = l;
        return (l == s) ? node : kGiAbsent;
}

// ---------------------------------------------------------------------------
// Step table.  The marching distances of cone_trace (voxel_octree.cc:252-281) do not depend on the cone:
// dist_0 = mindist, dist_(k+1) = dist_k + step * diam_k with diam_k = max(mindist, 2 * aperture * dist_k), and with
// them the sampled level int(log2f(maxdist / diam_k)) and the distance weight 1 / (1 + decay * dist_k) are functions
// of (min_voxel_size, root box) alone.  One thread evaluates the reference's expressions once per launch
// (gi_step_table_fill, the same float operations in the same order) and every cone of every ray reads the
// results: two IEEE divisions, the logarithm and the distance bookkeeping leave the per-sample loop.
// Layout: entry 0 = {count, truncated, dist to resume from, 0}; entry 1 + k = {dist_k, 1 / (1 + dist_k), level_k, 0}.
// A march longer than the table (kGiMaxSteps) continues with the original loop.
// ---------------------------------------------------------------------------
constexpr int kGiMaxSteps = 511;
constexpr float kGiAperture = 0.577350269f, kGiStep = .1f, kGiDecay = 1.f;

__device__ __forceinline__ float gi_maxdist(const float root[6])
{
        const float sx = fsub(root[3], root[0]), sy = fsub(root[4], root[1]), sz = fsub(root[5], root[2]);
        return __fsqrt_rn(dot3(sx, sy, sz, sx, sy, sz));
}

__device__ __forceinline__ void gi_step_table_fill(const float root[6], float min_voxel_size, float4* tab)
{
        const float mindist = fmul(1.414f, min_voxel_size);
        const float maxdist = gi_maxdist(root);
        float dist = mindist;
        int n = 0;
        bool truncated = false;
        while (dist < maxdist) {
                const float diam = std_max(mindist, fmul(fmul(kGiAperture, 2.f), dist));
                if (maxdist < diam)
                        break;
                if (n == kGiMaxSteps) {
                        truncated = true;
                        break;
                }
                tab[1 + n] = make_float4(dist, fdiv(1.f, fadd(1.f, fmul(kGiDecay, dist))),
                                         __int_as_float(gi_split_level(fdiv(maxdist, diam))), 0.f);
                ++n;
                dist = fadd(dist, fmul(kGiStep, diam));
        }
        tab[0] = make_float4(__int_as_float(n), __int_as_float(truncated ? 1 : 0), dist, 0.f);
}

// One sample of cone_trace (the loop body after the level is known) voxel_octree.cc:259-279.
__device__ __forceinline__ void gi_cone_sample(const TreeDev& tr, float* sc, int stride, GiPath& gp, const float o[3],
                                               const float d[3], const float coeff[6], int lobes, float dist,
                                               int split_level, float inv_w, float& opacity, float diffuse[3])
{
        const float p[3] = { fadd(o[0], fmul(d[0], dist)), fadd(o[1], fmul(d[1], dist)), fadd(o[2], fmul(d[2], dist)) };
        // descend `split_level` levels (or to a leaf): deeper than the tree, or into an absent child
        // (one of the reference's empty leaves, whose sample adds exact zeros) -> nothing to add
        uint32_t node = kGiAbsent;
        if (tr.num_nodes != 0 && split_level <= tr.L)
                node = (split_level == 0) ? 0u : gi_locate(tr, sc, stride, gp, p, split_level);
        if (node == kGiAbsent)
                return;
        const float4* g4 = reinterpret_cast<const float4*>(tr.gi + (size_t)kGiStride * node);
        float illum[3] = { 0.f, 0.f, 0.f };
        float coverage;
        if (lobes >= 0) {
                // Only the three lobes that face the cone: per axis one of (+a, -a) has the coefficient
                // clamp(<= 0) = +-0, and 0 * il = +-0 added to the running sum changes nothing -- the sum starts at
                // +0 and can never be -0 (x + y is -0 only for x = y = -0), the il are finite (TreeDev::gi_ok).  The
                // remaining terms are added in the reference's order: lobes = their indices, ascending, 3 bits each;
                // coeff[0..2] = their coefficients in that order.
                const float4 a = __ldg(g4 + (lobes & 7)), b = __ldg(g4 + ((lobes >> 3) & 7)), c = __ldg(g4 + (lobes >> 6));
                illum[0] = fadd(fadd(fadd(0.f, fmul(coeff[0], a.x)), fmul(coeff[1], b.x)), fmul(coeff[2], c.x));
                illum[1] = fadd(fadd(fadd(0.f, fmul(coeff[0], a.y)), fmul(coeff[1], b.y)), fmul(coeff[2], c.y));
                illum[2] = fadd(fadd(fadd(0.f, fmul(coeff[0], a.z)), fmul(coeff[1], b.z)), fmul(coeff[2], c.z));
                coverage = a.w;
        } else {
                const float4 q0 = __ldg(g4), q1 = __ldg(g4 + 1), q2 = __ldg(g4 + 2), q3 = __ldg(g4 + 3), q4 = __ldg(g4 + 4),
                             q5 = __ldg(g4 + 5);
                const float il[18] = { q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, q2.x, q2.y, q2.z,
                                       q3.x, q3.y, q3.z, q4.x, q4.y, q4.z, q5.x, q5.y, q5.z };
                coverage = q0.w;
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                                illum[k] = fadd(illum[k], fmul(coeff[i], il[3 * i + k]));
        }
        const float transparency = clampf(fsub(1.f, opacity), 0.f, 1.f);
        const float a = fmul(coverage, kGiStep);
        const float w = fmul(fmul(inv_w, transparency), coverage);
#pragma unroll
        for (int k = 0; k < 3; ++k)
                diffuse[k] = fadd(diffuse[k], fmul(w, illum[k]));
        opacity = fadd(opacity, fmul(transparency, a));
}

// cone_trace(root, cone, min_voxel_size) voxel_octree.cc:247-283.  steps: the launch's step table, or null.
__device__ __forceinline__ void gi_cone_one(const TreeDev& tr, const float root[6], float* sc, int stride, GiPath& gp,
                                            const float o[3], const float d[3], float min_voxel_size,
                                            const float4* __restrict__ steps, bool lobes3, float out[3])
{
        const float mindist = fmul(1.414f, min_voxel_size);
        const float maxdist = gi_maxdist(root);
        // compute_illum(-cone.d): the six lobe coefficients clamp(dot(illum_d[i], -d), 0, 1) do not depend
        // on the sample (illum_d[] = +x +y +z -x -y -z, voxel_octree.cc:19-20; jql::dot keeps the zero terms)
        float coeff[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
                const float sg = (i < 3) ? 1.f : -1.f;
                const float ax = (i % 3 == 0) ? sg : 0.f, ay = (i % 3 == 1) ? sg : 0.f, az = (i % 3 == 2) ? sg : 0.f;
                coeff[i] = clampf(dot3(ax, ay, az, -d[0], -d[1], -d[2]), 0.f, 1.f);
        }
        // the three lobes that face the cone, sorted by lobe index (= the order of the reference's sum), with their
        // coefficients moved to coeff[0..2]; or -1: all six (see gi_cone_sample)
        int lobes = -1;
        if (lobes3) {
                int k0 = (-d[0] > 0.f) ? 0 : 3, k1 = (-d[1] > 0.f) ? 1 : 4, k2 = (-d[2] > 0.f) ? 2 : 5;
                float c0 = (k0 == 0) ? coeff[0] : coeff[3], c1 = (k1 == 1) ? coeff[1] : coeff[4],
                      c2 = (k2 == 2) ? coeff[2] : coeff[5];
#define VRT_CSWAP(ka, ca, kb, cb)                 \
        if (ka > kb) {                            \
                const int tk = ka;                \
                ka = kb;                          \
                kb = tk;                          \
                const float tc = ca;              \
                ca = cb;                          \
                cb = tc;                          \
        }
                VRT_CSWAP(k0, c0, k1, c1)
                VRT_CSWAP(k1, c1, k2, c2)
                VRT_CSWAP(k0, c0, k1, c1)
#undef VRT_CSWAP
                lobes = k0 | (k1 << 3) | (k2 << 6);
                coeff[0] = c0;
                coeff[1] = c1;
                coeff[2] = c2;
        }
        float dist = mindist, opacity = 0.f;
        float diffuse[3] = { 0.f, 0.f, 0.f };
        bool rest = true;  // continue with the reference's own loop from `dist`
        if (steps != nullptr) {
                const float4 hdr = __ldg(steps);
                const int n = __float_as_int(hdr.x);
                int k = 0;
                // (the entry of step k + 1 is requested before step k is sampled: entry 1 + n exists, the table has
                // kGiMaxSteps + 2 entries)
                float4 st = __ldg(steps + 1);
                for (; k < n && opacity < 1.f; ++k) {
                        const float4 cur = st;
                        st = __ldg(steps + 2 + k);
                        gi_cone_sample(tr, sc, stride, gp, o, d, coeff, lobes, cur.x, __float_as_int(cur.z), cur.y, opacity, diffuse);
                }
                rest = (k == n) && (__float_as_int(hdr.y) != 0);
                dist = hdr.z;
        }
        while (rest && dist < maxdist && opacity < 1.f) {
                const float diam = std_max(mindist, fmul(fmul(kGiAperture, 2.f), dist));
                if (maxdist < diam)
                        break;
                gi_cone_sample(tr, sc, stride, gp, o, d, coeff, lobes, dist, gi_split_level(fdiv(maxdist, diam)),
                               fdiv(1.f, fadd(1.f, fmul(kGiDecay, dist))), opacity, diffuse);
                dist = fadd(dist, fmul(kGiStep, diam));
        }
        out[0] = diffuse[0];
        out[1] = diffuse[1];
        out[2] = diffuse[2];
}

// cone_trace(root, isect, min_voxel_size) voxel_octree.cc:285-303 with orthonormal_basis :236-245
// and the six HemiCones :227-234.
// sc: this thread's column of the path cache (kGiPathWords * L words, element stride `stride`).
__device__ __forceinline__ void gi_cone_trace_point(const TreeDev& tr, const float root[6], float* sc, int stride,
                                                    const float pos[3], const float n[3], float res,
                                                    const float4* __restrict__ steps, float out[3])
{
        GiPath gp;
        gi_path_reset(gp);
        const bool lobes3 = tr.gi_ok != nullptr && __ldg(tr.gi_ok) != 0u;
        const float hemi[6][4] = {
                { 0.000000f, 0.000000f, 1.0f, 0.25f },   { 0.000000f, 0.866025f, 0.5f, 0.15f },
                { 0.823639f, 0.267617f, 0.5f, 0.15f },   { 0.509037f, -0.700629f, 0.5f, 0.15f },
                { -0.509037f, -0.700629f, 0.5f, 0.15f }, { -0.823639f, 0.267617f, 0.5f, 0.15f },
        };
        const float s = (0.0f > n[2]) ? -1.0f : 1.0f;
        const float a0 = fdiv(-1.0f, fadd(s, n[2]));
        const float a1 = fmul(fmul(n[0], n[1]), a0);
        const float tv[3] = { fadd(1.0f, fmul(fmul(fmul(s, n[0]), n[0]), a0)), fmul(s, a1), fmul(-s, n[0]) };
        const float bv[3] = { a1, fadd(s, fmul(fmul(n[1], n[1]), a0)), -n[1] };
        float diffuse[3] = { 0.f, 0.f, 0.f };
#pragma unroll 1
        for (int i = 0; i < 6; ++i) {
                float d[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {  // dot(Mat3{t,b,n}, v): result += column_j * v[j], from 0
                        float r = fadd(0.f, fmul(tv[k], hemi[i][0]));
                        r = fadd(r, fmul(bv[k], hemi[i][1]));
                        r = fadd(r, fmul(n[k], hemi[i][2]));
                        d[k] = r;
                }
                normalize3(d[0], d[1], d[2]);
                float c[3];
                gi_cone_one(tr, root, sc, stride, gp, pos, d, res, steps, lobes3, c);
#pragma unroll
                for (int k = 0; k < 3; ++k)
                        diffuse[k] = fadd(diffuse[k], fmul(hemi[i][3], c[k]));
        }
        out[0] = diffuse[0];
        out[1] = diffuse[1];
        out[2] = diffuse[2];
}

}  // namespace vrt
