// oracle/ref_harness.cc -- TEST INFRASTRUCTURE, not product code.
//
// A thin extern "C" harness that is compiled TOGETHER WITH THE UNMODIFIED
// REFERENCE SOURCES (camera.cc voxel_octree.cc tribox2.cc raytri.cc
// tiny_obj_loader.cc, see oracle/build_ref.sh) into oracle/_ref/libvrt_ref.so.
// It contains no algorithm of its own: every result it returns is produced by
// the reference's functions
//   gi::ray_march_init   voxel_octree.cc:67-75
//   gi::ray_march        voxel_octree.cc:131-188
//   Camera::gen_rays1/4  camera.cc:77-112
//   render_mt            camera.h:41-68
//   triBoxOverlap        tribox2.cc:112-186
//   intersect_triangle3  raytri.cc:197-249
//   gi::cone_trace_init_filter / gi::cone_trace   voxel_octree.cc:190-303
//   VoxelOctree::compute_illum, Triangle::get_diffuse/get_albedo
//   (the two main.cc lambdas -- light-map splat main.cc:81-96 and trace()
//   main.cc:10-30 -- are not linkable, main.cc holds main(); the GI section at the
//   end of this file drives the same reference calls in the same order)
// and is only re-shaped into flat arrays so that Python tests (ctypes) and
// bench.py's `--impl reference` / `cpu_baseline` legs can read it.
//
// Only tests/, __graft_entry__.smoke() and bench.py's reference legs may load
// the library built from this file.

#include "voxel_octree.h"
#include "camera.h"
#include "stb_image.h"  // declarations only; the implementation is compiled into voxel_octree.cc:13-14

#include <chrono>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace {

struct LeafInfo {
        uint32_t x, y, z;  // cell coordinates at level max_depth-1
        uint32_t level;    // depth of the node below the root (root = 0)
};

struct RefScene {
        tinyobj::material_t mtl{};
        std::vector<tinyobj::material_t> mtls;  // ref_scene_create_mat: per-triangle materials
        std::vector<gi::Triangle> tris;
        std::vector<gi::VoxelBase*> ptrs;
        std::unique_ptr<gi::VoxelOctree> root;
        int max_depth = 0;
        double build_seconds = 0;
        // filled by index()
        std::unordered_map<const gi::VoxelOctree*, LeafInfo> leaf_of;
        std::unordered_map<const gi::VoxelBase*, uint32_t> tri_of;
        uint64_t n_nodes = 0, n_interior = 0, n_nonempty_leaves = 0, n_refs = 0;
        uint32_t max_leaf_refs = 0;
};

bool node_is_leaf(const gi::VoxelOctree* n)
{
        return n->children[0] == nullptr;
}

void index_walk(RefScene* s, const gi::VoxelOctree* n, uint32_t x, uint32_t y,
                uint32_t z, uint32_t level)
{
        s->n_nodes++;
        if (node_is_leaf(n)) {
                if (!n->voxels.empty()) {
                        s->n_nonempty_leaves++;
                        s->n_refs += n->voxels.size();
                        if (n->voxels.size() > s->max_leaf_refs)
                                s->max_leaf_refs = (uint32_t)n->voxels.size();
                        s->leaf_of[n] = LeafInfo{ x, y, z, level };
                }
                return;
        }
        s->n_interior++;
        for (int i = 0; i < 8; ++i) {
                // child i: x is bit 2, y is bit 1, z is bit 0 (voxel_octree.cc:33)
                index_walk(s, n->children[i].get(), 2 * x + ((i >> 2) & 1),
                           2 * y + ((i >> 1) & 1), 2 * z + (i & 1), level + 1);
        }
}

struct HitRec {
        const gi::VoxelOctree* leaf;
        const gi::VoxelBase* voxel;
        ISect isect;
        uint8_t hit;
};

void store_hit(const RefScene* s, const HitRec& h, uint8_t* hit, uint32_t* cell,
               uint32_t* tri, float* pos, float* nrm, uint64_t i)
{
        hit[i] = h.hit;
        if (!h.hit) {
                if (cell)
                        cell[3 * i] = cell[3 * i + 1] = cell[3 * i + 2] = 0xffffffffu;
                if (tri)
                        tri[i] = 0xffffffffu;
                if (pos)
                        pos[3 * i] = pos[3 * i + 1] = pos[3 * i + 2] = 0.f;
                if (nrm)
                        nrm[3 * i] = nrm[3 * i + 1] = nrm[3 * i + 2] = 0.f;
                return;
        }
        if (cell) {
                auto it = s->leaf_of.find(h.leaf);
                if (it == s->leaf_of.end()) {
                        cell[3 * i] = cell[3 * i + 1] = cell[3 * i + 2] = 0xfffffffeu;
                } else {
                        cell[3 * i] = it->second.x;
                        cell[3 * i + 1] = it->second.y;
                        cell[3 * i + 2] = it->second.z;
                }
        }
        if (tri)
                tri[i] = s->tri_of.at(h.voxel);
        if (pos) {
                pos[3 * i] = h.isect.hit.x;
                pos[3 * i + 1] = h.isect.hit.y;
                pos[3 * i + 2] = h.isect.hit.z;
        }
        if (nrm) {
                nrm[3 * i] = h.isect.normal.x;
                nrm[3 * i + 1] = h.isect.normal.y;
                nrm[3 * i + 2] = h.isect.normal.z;
        }
}

Camera make_camera(const float* cam /* fov, eye3, spot3, up3 */)
{
        return Camera{ cam[0],
                       Vec3{ cam[1], cam[2], cam[3] },
                       Vec3{ cam[4], cam[5], cam[6] },
                       Vec3{ cam[7], cam[8], cam[9] } };
}

}  // namespace

extern "C" {

// tri_xyz: [T][3][3] float; tri_nrm: [T][3][3] float or NULL (then the
// geometric normal of each triangle is fed to the reference's ctor).
void* ref_scene_create(const float* tri_xyz, const float* tri_nrm, uint32_t T)
{
        auto* s = new RefScene;
        s->tris.reserve(T);
        for (uint32_t t = 0; t < T; ++t) {
                const float* p = tri_xyz + 9 * (size_t)t;
                Vec3 p0{ p[0], p[1], p[2] }, p1{ p[3], p[4], p[5] },
                        p2{ p[6], p[7], p[8] };
                Vec3 n0, n1, n2;
                if (tri_nrm) {
                        const float* n = tri_nrm + 9 * (size_t)t;
                        n0 = Vec3{ n[0], n[1], n[2] };
                        n1 = Vec3{ n[3], n[4], n[5] };
                        n2 = Vec3{ n[6], n[7], n[8] };
                } else {
                        n0 = n1 = n2 = jql::cross(p1 - p0, p2 - p0);
                }
                s->tris.emplace_back(p0, p1, p2, n0, n1, n2, Vec2{}, Vec2{}, Vec2{},
                                     &s->mtl);
        }
        s->ptrs.reserve(T);
        for (auto& tri : s->tris)
                s->ptrs.push_back(&tri);
        for (uint32_t t = 0; t < T; ++t)
                s->tri_of[s->ptrs[t]] = t;
        return s;
}

void ref_scene_free(void* h)
{
        delete static_cast<RefScene*>(h);
}

// gi::ray_march_init, timed.  Returns seconds of the reference call alone.
double ref_scene_build(void* h, int max_depth)
{
        auto* s = static_cast<RefScene*>(h);
        s->root = std::make_unique<gi::VoxelOctree>();
        s->max_depth = max_depth;
        auto t0 = std::chrono::steady_clock::now();
        gi::ray_march_init(s->root.get(), s->ptrs, max_depth);
        auto t1 = std::chrono::steady_clock::now();
        s->build_seconds = std::chrono::duration<double>(t1 - t0).count();
        s->leaf_of.clear();
        s->n_nodes = s->n_interior = s->n_nonempty_leaves = s->n_refs = 0;
        s->max_leaf_refs = 0;
        index_walk(s, s->root.get(), 0, 0, 0, 0);
        return s->build_seconds;
}

// out[0..5] = counts: nodes, interior, non-empty leaves, tri refs, max refs per leaf, max_depth
void ref_scene_stats(void* h, uint64_t* out)
{
        auto* s = static_cast<RefScene*>(h);
        out[0] = s->n_nodes;
        out[1] = s->n_interior;
        out[2] = s->n_nonempty_leaves;
        out[3] = s->n_refs;
        out[4] = s->max_leaf_refs;
        out[5] = (uint64_t)s->max_depth;
}

void ref_scene_root_aabb(void* h, float* out6)
{
        auto* s = static_cast<RefScene*>(h);
        out6[0] = s->root->aabb.min.x;
        out6[1] = s->root->aabb.min.y;
        out6[2] = s->root->aabb.min.z;
        out6[3] = s->root->aabb.max.x;
        out6[4] = s->root->aabb.max.y;
        out6[5] = s->root->aabb.max.z;
}

namespace {
void dump_walk(const RefScene* s, const gi::VoxelOctree* n, uint32_t x, uint32_t y,
               uint32_t z, uint32_t level, uint32_t* cells, uint32_t* levels,
               uint32_t* counts, uint32_t* refs, float* boxes, uint64_t* nl,
               uint64_t* nr)
{
        if (node_is_leaf(n)) {
                if (n->voxels.empty())
                        return;
                uint64_t l = (*nl)++;
                cells[3 * l] = x;
                cells[3 * l + 1] = y;
                cells[3 * l + 2] = z;
                if (levels)
                        levels[l] = level;
                counts[l] = (uint32_t)n->voxels.size();
                if (boxes) {
                        boxes[6 * l + 0] = n->aabb.min.x;
                        boxes[6 * l + 1] = n->aabb.min.y;
                        boxes[6 * l + 2] = n->aabb.min.z;
                        boxes[6 * l + 3] = n->aabb.max.x;
                        boxes[6 * l + 4] = n->aabb.max.y;
                        boxes[6 * l + 5] = n->aabb.max.z;
                }
                for (auto* v : n->voxels)
                        refs[(*nr)++] = s->tri_of.at(v);
                return;
        }
        for (int i = 0; i < 8; ++i)
                dump_walk(s, n->children[i].get(), 2 * x + ((i >> 2) & 1),
                          2 * y + ((i >> 1) & 1), 2 * z + (i & 1), level + 1, cells,
                          levels, counts, refs, boxes, nl, nr);
}
}  // namespace

// Depth-first (child index order == Morton order) dump of every non-empty leaf:
// cells[L][3], levels[L] (may be NULL), counts[L], refs[sum counts] in stored
// order, boxes[L][6] (may be NULL) = the leaf AABBs exactly as the reference's
// split() recurrence produced them.
void ref_scene_dump_leaves(void* h, uint32_t* cells, uint32_t* levels,
                           uint32_t* counts, uint32_t* refs, float* boxes)
{
        auto* s = static_cast<RefScene*>(h);
        uint64_t nl = 0, nr = 0;
        dump_walk(s, s->root.get(), 0, 0, 0, 0, cells, levels, counts, refs, boxes,
                  &nl, &nr);
}

// gi::ray_march over an explicit ray batch.  rays: [R][8] = o3,d3,tmin,tmax,
// copied VERBATIM into jql::Ray (no re-normalisation: the Ray ctor is bypassed
// so that both sides see identical bits).
void ref_trace_rays(void* h, const float* rays, uint64_t R, int nthreads,
                    uint8_t* hit, uint32_t* cell, uint32_t* tri, float* pos,
                    float* nrm)
{
        auto* s = static_cast<RefScene*>(h);
        if (nthreads < 1)
                nthreads = 1;
        auto work = [&](uint64_t lo, uint64_t hi) {
                for (uint64_t i = lo; i < hi; ++i) {
                        const float* r = rays + 8 * i;
                        jql::Ray ray;
                        ray.o = Vec3{ r[0], r[1], r[2] };
                        ray.d = Vec3{ r[3], r[4], r[5] };
                        ray.tmin = r[6];
                        ray.tmax = r[7];
                        HitRec rec{};
                        gi::VoxelOctree* leaf = nullptr;
                        gi::VoxelBase* vox = nullptr;
                        rec.hit = gi::ray_march(s->root.get(), ray, &leaf, &vox,
                                                &rec.isect) ? 1 : 0;
                        rec.leaf = leaf;
                        rec.voxel = vox;
                        store_hit(s, rec, hit, cell, tri, pos, nrm, i);
                }
        };
        if (nthreads == 1) {
                work(0, R);
                return;
        }
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t)
                th.emplace_back(work, R * t / nthreads, R * (t + 1) / nthreads);
        for (auto& t : th)
                t.join();
}

// Camera::gen_rays1 / gen_rays4 for the pixel rectangle [x0,x1) x [y0,y1).
// rays_out: [(y1-y0)][(x1-x0)][spp][8], row-major by pixel then sample.
void ref_gen_rays(const float* cam10, float film_w, float film_h, int nx, int ny,
                  int spp, int x0, int y0, int x1, int y1, float* rays_out)
{
        Camera cam = make_camera(cam10);
        Film film(film_w, film_h, nx, ny);
        const Film& f = film;
        size_t k = 0;
        for (int py = y0; py < y1; ++py)
                for (int px = x0; px < x1; ++px) {
                        auto rays = (spp == 4) ? cam.gen_rays4(f, px, py)
                                               : cam.gen_rays1(f, px, py);
                        for (auto& r : rays) {
                                float* o = rays_out + 8 * k++;
                                o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z;
                                o[3] = r.d.x; o[4] = r.d.y; o[5] = r.d.z;
                                o[6] = r.tmin; o[7] = r.tmax;
                        }
                }
}

// The reference's own render loop: render_mt(&film, lambda) with the lambda
// = gen_rays{1,4} + gi::ray_march that only records the hit (the CPU baseline
// of SURVEY.md 8(d)).  Returns wall seconds of the render_mt call.  Outputs
// (any may be NULL) are [ny][nx][spp] on a harness-owned y*nx+x layout.
// NOTE render_mt only renders (nx/8*8) x (ny/8*8) pixels (camera.h:46).
double ref_render_mt(void* h, const float* cam10, float film_w, float film_h,
                     int nx, int ny, int spp, uint8_t* hit, uint32_t* cell,
                     uint32_t* tri, float* pos, float* nrm, uint64_t* rays_traced)
{
        auto* s = static_cast<RefScene*>(h);
        Camera cam = make_camera(cam10);
        Film film_obj(film_w, film_h, nx, ny);
        Film* film = &film_obj;
        std::vector<HitRec> recs((size_t)nx * ny * spp);
        gi::VoxelOctree* root = s->root.get();
        HitRec* recp = recs.data();
        auto t0 = std::chrono::steady_clock::now();
        render_mt(film, [&cam, root, recp, spp](Film* f, int px, int py) {
                auto rays = (spp == 4) ? cam.gen_rays4(*f, px, py)
                                       : cam.gen_rays1(*f, px, py);
                size_t base = ((size_t)py * f->nx + px) * spp;
                for (size_t k = 0; k < rays.size(); ++k) {
                        HitRec& rec = recp[base + k];
                        gi::VoxelOctree* leaf = nullptr;
                        gi::VoxelBase* vox = nullptr;
                        rec.hit = gi::ray_march(root, rays[k], &leaf, &vox,
                                                &rec.isect) ? 1 : 0;
                        rec.leaf = leaf;
                        rec.voxel = vox;
                }
        });
        auto t1 = std::chrono::steady_clock::now();
        if (rays_traced)
                *rays_traced = (uint64_t)(nx / 8 * 8) * (ny / 8 * 8) * spp;
        if (hit)
                for (size_t i = 0; i < recs.size(); ++i)
                        store_hit(s, recs[i], hit, cell, tri, pos, nrm, i);
        return std::chrono::duration<double>(t1 - t0).count();
}

int ref_hardware_concurrency(void)
{
        return (int)std::thread::hardware_concurrency();
}

// Direct predicate access (known-answer tests).
int ref_tribox(const float* center, const float* half, const float* tri9)
{
        float c[3] = { center[0], center[1], center[2] };
        float hs[3] = { half[0], half[1], half[2] };
        float tv[3][3];
        std::memcpy(tv, tri9, sizeof tv);
        return triBoxOverlap(c, hs, tv);
}

void ref_tribox_batch(const float* centers, const float* halves, const float* tris,
                      uint64_t n, uint8_t* out)
{
        for (uint64_t i = 0; i < n; ++i)
                out[i] = (uint8_t)ref_tribox(centers + 3 * i, halves + 3 * i,
                                             tris + 9 * i);
}

// Triangle::is_overlap (voxel_octree.cc:486-492) on an explicit AABB.
void ref_tri_overlap_aabb_batch(const float* aabbs6, const float* tris, uint64_t n,
                                uint8_t* out)
{
        tinyobj::material_t mtl{};
        for (uint64_t i = 0; i < n; ++i) {
                const float* p = tris + 9 * i;
                const float* b = aabbs6 + 6 * i;
                Vec3 p0{ p[0], p[1], p[2] }, p1{ p[3], p[4], p[5] },
                        p2{ p[6], p[7], p[8] };
                Vec3 nn{ 0, 1, 0 };
                gi::Triangle tri(p0, p1, p2, nn, nn, nn, Vec2{}, Vec2{}, Vec2{}, &mtl);
                AABB3D box{ Vec3{ b[0], b[1], b[2] }, Vec3{ b[3], b[4], b[5] } };
                out[i] = tri.is_overlap(box) ? 1 : 0;
        }
}

// intersect_triangle3 on doubles: in [n][15] = o3,d3,v0,v1,v2 ; out [n][3]=t,u,v
void ref_raytri_batch(const double* in, uint64_t n, uint8_t* res, double* tuv)
{
        for (uint64_t i = 0; i < n; ++i) {
                double a[15];
                std::memcpy(a, in + 15 * i, sizeof a);
                double t = 0, u = 0, v = 0;
                res[i] = (uint8_t)intersect_triangle3(a, a + 3, a + 6, a + 9, a + 12,
                                                      &t, &u, &v);
                tuv[3 * i] = t;
                tuv[3 * i + 1] = u;
                tuv[3 * i + 2] = v;
        }
}

// AABB<Vec3>::isect(ray, nullptr) (graphics_math.h:1312-1332), batch.
void ref_aabb_isect_batch(const float* aabbs6, const float* rays8, uint64_t n,
                          uint8_t* out)
{
        for (uint64_t i = 0; i < n; ++i) {
                const float* b = aabbs6 + 6 * i;
                const float* r = rays8 + 8 * i;
                AABB3D box{ Vec3{ b[0], b[1], b[2] }, Vec3{ b[3], b[4], b[5] } };
                jql::Ray ray;
                ray.o = Vec3{ r[0], r[1], r[2] };
                ray.d = Vec3{ r[3], r[4], r[5] };
                ray.tmin = r[6];
                ray.tmax = r[7];
                out[i] = box.isect(ray, nullptr) ? 1 : 0;
        }
}

// Camera ctor (camera.cc:65-75): returns the 16 floats of C_ (column-major,
// C_[col][row]) by probing the camera with unit vectors through gen_rays-free
// public API is impossible (C_ is private), so we rebuild it with the same jql
// calls the ctor uses.
void ref_camera_matrix(const float* cam10, float* out16)
{
        Vec3 eye{ cam10[1], cam10[2], cam10[3] };
        Vec3 spot{ cam10[4], cam10[5], cam10[6] };
        Vec3 up{ cam10[7], cam10[8], cam10[9] };
        const Vec3 forward_ = normalize(spot - eye);
        const Vec3 s = normalize(cross(forward_, up));
        const Vec3 up_ = normalize(cross(s, forward_));
        Mat4 C = jql::affine_transform(Mat3{ s, up_, -forward_ }, eye);
        std::memcpy(out16, jql::begin(C), 16 * sizeof(float));
}

}  // extern "C"

// ---------------------------------------------------------------------------
// GI rows of SURVEY.md 8(f): light-map splat, bottom-up filter, cone trace and
// the final trace() pixel.  All arithmetic is the reference's (ray_march,
// get_diffuse, compute_illum, cone_trace_init_filter, cone_trace); the harness
// only sequences the calls like main.cc does.
// ---------------------------------------------------------------------------
namespace {

void gi_reset_walk(gi::VoxelOctree* n)
{
        n->coverage = 0.f;
        for (int f = 0; f < 6; ++f)
                n->illum[f] = Vec3{ 0, 0, 0 };
        if (!node_is_leaf(n))
                for (int i = 0; i < 8; ++i)
                        gi_reset_walk(n->children[i].get());
}

struct GiDump {
        int level;
        uint32_t* cells;
        float* cov;
        float* illum;
        uint64_t cap, n;
};

void gi_dump_walk(const gi::VoxelOctree* n, uint32_t x, uint32_t y, uint32_t z, int level, GiDump* d)
{
        if (level == d->level) {
                if (n->coverage > 0.f) {
                        if (d->n < d->cap) {
                                if (d->cells) {
                                        d->cells[3 * d->n] = x;
                                        d->cells[3 * d->n + 1] = y;
                                        d->cells[3 * d->n + 2] = z;
                                }
                                if (d->cov)
                                        d->cov[d->n] = n->coverage;
                                if (d->illum)
                                        for (int f = 0; f < 6; ++f) {
                                                d->illum[18 * d->n + 3 * f] = n->illum[f].x;
                                                d->illum[18 * d->n + 3 * f + 1] = n->illum[f].y;
                                                d->illum[18 * d->n + 3 * f + 2] = n->illum[f].z;
                                        }
                        }
                        d->n++;
                }
                return;
        }
        if (node_is_leaf(n))
                return;
        for (int i = 0; i < 8; ++i)
                gi_dump_walk(n->children[i].get(), 2 * x + ((i >> 2) & 1), 2 * y + ((i >> 1) & 1),
                             2 * z + (i & 1), level + 1, d);
}

// main.cc:10-30 trace(), with the reference's calls in the reference's order.
Vec3 gi_trace_pixel(gi::VoxelOctree& root, const jql::Ray& ray, float res)
{
        gi::VoxelOctree* leaf_ptr{};
        gi::VoxelBase* voxel_ptr{};
        ISect isect{};
        if (!gi::ray_march(&root, ray, &leaf_ptr, &voxel_ptr, &isect, true)) {
                float t = 0.5 * (ray.d.y + 1.0);
                return jql::lerp(Vec3{ 1.0f, 1.0f, 1.0f }, Vec3{ 0.6f, 0.8f, 1.0f }, t);
        }
        auto indirect_light = gi::cone_trace(root, isect, res);
        Vec3 direct_light = leaf_ptr->compute_illum(-ray.d);
        if (voxel_ptr->is_visible())
                return voxel_ptr->get_albedo(isect) * (indirect_light + direct_light);
        return voxel_ptr->get_albedo(isect);
}

}  // namespace

extern "C" {

// tinyobj::material_t::diffuse of the scene's (single, untextured) material
void ref_scene_set_diffuse(void* h, const float* rgb)
{
        auto* s = static_cast<RefScene*>(h);
        for (int k = 0; k < 3; ++k)
                s->mtl.diffuse[k] = rgb[k];
}

void ref_gi_reset(void* h)
{
        gi_reset_walk(static_cast<RefScene*>(h)->root.get());
}

// The light-map pass of main.cc:81-96, executed SEQUENTIALLY in pixel order (py outer,
// px inner, samples in gen_rays order).  The reference runs the same lambda on pool
// threads without synchronising the += on the shared leaf, so its own result depends on
// the interleaving; the sequential order is the deterministic member of that family.
void ref_gi_splat(void* h, const float* cam10, float film_w, float film_h, int nx, int ny, int spp)
{
        auto* s = static_cast<RefScene*>(h);
        Camera scam = make_camera(cam10);
        Film film(film_w, film_h, nx, ny);
        gi::VoxelOctree& root = *s->root;
        for (int py = 0; py < ny; ++py)
                for (int px = 0; px < nx; ++px) {
                        auto rays = (spp == 4) ? scam.gen_rays4(film, px, py) : scam.gen_rays1(film, px, py);
                        for (const auto& ray : rays) {
                                gi::VoxelOctree* leaf_ptr{};
                                jql::ISect isect{};
                                gi::VoxelBase* voxel_ptr{};
                                if (!gi::ray_march(&root, ray, &leaf_ptr, &voxel_ptr, &isect))
                                        continue;
                                auto illum = voxel_ptr->get_diffuse(isect, ray, Vec3{ 1, 1, 1 });
                                for (int i = 0; i < 6; ++i) {
                                        float coeff = jql::dot(leaf_ptr->illum_d[i], isect.normal);
                                        coeff = jql::clamp(coeff, 0.f, 1.f);
                                        leaf_ptr->illum[i] += coeff * illum;
                                }
                        }
                }
}

void ref_gi_filter(void* h)
{
        gi::cone_trace_init_filter(static_cast<RefScene*>(h)->root.get());
}

// Nodes of tree level `level` (root = 0) with coverage > 0, in Morton order (x bit 2, y bit 1,
// z bit 0 per level).  Returns their number; fills up to `cap` entries of the non-NULL arrays.
uint64_t ref_gi_dump_level(void* h, int level, uint32_t* cells, float* cov, float* illum18, uint64_t cap)
{
        GiDump d{ level, cells, cov, illum18, cap, 0 };
        gi_dump_walk(static_cast<RefScene*>(h)->root.get(), 0, 0, 0, 0, &d);
        return d.n;
}

// gi::cone_trace(root, ISect{hit, normal}, res) for n surface points.
void ref_gi_cone_trace(void* h, const float* pos, const float* nrm, uint64_t n, float res, float* out3)
{
        auto* s = static_cast<RefScene*>(h);
        for (uint64_t i = 0; i < n; ++i) {
                ISect is{};
                is.hit = Vec3{ pos[3 * i], pos[3 * i + 1], pos[3 * i + 2] };
                is.normal = Vec3{ nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2] };
                Vec3 c = gi::cone_trace(*s->root, is, res);
                out3[3 * i] = c.x;
                out3[3 * i + 1] = c.y;
                out3[3 * i + 2] = c.z;
        }
}

// The final image loop of main.cc:117-123 on a harness-owned [ny][nx][3] film: every sample's
// trace() colour times 1/spp, added in sample order.  Rows are split over `nthreads` threads
// (the tree is read-only here).  Returns wall seconds.
double ref_gi_render(void* h, const float* cam10, float film_w, float film_h, int nx, int ny, int spp,
                     float res, float* film3, int nthreads)
{
        auto* s = static_cast<RefScene*>(h);
        Camera cam = make_camera(cam10);
        Film film(film_w, film_h, nx, ny);
        gi::VoxelOctree& root = *s->root;
        const float w = (spp == 4) ? .25f : 1.f;
        auto work = [&](int y0, int y1) {
                for (int py = y0; py < y1; ++py)
                        for (int px = 0; px < nx; ++px) {
                                Vec3 acc{ 0, 0, 0 };
                                auto rays = (spp == 4) ? cam.gen_rays4(film, px, py) : cam.gen_rays1(film, px, py);
                                for (const auto& ray : rays) {
                                        auto c = gi_trace_pixel(root, ray, res);
                                        acc += c * w;
                                }
                                float* o = film3 + 3 * ((size_t)py * nx + px);
                                o[0] = acc.x;
                                o[1] = acc.y;
                                o[2] = acc.z;
                        }
        };
        auto t0 = std::chrono::steady_clock::now();
        if (nthreads <= 1) {
                work(0, ny);
        } else {
                std::vector<std::thread> th;
                for (int t = 0; t < nthreads; ++t)
                        th.emplace_back(work, (int)((long long)ny * t / nthreads), (int)((long long)ny * (t + 1) / nthreads));
                for (auto& t : th)
                        t.join();
        }
        auto t1 = std::chrono::steady_clock::now();
        return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Materials and textures: Triangle::get_albedo voxel_octree.cc:471-484,
// VoxelBase::texel_fetch + unit_cycle + load_image (stb_image) voxel_octree.cc:373-422.
// ---------------------------------------------------------------------------
extern "C" {

// Like ref_scene_create, with per-vertex texture coordinates tri_uv[T][3][2] and a material id per
// triangle; material m has diffuse kd[m] and, when texpath[m] is a non-empty string, that image file
// as its diffuse texture (loaded by the reference's own load_image on first use).
void* ref_scene_create_mat(const float* tri_xyz, const float* tri_nrm, const float* tri_uv, const uint32_t* tri_mtl,
                           uint32_t T, uint32_t M, const float* kd, const char* const* texpath)
{
        auto* s = new RefScene;
        s->mtls.resize(M);
        for (uint32_t m = 0; m < M; ++m) {
                for (int k = 0; k < 3; ++k)
                        s->mtls[m].diffuse[k] = kd[3 * m + k];
                s->mtls[m].diffuse_texname = texpath && texpath[m] ? texpath[m] : "";
        }
        s->tris.reserve(T);
        for (uint32_t t = 0; t < T; ++t) {
                const float* p = tri_xyz + 9 * (size_t)t;
                const float* n = tri_nrm + 9 * (size_t)t;
                const float* u = tri_uv + 6 * (size_t)t;
                s->tris.emplace_back(Vec3{ p[0], p[1], p[2] }, Vec3{ p[3], p[4], p[5] }, Vec3{ p[6], p[7], p[8] },
                                     Vec3{ n[0], n[1], n[2] }, Vec3{ n[3], n[4], n[5] }, Vec3{ n[6], n[7], n[8] },
                                     Vec2{ u[0], u[1] }, Vec2{ u[2], u[3] }, Vec2{ u[4], u[5] }, &s->mtls[tri_mtl[t]]);
        }
        s->ptrs.reserve(T);
        for (auto& tri : s->tris)
                s->ptrs.push_back(&tri);
        for (uint32_t t = 0; t < T; ++t)
                s->tri_of[s->ptrs[t]] = t;
        return s;
}

// Triangle::get_albedo(ISect{hit = pos}) of triangle tri[i]
void ref_albedo(void* h, const uint32_t* tri, const float* pos, uint64_t n, float* out3)
{
        auto* s = static_cast<RefScene*>(h);
        for (uint64_t i = 0; i < n; ++i) {
                ISect is{};
                is.hit = Vec3{ pos[3 * i], pos[3 * i + 1], pos[3 * i + 2] };
                Vec3 a = s->tris[tri[i]].get_albedo(is);
                out3[3 * i] = a.x;
                out3[3 * i + 1] = a.y;
                out3[3 * i + 2] = a.z;
        }
}

// The bytes the reference sees for an image file: stbi_load(path, &w, &h, &channels, 0) as in load_image
// (voxel_octree.cc:373-388).  Returns w*h*channels (0 on failure); copies at most cap bytes.
uint64_t ref_load_image(const char* path, int* w, int* hgt, int* channels, uint8_t* out, uint64_t cap)
{
        unsigned char* data = stbi_load(path, w, hgt, channels, 0);
        if (!data)
                return 0;
        const uint64_t size = (uint64_t)(*w) * (*hgt) * (*channels);
        if (out)
                std::memcpy(out, data, size < cap ? size : cap);
        stbi_image_free(data);
        return size;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Film export (main.cc:125-126, camera.cc:27-63): what the reference does with its float film.
// ---------------------------------------------------------------------------
#define STB_IMAGE_WRITE_IMPLEMENTATION
#define STB_IMAGE_WRITE_STATIC
#include "stb_image_write.h"

extern "C" {

// Film::to_byte_array() of a film whose pixels were set() to rgb[ny][nx][3]
void ref_film_to_bytes(const float* rgb, int nx, int ny, uint8_t* out)
{
        Film film(1.f, 1.f, nx, ny);
        for (int y = 0; y < ny; ++y)
                for (int x = 0; x < nx; ++x) {
                        const float* p = rgb + 3 * ((size_t)y * nx + x);
                        film.set(x, y, Vec3{ p[0], p[1], p[2] });
                }
        const auto d = film.to_byte_array();
        std::memcpy(out, d.data(), d.size());
}

// the bytes of the file stbi_write_hdr(name, nx, ny, 3, film.to_float_array().data()) writes; returns the size,
// copies at most cap bytes
uint64_t ref_write_hdr(const float* rgb, int nx, int ny, uint8_t* out, uint64_t cap)
{
        Film film(1.f, 1.f, nx, ny);
        for (int y = 0; y < ny; ++y)
                for (int x = 0; x < nx; ++x) {
                        const float* p = rgb + 3 * ((size_t)y * nx + x);
                        film.set(x, y, Vec3{ p[0], p[1], p[2] });
                }
        const auto d = film.to_float_array();
        std::vector<uint8_t> bytes;
        stbi_write_hdr_to_func(
                [](void* ctx, void* data, int size) {
                        auto* v = static_cast<std::vector<uint8_t>*>(ctx);
                        v->insert(v->end(), static_cast<uint8_t*>(data), static_cast<uint8_t*>(data) + size);
                },
                &bytes, nx, ny, 3, d.data());
        if (out)
                std::memcpy(out, bytes.data(), bytes.size() < cap ? bytes.size() : cap);
        return bytes.size();
}

}  // extern "C"
