// chunkycu.cu - kernels and the C ABI (include/chunkycu.h) of libchunkycu.so.  sm_100a only.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <chrono>

#include "ccu_host.h"
#include "ccu_march.cuh"
#include "ccu_queue.cuh"
#include "ccu_tonemap.cuh"
#include "ccu_layout.h"

using namespace ccu;

// ======================================================================================================
// kernels
// ======================================================================================================

// Sun_new (sky.h:19-40) once per scene, on the device so it uses the device's arithmetic.
__global__ void k_sun_setup(const int *sun_words, float *out /* su sv sw radius_cos : 10 floats */) {
    float phi = i2f(sun_words[4]), theta = i2f(sun_words[5]);
    float r = fabsf(dm_cos(phi));
    float3 sw = f3(dm_cos(theta) * r, dm_sin(phi), dm_sin(theta) * r);
    float3 su = (fabsf(sw.x) > 0.1f) ? f3(0, 1, 0) : f3(1, 0, 0);
    float3 sv = normalize3(cross3(sw, su));
    su = cross3(sv, sw);
    out[0] = su.x; out[1] = su.y; out[2] = su.z;
    out[3] = sv.x; out[4] = sv.y; out[5] = sv.z;
    out[6] = sw.x; out[7] = sw.y; out[8] = sw.z;
    out[9] = dm_cos(0.03f);
}

// Thread-per-pixel path tracer: all passes of the batch in one launch, the running mean of
// rayTracer.cl:109-112 carried in registers between passes (identical arithmetic, no memory round trip).
// MODE: 0 = the reference's own node array, 1 = commit-time layouts, 2 = commit-time layouts of a deep world.
template <int MODE>
__global__ void __launch_bounds__(128) k_render_mega(const __grid_constant__ DScene s, const int *__restrict__ seeds, int n_passes,
                                                     int start_spp, float *res, const float *res_prev, int n_pixels) {
    stage_tables(s, nullptr);
    for (int gid = blockIdx.x * blockDim.x + threadIdx.x; gid < n_pixels; gid += gridDim.x * blockDim.x) {
        float *px = res + (size_t)gid * 3;
        const float *pp = res_prev + (size_t)gid * 3;    // what the reference's single buffer holds when the window starts
        float3 buf = f3(pp[0], pp[1], pp[2]);
        for (int pass = 0; pass < n_passes; pass++) {
            float3 col = sample_pixel<MODE>(s, gid, __ldg(seeds + pass));
            int spp = start_spp + pass;
            float fs = (float)spp, fs1 = (float)(spp + 1);
            buf.x = (buf.x * fs + col.x) / fs1;
            buf.y = (buf.y * fs + col.y) / fs1;
            buf.z = (buf.z * fs + col.z) / fs1;
        }
        px[0] = buf.x; px[1] = buf.y; px[2] = buf.z;
    }
}

__device__ __forceinline__ int face_of(float3 n) {
    if (n.x == -1 && n.y == 0 && n.z == 0) return 0;
    if (n.x == 1 && n.y == 0 && n.z == 0) return 1;
    if (n.x == 0 && n.y == -1 && n.z == 0) return 2;
    if (n.x == 0 && n.y == 1 && n.z == 0) return 3;
    if (n.x == 0 && n.y == 0 && n.z == -1) return 4;
    if (n.x == 0 && n.y == 0 && n.z == 1) return 5;
    return 6;
}

// First-hit pass (BASELINE config 2): closestIntersect (kernel.h:14-24) for the camera ray of every pixel.
// Pixels are taken in 32x32 screen tiles, 8x4 patches per warp (tile_order_pixel<true>: the lanes of a warp walk neighbouring
// rays, which keeps their step counts and brick loads together).  With the commit-time layouts (MODE 1 / 2) the warp marches on
// the air layout until every lane sits at a non-air leaf or has left the octree, THEN runs the block tests of all those lanes
// together (the block test is ~3x the instructions of a march step; run per lane as each ray arrives it executed at 6 of 32
// lanes), and repeats for the lanes whose block test missed.  The treeData index of the hit leaf (reference numbering,
// ClSceneLoader.java:56-59) is found by one root descent for the hit voxel only.  HAS_BVH = false compiles the entity BVHs out.
#ifndef CCU_FH_WARP_SKIP
#define CCU_FH_WARP_SKIP 0
#endif
// threads per block of the first-hit kernel and resident blocks per SM (the warps of a block do not depend on each other, so
// the block size only decides how many warps wait for the slowest one before their registers are handed on)
#ifndef CCU_FH_THREADS
#define CCU_FH_THREADS 128
#endif
#ifndef CCU_FH_MIN_BLOCKS
#define CCU_FH_MIN_BLOCKS (1024 / CCU_FH_THREADS)
#endif
struct FirstHitOut { int *block, *face, *node, *kind; float *t, *normal, *color; };
__device__ __forceinline__ void first_hit_store(const FirstHitOut &out, int gid, bool hit, int material, float3 n, int node, int kind, float t, float4 color) {
    if (out.block) out.block[gid] = hit ? material : 0;
    if (out.face) out.face[gid] = hit ? face_of(n) : 6;
    if (out.node) out.node[gid] = hit ? node : -1;
    if (out.kind) out.kind[gid] = hit ? kind : 0;
    if (out.t) out.t[gid] = hit ? t : inff_();
    if (out.normal) {
        out.normal[gid * 3 + 0] = hit ? n.x : 0.0f;
        out.normal[gid * 3 + 1] = hit ? n.y : 0.0f;
        out.normal[gid * 3 + 2] = hit ? n.z : 0.0f;
    }
    if (out.color) reinterpret_cast<float4 *>(out.color)[gid] = hit ? color : make_float4(0, 0, 0, 0);
}

// The warp-synchronous march loop of the thread-per-ray kernels: every lane steps its ray until no lane of the warp is marching
// any more.  Kept out of line so that the loop is register-allocated by itself - inlined into the 64-register kernel, next to the
// block test, the ray was spilled and re-loaded inside the loop.  The ray travels through local memory once per call.
// st: 0 marching, anything else: not marching (returned unchanged).
template <bool DEEP>
static __device__ __noinline__ int first_hit_march(const DScene &s, LeanRay *ray, int st) {
    LeanRay r = *ray;
    while (__any_sync(0xffffffffu, st == 0)) {
#if CCU_FLAT_MARCH
        if (!DEEP) st = lean_step_flat<false, CCU_FH_WARP_SKIP != 0>(s, s.air_top, r, st == 0, st);
        else
#endif
        if (st == 0) st = lean_probe<DEEP, false>(s, s.air_top, r);
    }
    ray->t = r.t;
    ray->steps = r.steps;
    return st;
}

template <int MODE, bool HAS_BVH>
__global__ void __launch_bounds__(CCU_FH_THREADS, CCU_FH_MIN_BLOCKS) k_first_hit(const __grid_constant__ DScene s, int seed, int n_pixels, const __grid_constant__ FirstHitOut out) {
    stage_tables_per_warp(s);
    const unsigned full = 0xffffffffu;
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = k < (unsigned)n_pixels;
    unsigned px = 0, py = 0;
    if (valid) tile_order_xy<true>(k, s.width, s.height, px, py);
    const int gid = (int)(py * (unsigned)s.width + px);
    uint32_t rng = (uint32_t)seed + (uint32_t)gid;
    rng_next(rng);
    float3 o, d;
    camera_ray_xy<false>(s, gid, (int)px, (int)py, rng, o, d);
    if (MODE == 0) {
        Record rec;
        rec.distance = inff_(); rec.material = 0; rec.surf.normal = f3(0, 0, 0); rec.point = f3(0, 0, 0);
        rec.surf.color = make_float4(0, 0, 0, 0); rec.surf.emittance = 0;
        HitInfo hi = {-1, 0, 0, 0, 0};
        bool hit = false;
        if (valid) hit = HAS_BVH ? closest_intersect_ref(s, o, d, rec, hi) : octree_intersect_ref(s, o, d, rec, hi);
        if (valid) first_hit_store(out, gid, hit, rec.material, rec.surf.normal, hi.node, hi.kind, rec.distance, rec.surf.color);
        return;
    }
    March m;
    const bool entered = march_begin(s, m, o, d, inff_());
    LeanRay r;
    r.o = m.o; r.d = m.d; r.inv = m.inv; r.t = m.t; r.limit = m.limit; r.steps = m.steps;
    lean_prepare(r);
    int st = (valid && entered) ? 0 : 2;      // 0 marching, 1 at a non-air leaf, 2 left the octree without a hit, 3 hit
    // what an octree hit leaves behind; without BVHs it is written out at once, so that nothing of it stays in registers while
    // the other lanes of the warp keep marching
    Surf surf;
    surf.normal = f3(0, 0, 0); surf.color = make_float4(0, 0, 0, 0); surf.emittance = 0;
    float distance = inff_();
    int material = 0, hbx = 0, hby = 0, hbz = 0;
    for (;;) {
        st = first_hit_march<MODE == 2>(s, &r, st);
        if (st == 1) {
            m.t = r.t; m.steps = r.steps;
            const Cell c = march_cell(m);
            int level;
            const int data = find_leaf_wide(s, c.bx, c.by, c.bz, level);
            float th;
            Surf hs;
            if (march_block(s, m, data, level, hs, th)) {
                st = 3;
                if (!HAS_BVH) {
                    int hnode = -1;
                    if (out.node) { int lv; find_leaf(s, c.bx, c.by, c.bz, lv, hnode); }
                    first_hit_store(out, gid, true, data, hs.normal, hnode, 1, th, hs.color);
                } else {
                    surf = hs; distance = th; material = data;
                    hbx = c.bx; hby = c.by; hbz = c.bz;
                }
            } else {
                r.t = m.t; r.steps = m.steps;
                st = 0;
            }
        }
        if (!__any_sync(full, st == 0)) break;
    }
    if (!valid) return;
    if (!HAS_BVH) {
        if (st != 3) first_hit_store(out, gid, false, 0, f3(0, 0, 0), -1, 0, inff_(), make_float4(0, 0, 0, 0));
        return;
    }
    bool hit = st == 3;
    int kind = hit ? 1 : 0, bk = 0;
    if (bvh_pair(s, o, d, distance, surf, bk)) { hit = true; kind = bk; }
    int hnode = -1;
    if (out.node && hit && kind == 1) { int lv; find_leaf(s, hbx, hby, hbz, lv, hnode); }
    first_hit_store(out, gid, hit, material, surf.normal, hnode, kind, distance, surf.color);
}

// rayTracer.cl:141-216
template <int MODE>
__global__ void __launch_bounds__(256) k_preview(const __grid_constant__ DScene s, int n_pixels, int *res) {
    stage_tables(s, nullptr);
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (unsigned)n_pixels) return;
    const int gid = tile_order_pixel<true>(k, s.width, s.height);
    int W = s.width, H = s.height;
    int px = gid % W, py = gid / W;
    if ((px == W / 2 && (py >= H / 2 - 5 && py <= H / 2 + 5)) || (py == H / 2 && (px >= W / 2 - 5 && px <= W / 2 + 5))) {
        res[gid] = (int)0xFFFFFFFFu;
        return;
    }
    uint32_t rng = 0;
    rng_next(rng);
    float3 o, d;
    camera_ray<true>(s, gid, rng, o, d);
    Record rec;
    rec.distance = inff_(); rec.material = 0; rec.surf.normal = f3(0, 0, 0); rec.point = f3(0, 0, 0);
    rec.surf.color = make_float4(0, 0, 0, 0); rec.surf.emittance = 0;
    HitInfo hi = {-1, 0, 0, 0, 0};
    float3 c;
    if (closest_intersect_mode<MODE>(s, o, d, rec, hi)) {
        float shading = dot3(rec.surf.normal, f3(0.25f, 0.866f, 0.433f));
        shading = fmaxf(0.3f, shading);
        c = f3(rec.surf.color.x * shading, rec.surf.color.y * shading, rec.surf.color.z * shading);
    } else {
        c = sky_radiance(s, d);
    }
    float v[3] = {c.x, c.y, c.z};
    int rgb[3];
    for (int i = 0; i < 3; i++) {
        float q = sqrtf(v[i]) * 255.0f;
        q = fminf(fmaxf(q, 0.0f), 255.0f);
        rgb[i] = f2i(floorf(q));
    }
    res[gid] = (int)(0xFF000000u | ((uint32_t)rgb[0] << 16) | ((uint32_t)rgb[1] << 8) | (uint32_t)rgb[2]);
}

// post_processing_filter.cl:5-51: one work-item per pixel
__global__ void __launch_bounds__(256) k_tonemap(int n_pixels, float exposure, const double *__restrict__ input, int type, uint32_t *__restrict__ res) {
    for (int gid = blockIdx.x * blockDim.x + threadIdx.x; gid < n_pixels; gid += gridDim.x * blockDim.x)
        res[gid] = tonemap_pixel(input + (size_t)gid * 3, exposure, type);
}

__global__ void k_unorm_table(float *t) { t[threadIdx.x] = (float)threadIdx.x / 255.0f; }

__global__ void k_scale(float *buf, float factor, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] *= factor;
}

// Re-tile a row-major RGBA8 rectangle into the tile-linear atlas (16x16 texel tiles contiguous = 1 KiB each).
__global__ void k_atlas_write(uchar4 *atlas, int tiles_x, int tiles_y, int x0, int y0, int layer, int w, int h, const uchar4 *src) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * h) return;
    int x = x0 + i % w, y = y0 + i / w;
    size_t tile = ((size_t)layer * tiles_y + (y >> 4)) * tiles_x + (x >> 4);
    atlas[tile * 256 + ((y & 15) << 4) + (x & 15)] = src[i];
}

// Roofline denominators this path needs: random 32-byte-sector gathers.  Every thread issues `loads`
// independent 16-byte L2 loads (ld.global.cg) at hashed sector addresses of an array of `sectors` sectors
// (dependent = 0), or walks a dependent chain (dependent = 1: the next address comes from the loaded word).
__global__ void k_gather_fill(uint4 *a, size_t sectors, uint32_t seed) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < sectors * 2; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        a[i] = make_uint4(h, h * 3u + 1u, h ^ 0x9e3779b9u, (uint32_t)i);
    }
}
__global__ void __launch_bounds__(256) k_gather(const uint4 *__restrict__ a, uint32_t sector_mask, int loads, int dependent, uint32_t salt, uint32_t *sink) {
    // salt: every timed launch walks different sectors, so that a repeat does not find its own footprint in L2
    uint32_t st = (blockIdx.x * blockDim.x + threadIdx.x + salt * 0x9e3779b9u) * 747796405u + 2891336453u;
    uint32_t acc = 0;
    if (dependent) {
        // one thread per SM walks the chain: the time per load is the latency of one dependent gather
        if (threadIdx.x != 0) return;
        uint32_t at = st & sector_mask;
        for (int i = 0; i < loads; i++) {
            uint4 v = __ldcg(a + (size_t)at * 2);
            acc ^= v.y;
            at = (v.x ^ (uint32_t)i * 0x9e3779b9u) & sector_mask;
        }
    } else {
#pragma unroll 8
        for (int i = 0; i < loads; i++) {
            st = st * 747796405u + 2891336453u;
            uint32_t at = ((st >> 9) ^ st) & sector_mask;
            uint4 v = __ldcg(a + (size_t)at * 2);
            acc ^= v.x ^ v.w;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

// ======================================================================================================
// host side
// ======================================================================================================
namespace ccu_host {
static thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
const char *last_error() { return g_err.c_str(); }
}  // namespace ccu_host

using ccu_host::DevBuf;
using ccu_host::DeviceGuard;
using ccu_host::fail;
using ccu_host::SEED_SLOTS;
using ccu_host::SEED_SLOT_INTS;

namespace {

bool bvh_is_empty(const std::vector<int> &bvh) {   // bvh.h:23-32
    if (bvh.size() < 7) return true;
    if (bvh[0] != 0) return false;
    for (int i = 1; i < 7; i++) {
        float f;
        memcpy(&f, &bvh[i], 4);
        if (!std::isnan(f)) return false;
    }
    return true;
}

int air_layout_id(const ccu_ctx *c) {   // template parameter LAY of k_render_queue
    if (c->air_deep) return 2;
    return (c->air_top.n <= (size_t)Q_TOP_WORDS && getenv("CCU_NO_TOPS") == nullptr) ? 0 : 1;
}

void fill_scene(ccu_ctx *c) {
    DScene &s = c->scene;
    s.tree = c->tree.p;
    s.depth = c->depth;
    s.top = c->top.p;
    s.wide = c->wide.p;
    s.cell_level = c->cell_level;
    s.top_log2 = c->top_log2;
    s.use_wide = c->use_wide;
    s.air_top = c->air_top.p;
    s.air_wide = c->air_wide.p;
    s.air_bricks = c->air_bricks.p;
    s.air_cell_level = c->air_cell_level;
    s.air_top_log2 = c->air_top_log2;
    s.world_rec = reinterpret_cast<const int4 *>(c->world_rec.p);
    s.actor_rec = reinterpret_cast<const int4 *>(c->actor_rec.p);
    s.tris2 = c->tris2.p;
    s.world_root = c->world_root;
    s.actor_root = c->actor_root;
    s.block_palette = c->block_palette.p;
    s.block_palette_len = (int)c->block_palette.n;
    s.quad_models = c->quad_models.p;
    s.aabb_models = c->aabb_models.p;
    s.mat_palette = c->mat_palette.p;
    s.block_rec = reinterpret_cast<const int4 *>(c->block_rec.p);
    s.mat_rec = reinterpret_cast<const int4 *>(c->mat_rec.p);
    s.quad_rec = reinterpret_cast<const int4 *>(c->quad_rec.p);
    s.aabb_rec = reinterpret_cast<const int4 *>(c->aabb_rec.p);
    s.world_bvh = c->world_bvh.p;
    s.actor_bvh = c->actor_bvh.p;
    s.trigs = c->trigs.p;
    const bool no_entities = (c->params.flags & CCU_RENDER_NO_ENTITIES) != 0;
    s.world_bvh_empty = no_entities || bvh_is_empty(c->world_host);
    s.actor_bvh_empty = no_entities || bvh_is_empty(c->actor_host);
    s.atlas = c->atlas.p;
    s.atlas_w = c->atlas_w; s.atlas_h = c->atlas_h; s.atlas_layers = c->atlas_layers;
    s.atlas_tiles_x = (c->atlas_w + 15) / 16; s.atlas_tiles_y = (c->atlas_h + 15) / 16;
    s.sky = c->sky.p;
    s.sky_res = c->sky_res;
    s.sky_intensity = c->sky_intensity;
    s.unorm = c->unorm;
    s.sun_flags = c->sun_host[0]; s.sun_tex_size = c->sun_host[1]; s.sun_tex = c->sun_host[2];
    if (c->params.flags & CCU_RENDER_NO_SUN) s.sun_flags &= ~1;
    memcpy(&s.sun_intensity, &c->sun_host[3], 4);
    s.projector_type = c->projector_type;
    memcpy(s.cam, c->cam, sizeof s.cam);
    s.rays = c->rays[c->rays_active].p;
    s.width = c->width; s.height = c->height;
    if (c->height > 0) {
        s.half_width = (float)(c->width / (2.0 * c->height));   // rayTracer.cl:66
        s.inv_height = (float)(1.0 / c->height);                // rayTracer.cl:67
    }
    s.draw_depth = c->params.draw_depth; s.max_depth = c->params.max_depth; s.emitter_scale = c->params.emitter_scale;
}

// Upload one of the reference's packed int arrays; `after` runs under the context lock once the upload succeeded
// (host-side copies and flags change together with the device buffer).
template <class F>
int upload_words(ccu_ctx *c, DevBuf<int> ccu_ctx::*dst, const int32_t *words, int64_t n, const char *what, F after) {
    if (!c) return fail(CCU_EINVAL, "%s: null context", what);
    if (n < 0 || (n > 0 && !words)) return fail(CCU_EINVAL, "%s: bad array (n=%lld)", what, (long long)n);
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    c->committed = false;
    CU((c->*dst).upload(words, (size_t)n, c->stream));
    after();
    return CCU_OK;
}

// device time of the last render / first-hit / preview / tonemap launch(es); waits for them (caller holds c->mu)
void stop_timer(ccu_ctx *c) {
    if (c->timing_pending) {
        cudaEventSynchronize(c->ev1);
        cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1);
        c->timing_pending = false;
    }
}

// waits for everything queued on the render stream WITHOUT holding the context lock
int wait_render_stream(ccu_ctx *c) {
    cudaEvent_t ev = nullptr;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        DeviceGuard g(c->device);
        CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        cudaError_t e = cudaEventRecord(ev, c->stream);
        if (e != cudaSuccess) { cudaEventDestroy(ev); return fail(CCU_ECUDA, "cudaEventRecord: %s", cudaGetErrorString(e)); }
    }
    cudaError_t e = cudaEventSynchronize(ev);
    cudaEventDestroy(ev);
    if (e != cudaSuccess) return fail(CCU_ECUDA, "render stream: %s", cudaGetErrorString(e));
    return CCU_OK;
}

int join_merge_locked(ccu_ctx *c, std::unique_lock<std::mutex> &lk) {
    if (!c->merge.active) return CCU_OK;
    std::thread t = std::move(c->merge.worker);
    lk.unlock();
    if (t.joinable()) t.join();
    lk.lock();
    c->merge.active = false;
    if (c->merge.status != CCU_OK) {
        const int rc = c->merge.status;
        c->merge.status = CCU_OK;
        return fail(rc, "%s", c->merge.error.c_str());
    }
    return CCU_OK;
}

void free_target(ccu_ctx *c) {
    for (int i = 0; i < 2; i++) {
        if (c->accum[i]) cudaFree(c->accum[i]);
        c->accum[i] = nullptr;
    }
    if (c->pinned) cudaFreeHost(c->pinned);
    c->pinned = nullptr;
    c->accum_floats = 0;
}

template <class K>
cudaError_t allow_smem(K kernel, int bytes) { return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); }

}  // namespace

namespace ccu_host {

// Read-back of a finished window, in two phases so that the copy overlaps whatever the host does in between:
// start_readback queues the copy of floats [lo, hi) of `src_dev` (indexed by absolute float offset) into the pinned staging
// buffer in chunks on `st`, one event per chunk; merge_readback waits chunk by chunk and folds each into the host sample buffer
// (OpenClPathTracingRenderer.java:164-173):  sample[i] = (sample[i] * ds + mean[i] * dp) * sinv
constexpr int MERGE_CHUNKS = 8;
static size_t merge_chunk_len(size_t n) { return ((n + MERGE_CHUNKS - 1) / MERGE_CHUNKS + 63) & ~(size_t)63; }

int start_readback(ccu_ctx *c, const float *src_dev, size_t lo, size_t hi, cudaStream_t st) {
    const size_t n = hi > lo ? hi - lo : 0;
    const size_t chunk = merge_chunk_len(n);
    for (int k = 0; k < MERGE_CHUNKS; k++) {
        const size_t a = lo + std::min(n, k * chunk), b = lo + std::min(n, (k + 1) * chunk);
        if (b > a) CU(cudaMemcpyAsync(c->pinned + a, src_dev + a, (b - a) * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(c->chunk_ev[k], st));
    }
    return CCU_OK;
}

int merge_readback(ccu_ctx *c, size_t lo, size_t hi, double *sample_buffer, double ds, double dp, double sinv, unsigned max_threads) {
    if (hi <= lo) return CCU_OK;
    const size_t n = hi - lo;
    const size_t chunk = merge_chunk_len(n);
    unsigned hw = std::thread::hardware_concurrency();
    unsigned nt = std::max(1u, std::min(std::min(16u, max_threads), hw ? hw : 4u));
    if (n < (1u << 20)) nt = 1;
    const float *src = c->pinned;
    cudaEvent_t *evs = c->chunk_ev;
    const int device = c->device;
    std::vector<cudaError_t> errs(nt, cudaSuccess);
    auto work = [&, device](unsigned t) {
        cudaSetDevice(device);
        for (int k = 0; k < MERGE_CHUNKS; k++) {
            const size_t a = lo + std::min(n, k * chunk), b = lo + std::min(n, (k + 1) * chunk);
            const cudaError_t e = cudaEventSynchronize(evs[k]);
            if (e != cudaSuccess) { errs[t] = e; return; }
            const size_t len = b - a, part = (len + nt - 1) / nt;
            const size_t i0 = a + std::min(len, t * part), i1 = a + std::min(len, (t + 1) * part);
            for (size_t i = i0; i < i1; i++) sample_buffer[i] = (sample_buffer[i] * ds + (double)src[i] * dp) * sinv;
        }
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; t++) th.emplace_back(work, t);
        for (auto &t : th) t.join();
    }
    for (cudaError_t e : errs)
        if (e != cudaSuccess) return fail(CCU_ECUDA, "window read-back: %s", cudaGetErrorString(e));
    return CCU_OK;
}

int merge_window_range(ccu_ctx *c, const float *src_dev, size_t lo, size_t hi, double *sample_buffer, double ds, double dp, double sinv,
                       cudaStream_t st, unsigned max_threads) {
    int rc = start_readback(c, src_dev, lo, hi, st);
    if (rc != CCU_OK) return rc;
    return merge_readback(c, lo, hi, sample_buffer, ds, dp, sinv, max_threads);
}

template <class T>
static int copy_buf(ccu_ctx *src, ccu_ctx *dst, DevBuf<T> ccu_ctx::*member) {
    DevBuf<T> &a = src->*member, &b = dst->*member;
    if (!a.p) { b.release(); return CCU_OK; }
    CU(b.alloc(a.n));
    CU(cudaMemcpyPeerAsync(b.p, dst->device, a.p, src->device, a.bytes(), dst->stream));
    return CCU_OK;
}

// Copy the committed scene of `src` (uploaded arrays and commit-time layouts) to `dst` device-to-device.
int replicate_scene(ccu_ctx *src, ccu_ctx *dst) {
    if (src == dst) return CCU_OK;
    std::lock(src->mu, dst->mu);
    std::lock_guard<std::mutex> l1(src->mu, std::adopt_lock), l2(dst->mu, std::adopt_lock);
    if (!src->committed) return fail(CCU_ESTATE, "replicate: source scene not committed");
    DeviceGuard g(dst->device);
    cudaStreamSynchronize(dst->stream);
    int rc = CCU_OK;
#define CP(m) if (rc == CCU_OK) rc = copy_buf(src, dst, &ccu_ctx::m)
    CP(tree); CP(block_palette); CP(quad_models); CP(aabb_models); CP(mat_palette); CP(trigs); CP(world_bvh); CP(actor_bvh); CP(sun_words);
    CP(top); CP(wide); CP(air_top); CP(air_wide); CP(air_bricks); CP(world_rec); CP(actor_rec); CP(tris2); CP(block_rec); CP(mat_rec); CP(quad_rec); CP(aabb_rec); CP(atlas); CP(sky); CP(sun_basis);
#undef CP
    if (rc != CCU_OK) return rc;
    dst->world_host = src->world_host.size() >= 7 ? std::vector<int>(src->world_host.begin(), src->world_host.begin() + 7) : src->world_host;
    dst->actor_host = src->actor_host.size() >= 7 ? std::vector<int>(src->actor_host.begin(), src->actor_host.begin() + 7) : src->actor_host;
    dst->tree_host.clear(); dst->trigs_host.clear(); dst->block_host.clear(); dst->mat_host.clear();
    dst->world_root = src->world_root; dst->actor_root = src->actor_root; dst->use_bvh2 = src->use_bvh2; dst->use_air = src->use_air;
    dst->air_deep = src->air_deep; dst->cell_level = src->cell_level; dst->top_log2 = src->top_log2; dst->use_wide = src->use_wide;
    dst->air_cell_level = src->air_cell_level; dst->air_top_log2 = src->air_top_log2;
    dst->atlas_w = src->atlas_w; dst->atlas_h = src->atlas_h; dst->atlas_layers = src->atlas_layers;
    dst->depth = src->depth; dst->sky_res = src->sky_res; dst->sky_intensity = src->sky_intensity;
    memcpy(dst->sun_host, src->sun_host, sizeof dst->sun_host);
    dst->have_sun = dst->have_octree = dst->have_blocks = dst->have_mats = dst->have_atlas = dst->have_sky = true;
    dst->scene.su = src->scene.su; dst->scene.sv = src->scene.sv; dst->scene.sw = src->scene.sw; dst->scene.sun_radius_cos = src->scene.sun_radius_cos;
    CU(cudaStreamSynchronize(dst->stream));
    dst->committed = true;
    return CCU_OK;
}

}  // namespace ccu_host

extern "C" {

const char *ccu_last_error(void) { return ccu_host::last_error(); }
const char *ccu_version(void) { return "chunkycu 0.2 (sm_100a)"; }

int ccu_device_count(int *count) {
    if (!count) return fail(CCU_EINVAL, "ccu_device_count: null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(CCU_ENODEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    }
    *count = n;
    return CCU_OK;
}

int ccu_device_info(int index, char *name, int name_len, int *sm_count, int *clock_khz, uint64_t *mem_bytes) {
    int n = 0;
    int rc = ccu_device_count(&n);
    if (rc != CCU_OK) return rc;
    if (index < 0 || index >= n) return fail(CCU_EINVAL, "device index %d out of range [0,%d)", index, n);
    cudaDeviceProp p;
    CU(cudaGetDeviceProperties(&p, index));
    if (name && name_len > 0) snprintf(name, (size_t)name_len, "%s", p.name);
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (clock_khz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, index);
        *clock_khz = khz;
    }
    if (mem_bytes) *mem_bytes = (uint64_t)p.totalGlobalMem;
    return CCU_OK;
}

int ccu_ctx_create(int device_index, ccu_ctx **out) {
    if (!out) return fail(CCU_EINVAL, "ccu_ctx_create: null out");
    *out = nullptr;
    int n = 0;
    int rc = ccu_device_count(&n);
    if (rc != CCU_OK) return rc;
    if (n == 0) return fail(CCU_ENODEVICE, "no CUDA device");
    if (device_index < 0 || device_index >= n) return fail(CCU_EINVAL, "device index %d out of range [0,%d)", device_index, n);
    cudaDeviceProp p;
    CU(cudaGetDeviceProperties(&p, device_index));
    if (p.major != 10) return fail(CCU_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device_index, p.major, p.minor);
    ccu_ctx *c = new ccu_ctx();
    c->device = device_index;
    c->sm_count = p.multiProcessorCount;
    if (const char *e = getenv("CCU_Q_MARCH_WARPS")) c->q_march_warps = std::max(1, std::min(64, atoi(e)));
    if (const char *e = getenv("CCU_Q_BVH_WARPS")) c->q_bvh_warps = std::max(1, std::min(64, atoi(e)));
    if (const char *e = getenv("CCU_Q_LEAF_MIN")) c->q_leaf_min = std::max(1, std::min(32, atoi(e)));
    if (const char *e = getenv("CCU_Q_MARCH_BIAS")) c->q_march_bias = std::max(-32, std::min(32, atoi(e)));
    if (const char *e = getenv("CCU_Q_REFILL_MIN")) c->q_refill_min = std::max(1, std::min(32, atoi(e)));
    if (const char *e = getenv("CCU_Q_STICKY")) c->q_sticky_min = std::max(1, std::min(33, atoi(e)));
    if (const char *e = getenv("CCU_Q_SHADE_MIN")) c->q_shade_min = std::max(0, std::min(32, atoi(e)));
    if (const char *e = getenv("CCU_YIELD_BELOW")) c->yield_below = std::max(0, std::min(33, atoi(e)));
    DeviceGuard g(device_index);
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->window_ev, cudaEventDisableTiming);
    for (int k = 0; k < 8 && e == cudaSuccess; k++) e = cudaEventCreateWithFlags(&c->chunk_ev[k], cudaEventDisableTiming);
    for (int k = 0; k < 2 && e == cudaSuccess; k++) e = cudaEventCreateWithFlags(&c->rays_used[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = allow_smem(k_render_queue<false, 0>, Q_SMEM_LIMIT);
    if (e == cudaSuccess) e = allow_smem(k_render_queue<false, 1>, Q_SMEM_LIMIT);
    if (e == cudaSuccess) e = allow_smem(k_render_queue<false, 2>, Q_SMEM_LIMIT);
    if (e == cudaSuccess) e = allow_smem(k_render_queue<true, 0>, Q_SMEM_LIMIT);
    if (e == cudaSuccess) e = allow_smem(k_render_queue<true, 1>, Q_SMEM_LIMIT);
    if (e == cudaSuccess) e = allow_smem(k_render_queue<true, 2>, Q_SMEM_LIMIT);
    if (e == cudaSuccess) e = cudaMalloc(&c->unorm, 256 * sizeof(float));
    if (e == cudaSuccess) {
        k_unorm_table<<<1, 256, 0, c->stream>>>(c->unorm);
        c->launches++;
        e = cudaStreamSynchronize(c->stream);
    }
    if (e != cudaSuccess) {
        delete c;
        return fail(CCU_ECUDA, "context setup: %s", cudaGetErrorString(e));
    }
    *out = c;
    return CCU_OK;
}

int ccu_ctx_destroy(ccu_ctx *c) {
    if (!c) return CCU_OK;
    {
        std::unique_lock<std::mutex> lk(c->mu);
        join_merge_locked(c, lk);
        DeviceGuard g(c->device);
        cudaStreamSynchronize(c->stream);
        cudaStreamSynchronize(c->copy_stream);
        c->tree.release(); c->top.release(); c->wide.release(); c->air_top.release(); c->air_wide.release(); c->air_bricks.release();
        c->world_rec.release(); c->actor_rec.release(); c->tris2.release(); c->block_rec.release(); c->mat_rec.release(); c->quad_rec.release(); c->aabb_rec.release(); c->block_palette.release();
        c->quad_models.release(); c->aabb_models.release(); c->mat_palette.release(); c->trigs.release(); c->world_bvh.release();
        c->actor_bvh.release(); c->sun_words.release(); c->atlas.release(); c->sky.release(); c->rays[0].release(); c->rays[1].release();
        c->sun_basis.release();
        free_target(c);
        if (c->seeds_dev) cudaFree(c->seeds_dev);
        if (c->seeds_pinned) cudaFreeHost(c->seeds_pinned);
        for (auto &e : c->seeds_ev) if (e) cudaEventDestroy(e);
        if (c->fh_scratch) cudaFree(c->fh_scratch);
        if (c->bvh_deep) cudaFree(c->bvh_deep);
        if (c->work_counter) cudaFree(c->work_counter);
        if (c->unorm) cudaFree(c->unorm);
        for (auto &e : c->chunk_ev) if (e) cudaEventDestroy(e);
        for (auto &e : c->rays_used) if (e) cudaEventDestroy(e);
        if (c->window_ev) cudaEventDestroy(c->window_ev);
        cudaEventDestroy(c->ev0);
        cudaEventDestroy(c->ev1);
        cudaStreamDestroy(c->copy_stream);
        cudaStreamDestroy(c->stream);
    }
    delete c;
    return CCU_OK;
}

// A new scene starts from nothing, as the reference's loader does (fresh palettes and EMPTY_NODE BVHs on every load,
// AbstractSceneLoader.java:70-140): arrays of the previous scene do not leak into a scene that does not set them.
int ccu_scene_begin(ccu_ctx *c) {
    if (!c) return fail(CCU_EINVAL, "ccu_scene_begin: null context");
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    cudaStreamSynchronize(c->stream);
    c->committed = false;
    c->quad_models.release(); c->aabb_models.release(); c->trigs.release(); c->world_bvh.release(); c->actor_bvh.release();
    c->world_host.clear(); c->actor_host.clear(); c->trigs_host.clear(); c->quad_host.clear(); c->aabb_host.clear();
    c->have_octree = c->have_blocks = c->have_mats = c->have_atlas = c->have_sky = c->have_sun = false;
    return CCU_OK;
}

int ccu_scene_set_octree(ccu_ctx *c, const int32_t *tree, int64_t n, int32_t depth) {
    if (depth < 0 || depth > 30) return fail(CCU_EINVAL, "octree depth %d out of range", depth);
    if (n < 1) return fail(CCU_EINVAL, "octree needs at least the root word");
    return upload_words(c, &ccu_ctx::tree, tree, n, "ccu_scene_set_octree", [&] {
        c->depth = depth;
        c->tree_host.assign(tree, tree + n);
        c->have_octree = true;
    });
}
int ccu_scene_set_block_palette(ccu_ctx *c, const int32_t *w, int64_t n) {
    return upload_words(c, &ccu_ctx::block_palette, w, n, "ccu_scene_set_block_palette", [&] { c->block_host.assign(w, w + n); c->have_blocks = true; });
}
int ccu_scene_set_quad_models(ccu_ctx *c, const int32_t *w, int64_t n) {
    return upload_words(c, &ccu_ctx::quad_models, w, n, "ccu_scene_set_quad_models", [&] { c->quad_host.assign(w, w + n); });
}
int ccu_scene_set_aabb_models(ccu_ctx *c, const int32_t *w, int64_t n) {
    return upload_words(c, &ccu_ctx::aabb_models, w, n, "ccu_scene_set_aabb_models", [&] { c->aabb_host.assign(w, w + n); });
}
int ccu_scene_set_material_palette(ccu_ctx *c, const int32_t *w, int64_t n) {
    return upload_words(c, &ccu_ctx::mat_palette, w, n, "ccu_scene_set_material_palette", [&] { c->mat_host.assign(w, w + n); c->have_mats = true; });
}
int ccu_scene_set_triangles(ccu_ctx *c, const int32_t *w, int64_t n) {
    return upload_words(c, &ccu_ctx::trigs, w, n, "ccu_scene_set_triangles", [&] { c->trigs_host.assign(w, w + n); });
}
int ccu_scene_set_world_bvh(ccu_ctx *c, const int32_t *w, int64_t n) {
    return upload_words(c, &ccu_ctx::world_bvh, w, n, "ccu_scene_set_world_bvh", [&] { c->world_host.assign(w, w + n); });
}
int ccu_scene_set_actor_bvh(ccu_ctx *c, const int32_t *w, int64_t n) {
    return upload_words(c, &ccu_ctx::actor_bvh, w, n, "ccu_scene_set_actor_bvh", [&] { c->actor_host.assign(w, w + n); });
}

int ccu_scene_atlas_create(ccu_ctx *c, int32_t width, int32_t height, int32_t layers) {
    if (!c) return fail(CCU_EINVAL, "ccu_scene_atlas_create: null context");
    if (width <= 0 || height <= 0 || layers <= 0 || width > 16384 || height > 16384) return fail(CCU_EINVAL, "bad atlas extents %dx%dx%d", width, height, layers);
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    c->committed = false;
    c->have_atlas = false;
    const size_t tiles = (size_t)((width + 15) / 16) * ((height + 15) / 16) * layers;
    CU(c->atlas.alloc(tiles * 256));
    CU(cudaMemsetAsync(c->atlas.p, 0, c->atlas.n * sizeof(uchar4), c->stream));
    c->atlas_w = width; c->atlas_h = height; c->atlas_layers = layers;
    c->have_atlas = true;
    return CCU_OK;
}

int ccu_scene_atlas_write(ccu_ctx *c, int32_t x, int32_t y, int32_t layer, int32_t w, int32_t h, const uint8_t *rgba) {
    if (!c) return fail(CCU_EINVAL, "ccu_scene_atlas_write: null context");
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->atlas.p || !c->have_atlas) return fail(CCU_ESTATE, "atlas not created");
    if (!rgba || w <= 0 || h <= 0 || x < 0 || y < 0 || layer < 0 || x + w > c->atlas_w || y + h > c->atlas_h || layer >= c->atlas_layers)
        return fail(CCU_EINVAL, "atlas write %dx%d at (%d,%d,%d) outside %dx%dx%d", w, h, x, y, layer, c->atlas_w, c->atlas_h, c->atlas_layers);
    DeviceGuard g(c->device);
    c->committed = false;
    uchar4 *tmp = nullptr;
    size_t bytes = (size_t)w * h * 4;
    CU(cudaMalloc(&tmp, bytes));
    cudaError_t e = cudaMemcpyAsync(tmp, rgba, bytes, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        int n = w * h;
        k_atlas_write<<<(n + 255) / 256, 256, 0, c->stream>>>(c->atlas.p, (c->atlas_w + 15) / 16, (c->atlas_h + 15) / 16, x, y, layer, w, h, tmp);
        c->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(CCU_ECUDA, "atlas write: %s", cudaGetErrorString(e));
    return CCU_OK;
}

int ccu_scene_set_atlas(ccu_ctx *c, const uint8_t *rgba, int32_t width, int32_t height, int32_t layers) {
    int rc = ccu_scene_atlas_create(c, width, height, layers);
    if (rc != CCU_OK) return rc;
    for (int l = 0; l < layers; l++) {
        rc = ccu_scene_atlas_write(c, 0, 0, l, width, height, rgba + (size_t)l * width * height * 4);
        if (rc != CCU_OK) return rc;
    }
    return CCU_OK;
}

int ccu_scene_set_sky(ccu_ctx *c, const uint8_t *rgba, int32_t res, float sky_intensity) {
    if (!c) return fail(CCU_EINVAL, "ccu_scene_set_sky: null context");
    if (!rgba || res <= 0 || res > 16384) return fail(CCU_EINVAL, "bad sky texture (res=%d)", res);
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    c->committed = false;
    CU(c->sky.upload(reinterpret_cast<const uchar4 *>(rgba), (size_t)res * res, c->stream));
    c->sky_res = res;
    c->sky_intensity = sky_intensity;
    c->have_sky = true;
    return CCU_OK;
}

int ccu_scene_set_sun(ccu_ctx *c, const int32_t sun_words[6]) {
    if (!c || !sun_words) return fail(CCU_EINVAL, "ccu_scene_set_sun: null argument");
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    c->committed = false;
    memcpy(c->sun_host, sun_words, sizeof c->sun_host);
    CU(c->sun_words.upload(sun_words, 6, c->stream));
    c->have_sun = true;
    return CCU_OK;
}

int ccu_scene_commit(ccu_ctx *c) {
    if (!c) return fail(CCU_EINVAL, "ccu_scene_commit: null context");
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->tree.p || !c->have_octree) return fail(CCU_ESTATE, "commit: octree not set");
    if (!c->have_blocks || !c->have_mats) return fail(CCU_ESTATE, "commit: block/material palette not set");
    if (!c->have_atlas) return fail(CCU_ESTATE, "commit: texture atlas not set");
    if (!c->have_sky) return fail(CCU_ESTATE, "commit: sky not set");
    if (!c->have_sun) return fail(CCU_ESTATE, "commit: sun not set");
    DeviceGuard g(c->device);
    const auto t0 = std::chrono::steady_clock::now();
    // absent optional palettes behave as the reference's single zero word / EMPTY_NODE
    int zero = 0;
    if (!c->quad_models.p) { CU(c->quad_models.upload(&zero, 0, c->stream)); c->quad_host.clear(); }
    if (!c->aabb_models.p) { CU(c->aabb_models.upload(&zero, 0, c->stream)); c->aabb_host.clear(); }
    if (!c->trigs.p) { CU(c->trigs.upload(&zero, 0, c->stream)); c->trigs_host.clear(); }
    if (!c->world_bvh.p) { CU(c->world_bvh.upload(&zero, 0, c->stream)); c->world_host.clear(); }
    if (!c->actor_bvh.p) { CU(c->actor_bvh.upload(&zero, 0, c->stream)); c->actor_host.clear(); }
    // octree layouts: value-carrying (top table + 64-ary nodes) and march layout (top table + bricks); a tree that cannot be
    // encoded (malformed, leaf values beyond 26 bits) leaves the flags at 0 and rendering uses the reference's own array
    {
        const bool enable = getenv("CCU_NO_WIDE") == nullptr;
        WideLayout wl = build_wide_layout(c->tree_host.data(), c->tree_host.size(), c->depth);
        c->use_wide = (enable && wl.ok) ? 1 : 0;
        c->cell_level = wl.cell_level;
        c->top_log2 = wl.top_log2;
        CU(c->top.upload(wl.top.data(), wl.top.size(), c->stream));
        CU(c->wide.upload(wl.wide.data(), wl.wide.size(), c->stream));
        AirLayout al = build_air_layout(c->tree_host.data(), c->tree_host.size(), c->depth);
        c->use_air = (enable && al.ok) ? 1 : 0;
        c->air_cell_level = al.cell_level;
        c->air_top_log2 = al.top_log2;
        c->air_deep = al.cell_level > 4 ? 1 : 0;
        CU(c->air_top.upload(al.top.data(), al.top.size(), c->stream));
        CU(c->air_wide.upload(al.wide.data(), al.wide.size(), c->stream));
        CU(c->air_bricks.upload(al.bricks.data(), al.bricks.size(), c->stream));
    }
    // 16-byte-vectorised block / material / model palettes
    {
        PaletteRecs pr = build_palette_recs(c->block_host, c->mat_host, c->quad_host.data(), c->quad_host.size(), c->aabb_host.data(), c->aabb_host.size());
        CU(c->block_rec.upload(pr.block.data(), pr.block.size(), c->stream));
        CU(c->mat_rec.upload(pr.mat.data(), pr.mat.size(), c->stream));
        CU(c->quad_rec.upload(pr.quad.data(), pr.quad.size(), c->stream));
        CU(c->aabb_rec.upload(pr.aabb.data(), pr.aabb.size(), c->stream));
    }
    // BVH stage layout (pair records + aligned triangle blocks); a BVH it cannot hold (malformed, or deeper than the 64 entries
    // of the reference's traversal stack, bvh.h:38) is rendered by the thread-per-pixel kernel on the reference's own arrays
    {
        BvhLayout wb, ab;
        TriRepack tr;
        std::unordered_map<int, int> leaf_map;
        const bool we = bvh_is_empty(c->world_host), ae = bvh_is_empty(c->actor_host);
        if (!we) wb.root = bvh_ref(c->world_host, c->trigs_host, 0, 0, wb, tr, leaf_map);
        if (!ae) ab.root = bvh_ref(c->actor_host, c->trigs_host, 0, 0, ab, tr, leaf_map);
        c->use_bvh2 = (wb.ok && ab.ok) ? 1 : 0;
        c->world_root = wb.root;
        c->actor_root = ab.root;
        if (!c->use_bvh2) { wb.rec.clear(); ab.rec.clear(); tr.tris.clear(); }
        tr.tris.insert(tr.tris.end(), 8, 0);      // the leaf stage reads the 32 bytes behind a count word before it looks at the count
        CU(c->world_rec.upload(wb.rec.data(), wb.rec.size(), c->stream));
        CU(c->actor_rec.upload(ab.rec.data(), ab.rec.size(), c->stream));
        CU(c->tris2.upload(tr.tris.data(), tr.tris.size(), c->stream));
    }
    // sun basis on the device
    if (!c->sun_basis.p) {
        float z[10] = {0};
        CU(c->sun_basis.upload(z, 10, c->stream));
    }
    k_sun_setup<<<1, 1, 0, c->stream>>>(c->sun_words.p, c->sun_basis.p);
    c->launches++;
    CU(cudaGetLastError());
    float b[10];
    CU(cudaMemcpyAsync(b, c->sun_basis.p, sizeof b, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    DScene &s = c->scene;
    s.su = make_float3(b[0], b[1], b[2]);
    s.sv = make_float3(b[3], b[4], b[5]);
    s.sw = make_float3(b[6], b[7], b[8]);
    s.sun_radius_cos = b[9];
    c->commit_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    c->committed = true;
    return CCU_OK;
}

// ClCamera.java:33-70 (settings) / :72-105 (pre-generated rays).  Ray uploads go to the buffer no launch is reading, on the
// copy stream, so that they overlap the passes in flight (the reference's cameraGenTask, OpenClPathTracingRenderer.java:146-148);
// the next ccu_render_passes uses the new rays.
int ccu_camera_set(ccu_ctx *c, int32_t projector_type, const float *settings, int64_t n) {
    if (!c || !settings) return fail(CCU_EINVAL, "ccu_camera_set: null argument");
    std::lock_guard<std::mutex> cam(c->cam_mu);
    if (projector_type != -1) {
        if (n < 15) return fail(CCU_EINVAL, "camera settings need 15 floats (got %lld)", (long long)n);
        std::lock_guard<std::mutex> lk(c->mu);
        memcpy(c->cam, settings, sizeof c->cam);
        c->projector_type = projector_type;
        c->have_camera = true;
        return CCU_OK;
    }
    if (n % 6 != 0 || n <= 0) return fail(CCU_EINVAL, "pre-generated rays need 6 floats per pixel (got %lld)", (long long)n);
    int target;
    cudaEvent_t used;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        target = (c->have_camera && c->projector_type == -1) ? 1 - c->rays_active : c->rays_active;
        used = c->rays_used[target];
    }
    DeviceGuard g(c->device);
    CU(cudaEventSynchronize(used));              // launches that read this buffer are done (outside the context lock)
    DevBuf<float> &buf = c->rays[target];        // only camera updates (serialised by cam_mu) touch the inactive buffer
    if (buf.n != (size_t)n) CU(buf.alloc((size_t)n));
    CU(cudaMemcpyAsync(buf.p, settings, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaStreamSynchronize(c->copy_stream));   // host memory is not retained past the call
    std::lock_guard<std::mutex> lk(c->mu);
    c->rays_active = target;
    c->projector_type = -1;
    c->have_camera = true;
    return CCU_OK;
}

int ccu_render_end(ccu_ctx *c) {
    if (!c) return fail(CCU_EINVAL, "ccu_render_end: null context");
    int rc = wait_render_stream(c);
    std::unique_lock<std::mutex> lk(c->mu);
    const int rc2 = join_merge_locked(c, lk);
    DeviceGuard g(c->device);
    stop_timer(c);
    // the buffers stay cached for the next render of the same size (freed by ccu_ctx_destroy / a size change)
    c->target_live = false;
    c->window_spp = 0;
    c->closed_spp = 0;
    return rc != CCU_OK ? rc : rc2;
}

int ccu_render_begin(ccu_ctx *c, int32_t width, int32_t height) {
    if (!c) return fail(CCU_EINVAL, "ccu_render_begin: null context");
    if (width <= 0 || height <= 0 || (int64_t)width * height > (1ll << 30)) return fail(CCU_EINVAL, "bad canvas %dx%d", width, height);
    std::unique_lock<std::mutex> lk(c->mu);
    int rc = join_merge_locked(c, lk);
    if (rc != CCU_OK) return rc;
    DeviceGuard g(c->device);
    cudaStreamSynchronize(c->stream);
    const size_t align = std::max<size_t>(1, c->accum_align);
    const size_t n = (((size_t)width * height * 3 + align - 1) / align) * align;
    if (c->accum[0] && c->accum_floats != n) free_target(c);
    if (!c->accum[0]) {
        cudaError_t e = cudaMalloc(&c->accum[0], n * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&c->accum[1], n * sizeof(float));
        if (e == cudaSuccess) e = cudaMallocHost(&c->pinned, n * sizeof(float));
        if (e != cudaSuccess) {
            free_target(c);       // nothing half-allocated survives a failed begin
            return fail(e == cudaErrorMemoryAllocation ? CCU_ENOMEM : CCU_ECUDA, "ccu_render_begin: %s", cudaGetErrorString(e));
        }
        c->accum_floats = n;
        CU(cudaMemsetAsync(c->accum[1], 0, n * sizeof(float), c->stream));
    }
    c->accum_active = 0;
    CU(cudaMemsetAsync(c->accum[0], 0, n * sizeof(float), c->stream));
    c->window_base = c->accum[0];
    c->width = width;
    c->height = height;
    c->window_spp = 0;
    c->closed_spp = 0;
    c->target_live = true;
    return CCU_OK;
}

int ccu_render_set_params(ccu_ctx *c, const ccu_render_params *p) {
    if (!c || !p) return fail(CCU_EINVAL, "ccu_render_set_params: null argument");
    if (p->draw_depth < 0 || p->draw_depth >= (1 << 24) || p->max_depth < 1 || p->max_depth > 255) return fail(CCU_EINVAL, "bad render params");
    if (p->kernel != 0 && p->kernel != 1 && p->kernel != 4) return fail(CCU_EINVAL, "unknown kernel %d (0 auto, 1 thread-per-pixel, 4 wavefront)", p->kernel);
    if (p->flags & ~(CCU_RENDER_NO_ENTITIES | CCU_RENDER_NO_SUN)) return fail(CCU_EINVAL, "unknown render flags 0x%x", p->flags);
    std::lock_guard<std::mutex> lk(c->mu);
    c->params = *p;
    return CCU_OK;
}

static int render_ready(ccu_ctx *c, const char *what) {
    if (!c->committed) return fail(CCU_ESTATE, "%s: scene not committed", what);
    if (!c->have_camera) return fail(CCU_ESTATE, "%s: camera not set", what);
    if (!c->accum[0] || !c->target_live) return fail(CCU_ESTATE, "%s: ccu_render_begin not called", what);
    if (c->projector_type == -1 && c->rays[c->rays_active].n != (size_t)c->width * c->height * 6)
        return fail(CCU_ESTATE, "%s: ray buffer holds %zu floats, canvas needs %zu", what, c->rays[c->rays_active].n, (size_t)c->width * c->height * 6);
    return CCU_OK;
}

// which implementation closestIntersect runs on: 0 = the reference's own arrays, 1 / 2 = commit-time layouts (2: deep world)
static int layout_mode(const ccu_ctx *c) { return (c->use_wide && c->use_air) ? (c->air_deep ? 2 : 1) : 0; }

static int render_passes_locked(ccu_ctx *c, const int32_t *seeds, int32_t n_passes, bool first_chunk) {
    // The seed words are the only per-pass host->device traffic (OpenClPathTracingRenderer.java:106-109).  They are staged in a
    // small ring of pinned slots so that the call neither retains the caller's array nor waits for the passes in flight.
    if (!c->seeds_dev) {
        CU(cudaMalloc(&c->seeds_dev, (size_t)SEED_SLOTS * SEED_SLOT_INTS * sizeof(int)));
        CU(cudaMallocHost(&c->seeds_pinned, (size_t)SEED_SLOTS * SEED_SLOT_INTS * sizeof(int)));
        for (int k = 0; k < SEED_SLOTS; k++) CU(cudaEventCreateWithFlags(&c->seeds_ev[k], cudaEventDisableTiming));
    }
    const int slot = c->seeds_slot;
    c->seeds_slot = (slot + 1) % SEED_SLOTS;
    CU(cudaEventSynchronize(c->seeds_ev[slot]));        // blocks only when SEED_SLOTS batches are already queued
    int *seeds_host = c->seeds_pinned + (size_t)slot * SEED_SLOT_INTS, *seeds_dev = c->seeds_dev + (size_t)slot * SEED_SLOT_INTS;
    memcpy(seeds_host, seeds, (size_t)n_passes * sizeof(int));
    CU(cudaMemcpyAsync(seeds_dev, seeds_host, (size_t)n_passes * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    fill_scene(c);
    const int n_pixels = c->width * c->height;
    if (!c->work_counter) CU(cudaMalloc(&c->work_counter, sizeof(unsigned int)));
    if (first_chunk) CU(cudaEventRecord(c->ev0, c->stream));
    const bool bvh = !(c->scene.world_bvh_empty && c->scene.actor_bvh_empty);
    const int mode = layout_mode(c);
    float *res = c->accum[c->accum_active];
    const float *res_prev = c->window_spp == 0 ? c->window_base : res;
    int kernel = c->params.kernel;
    if (kernel == 0) kernel = (mode != 0 && (!bvh || c->use_bvh2)) ? 4 : 1;   // layouts refused at commit: the reference's own arrays
    if (kernel == 1) {
        const int threads = 128, blocks = (n_pixels + threads - 1) / threads;
        if (mode == 0) k_render_mega<0><<<blocks, threads, 0, c->stream>>>(c->scene, seeds_dev, n_passes, c->window_spp, res, res_prev, n_pixels);
        else if (mode == 1) k_render_mega<1><<<blocks, threads, 0, c->stream>>>(c->scene, seeds_dev, n_passes, c->window_spp, res, res_prev, n_pixels);
        else k_render_mega<2><<<blocks, threads, 0, c->stream>>>(c->scene, seeds_dev, n_passes, c->window_spp, res, res_prev, n_pixels);
    } else {
        // persistent wavefront kernel (ccu_queue.cuh): one CTA per SM, pixels handed out through a counter
        if (mode == 0) return fail(CCU_ESTATE, "ccu_render_passes: kernel 4 needs the commit-time octree layouts (malformed octree?)");
        if (bvh && !c->use_bvh2) return fail(CCU_ESTATE, "ccu_render_passes: kernel 4 cannot hold this BVH (malformed, or deeper than the 64 entries of the reference's traversal stack, bvh.h:38); use kernel 0 or 1");
        CU(cudaMemsetAsync(c->work_counter, 0, sizeof(unsigned int), c->stream));
        QueueParams qp;
        qp.w.seeds = seeds_dev;
        qp.w.n_passes = n_passes;
        qp.w.start_spp = c->window_spp;
        qp.w.res = res;
        qp.w.res_prev = res_prev;
        qp.w.n_pixels = n_pixels;
        qp.w.next_pixel = c->work_counter;
        qp.yield_below = c->yield_below;
        qp.refill_min = c->q_refill_min;
        qp.march_bias = c->q_march_bias;
        qp.shade_min = c->q_shade_min;
        qp.sticky_min = c->q_sticky_min > 0 ? c->q_sticky_min : (bvh ? 16 : 33);   // measured: -3 % with BVH stages, +5 % without
        qp.leaf_min = c->q_leaf_min;
        qp.bvh_warps = c->q_bvh_warps;
        qp.march_warps = bvh ? 64 : c->q_march_warps;   // with BVHs the service warps are the ones that do not walk
        const int lay = air_layout_id(c);
        const int grid = c->sm_count, block = Q_WARPS * 32;
        // The sky table is staged in shared memory when it is small: shared memory and L1 are one array on this chip, and a
        // 128 x 128 table (64 KiB) taken from L1 costs more brick / atlas hits than the sky lookups gain (measured: +1.8 % per
        // pass on config 1).  CCU_SKY_SMEM_MAX (bytes, default 16 KiB = 64 x 64 texels) moves the threshold.
        const int sky_texels = c->sky_res * c->sky_res;
        const int base_smem = q_smem_bytes(bvh, lay == 0);
        const char *sky_env = getenv("CCU_SKY_SMEM_MAX");
        const int sky_max = sky_env ? atoi(sky_env) : 16384;
        const bool sky_smem = sky_texels * 4 <= sky_max && base_smem + sky_texels * 4 <= Q_SMEM_LIMIT;
        qp.sky_texels = sky_smem ? sky_texels : 0;
        qp.bvh_deep = nullptr;
#if CCU_BVH_PARK
        if (bvh) {
            // scratch for traversal-stack entries beyond the shared-memory part (bvh.h:38 allows 64 pending nodes)
            if (!c->bvh_deep) CU(cudaMalloc(&c->bvh_deep, (size_t)c->sm_count * Q_SLOTS * Q_DEEP * sizeof(int)));
            qp.bvh_deep = c->bvh_deep;
        }
#endif
        const int smem = base_smem + qp.sky_texels * 4;
        if (bvh) {
            if (lay == 0) k_render_queue<true, 0><<<grid, block, smem, c->stream>>>(c->scene, qp);
            else if (lay == 1) k_render_queue<true, 1><<<grid, block, smem, c->stream>>>(c->scene, qp);
            else k_render_queue<true, 2><<<grid, block, smem, c->stream>>>(c->scene, qp);
        } else {
            if (lay == 0) k_render_queue<false, 0><<<grid, block, smem, c->stream>>>(c->scene, qp);
            else if (lay == 1) k_render_queue<false, 1><<<grid, block, smem, c->stream>>>(c->scene, qp);
            else k_render_queue<false, 2><<<grid, block, smem, c->stream>>>(c->scene, qp);
        }
#ifdef CCU_Q_STATS
        {
            cudaStreamSynchronize(c->stream);
            unsigned long long st[32];
            cudaMemcpyFromSymbol(st, g_qstats, sizeof st);
            const char *names[4] = {"march", "block", "exit", "end"};
            fprintf(stderr, "[qstats] ");
            for (int i = 1; i < 4; i++) fprintf(stderr, "%s: %llu x %.1f lanes  ", names[i], st[2 * i], st[2 * i] ? (double)st[2 * i + 1] / st[2 * i] : 0.0);
            fprintf(stderr, "\n[qstats] march stages %llu, iterations %llu x %.1f lanes in flight, yields %llu, idle rounds %llu, pops %llu retries %llu\n", st[0], st[10],
                    st[10] ? (double)st[11] / st[10] : 0.0, st[13], st[12], st[14], st[15]);
            fprintf(stderr, "[qstats] bvh stages %llu, steps %llu x %.1f walking lanes, leaf turns %llu x %.1f lanes, shade %llu x %.1f lanes\n", st[16], st[18],
                    st[18] ? (double)st[19] / st[18] : 0.0, st[20], st[20] ? (double)st[21] / st[20] : 0.0, st[22], st[22] ? (double)st[23] / st[22] : 0.0);
            fprintf(stderr, "[qstats] march refills %llu, bvh refills %llu\n", st[24], st[25]);
            unsigned long long z[32] = {0};
            cudaMemcpyToSymbol(g_qstats, z, sizeof z);
        }
#endif
    }
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaEventRecord(c->seeds_ev[slot], c->stream));
    if (c->projector_type == -1) CU(cudaEventRecord(c->rays_used[c->rays_active], c->stream));
    c->timing_pending = true;
    c->window_spp += n_passes;
    c->window_base = res;
    return CCU_OK;
}

int ccu_render_passes_async(ccu_ctx *c, const int32_t *seeds, int32_t n_passes) {
    if (!c) return fail(CCU_EINVAL, "ccu_render_passes: null context");
    if (n_passes < 0 || (n_passes > 0 && !seeds)) return fail(CCU_EINVAL, "ccu_render_passes: bad seeds");
    std::lock_guard<std::mutex> lk(c->mu);
    int rc = render_ready(c, "ccu_render_passes");
    if (rc != CCU_OK) return rc;
    if (n_passes == 0) return CCU_OK;
    DeviceGuard g(c->device);
    // one launch covers at most 32768 passes (the wavefront kernel keeps a path's pass index in 16 bits); the reference's
    // windows are at most 1024 passes (OpenClPathTracingRenderer.java:158)
    const int32_t kChunk = SEED_SLOT_INTS;
    for (int32_t off = 0; off < n_passes; off += kChunk) {
        rc = render_passes_locked(c, seeds + off, std::min(kChunk, n_passes - off), off == 0);
        if (rc != CCU_OK) return rc;
    }
    return CCU_OK;
}

int ccu_render_sync(ccu_ctx *c) {
    if (!c) return fail(CCU_EINVAL, "ccu_render_sync: null context");
    int rc = wait_render_stream(c);       // the wait itself runs outside the context lock
    if (rc != CCU_OK) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    stop_timer(c);
    return CCU_OK;
}

int ccu_render_passes(ccu_ctx *c, const int32_t *seeds, int32_t n_passes) {
    int rc = ccu_render_passes_async(c, seeds, n_passes);
    if (rc != CCU_OK) return rc;
    return ccu_render_sync(c);
}

int ccu_render_read(ccu_ctx *c, float *mean_rgb, int32_t *window_spp) {
    if (!c || !mean_rgb) return fail(CCU_EINVAL, "ccu_render_read: null argument");
    int rc = wait_render_stream(c);
    if (rc != CCU_OK) return rc;
    std::unique_lock<std::mutex> lk(c->mu);
    if (!c->accum[0] || !c->target_live) return fail(CCU_ESTATE, "ccu_render_read: no render target");
    rc = join_merge_locked(c, lk);
    if (rc != CCU_OK) return rc;
    if (c->closed_spp > 0) return fail(CCU_ESTATE, "ccu_render_read: a closed window is waiting for ccu_render_window_merge (it owns the staging buffer)");
    DeviceGuard g(c->device);
    const size_t n = (size_t)c->width * c->height * 3;
    CU(cudaMemcpyAsync(c->pinned, c->accum[c->accum_active], n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    stop_timer(c);
    memcpy(mean_rgb, c->pinned, n * sizeof(float));
    if (window_spp) *window_spp = c->window_spp;
    return CCU_OK;
}

int ccu_render_merge_wait(ccu_ctx *c) {
    if (!c) return fail(CCU_EINVAL, "ccu_render_merge_wait: null context");
    std::unique_lock<std::mutex> lk(c->mu);
    return join_merge_locked(c, lk);
}

// closes the window under the lock: the passes that follow accumulate in the other buffer (the first of them still reads
// what this buffer holds, like the reference's single buffer at bufferSpp = 0, rayTracer.cl:111); the read-back starts
static int window_close_locked(ccu_ctx *c, std::unique_lock<std::mutex> &lk, int32_t *window_spp) {
    if (!c->accum[0] || !c->target_live) return fail(CCU_ESTATE, "no render target");
    int rc = join_merge_locked(c, lk);       // one merge at a time: they share the staging buffer
    if (rc != CCU_OK) return rc;
    if (c->closed_spp > 0) return fail(CCU_ESTATE, "the previous window was closed but not merged (ccu_render_window_merge)");
    const int pass_spp = c->window_spp;
    if (window_spp) *window_spp = pass_spp;
    if (pass_spp == 0) return CCU_OK;
    DeviceGuard g(c->device);
    CU(cudaEventRecord(c->window_ev, c->stream));
    const float *src = c->accum[c->accum_active];
    c->accum_active ^= 1;
    c->window_base = src;
    c->window_spp = 0;   // bufferSppReal = 0 (:170)
    c->closed_spp = pass_spp;
    CU(cudaStreamWaitEvent(c->copy_stream, c->window_ev, 0));
    return ccu_host::start_readback(c, src, 0, (size_t)c->width * c->height * 3, c->copy_stream);
}

int ccu_render_window_close(ccu_ctx *c, int32_t *window_spp) {
    if (!c) return fail(CCU_EINVAL, "ccu_render_window_close: null context");
    std::unique_lock<std::mutex> lk(c->mu);
    return window_close_locked(c, lk, window_spp);
}

int ccu_render_window_merge(ccu_ctx *c, double *sample_buffer, int32_t sample_spp) {
    if (!c || !sample_buffer) return fail(CCU_EINVAL, "ccu_render_window_merge: null argument");
    if (sample_spp < 0) return fail(CCU_EINVAL, "ccu_render_window_merge: negative spp");
    int pass_spp;
    size_t n;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        pass_spp = c->closed_spp;
        n = (size_t)c->width * c->height * 3;
    }
    if (pass_spp == 0) return CCU_OK;
    const double sinv = 1.0 / (double)(sample_spp + pass_spp);
    // the wait for the copies and the merge itself run outside the context lock: the next window renders meanwhile
    int rc = ccu_host::merge_readback(c, 0, n, sample_buffer, (double)sample_spp, (double)pass_spp, sinv, 16);
    std::lock_guard<std::mutex> lk(c->mu);
    c->closed_spp = 0;
    return rc;
}

int ccu_render_merge_async(ccu_ctx *c, double *sample_buffer, int32_t sample_spp, int32_t *merged_spp) {
    if (!c || !sample_buffer) return fail(CCU_EINVAL, "ccu_render_merge: null argument");
    if (sample_spp < 0) return fail(CCU_EINVAL, "ccu_render_merge: negative spp");
    std::unique_lock<std::mutex> lk(c->mu);
    int32_t pass_spp = 0;
    int rc = window_close_locked(c, lk, &pass_spp);
    if (merged_spp) *merged_spp = pass_spp;
    if (rc != CCU_OK || pass_spp == 0) return rc;
    c->merge.active = true;
    c->merge.status = CCU_OK;
    c->merge.worker = std::thread([=] {
        const int st = ccu_render_window_merge(c, sample_buffer, sample_spp);
        if (st != CCU_OK) { c->merge.status = st; c->merge.error = ccu_host::last_error(); }
    });
    return CCU_OK;
}

int ccu_render_merge(ccu_ctx *c, double *sample_buffer, int32_t sample_spp, int32_t *merged_spp) {
    if (!c || !sample_buffer) return fail(CCU_EINVAL, "ccu_render_merge: null argument");
    if (sample_spp < 0) return fail(CCU_EINVAL, "ccu_render_merge: negative spp");
    int32_t pass_spp = 0;
    {
        std::unique_lock<std::mutex> lk(c->mu);
        int rc = window_close_locked(c, lk, &pass_spp);
        if (merged_spp) *merged_spp = pass_spp;
        if (rc != CCU_OK) return rc;
    }
    int rc = ccu_render_window_merge(c, sample_buffer, sample_spp);
    if (rc != CCU_OK) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    stop_timer(c);
    return CCU_OK;
}

int ccu_render_reset_window(ccu_ctx *c) {
    if (!c) return fail(CCU_EINVAL, "ccu_render_reset_window: null context");
    std::lock_guard<std::mutex> lk(c->mu);
    c->window_spp = 0;
    return CCU_OK;
}

int ccu_render_set_window_spp(ccu_ctx *c, int32_t window_spp) {
    if (!c) return fail(CCU_EINVAL, "ccu_render_set_window_spp: null context");
    if (window_spp < 0) return fail(CCU_EINVAL, "ccu_render_set_window_spp: negative pass count");
    std::lock_guard<std::mutex> lk(c->mu);
    c->window_spp = window_spp;
    return CCU_OK;
}

int ccu_render_device_buffer(ccu_ctx *c, void **device_ptr, int64_t *n_floats) {
    if (!c || !device_ptr) return fail(CCU_EINVAL, "ccu_render_device_buffer: null argument");
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->accum[0] || !c->target_live) return fail(CCU_ESTATE, "ccu_render_device_buffer: no render target");
    *device_ptr = c->accum[c->accum_active];
    if (n_floats) *n_floats = (int64_t)c->width * c->height * 3;
    return CCU_OK;
}

int ccu_render_scale(ccu_ctx *c, float factor) {
    if (!c) return fail(CCU_EINVAL, "ccu_render_scale: null context");
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->accum[0] || !c->target_live) return fail(CCU_ESTATE, "ccu_render_scale: no render target");
    DeviceGuard g(c->device);
    size_t n = (size_t)c->width * c->height * 3;
    k_scale<<<c->sm_count * 4, 256, 0, c->stream>>>(c->accum[c->accum_active], factor, n);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return CCU_OK;
}

int ccu_stream_handle(ccu_ctx *c, void **cuda_stream) {
    if (!c || !cuda_stream) return fail(CCU_EINVAL, "ccu_stream_handle: null argument");
    std::lock_guard<std::mutex> lk(c->mu);
    *cuda_stream = (void *)c->stream;
    return CCU_OK;
}

int ccu_first_hit(ccu_ctx *c, int32_t seed, int32_t *block, int32_t *face, int32_t *node, int32_t *kind, float *t, float *normal,
                  float *color) {
    if (!c) return fail(CCU_EINVAL, "ccu_first_hit: null context");
    std::lock_guard<std::mutex> lk(c->mu);
    int rc = render_ready(c, "ccu_first_hit");
    if (rc != CCU_OK) return rc;
    DeviceGuard g(c->device);
    size_t n = (size_t)c->width * c->height;
    // one scratch allocation (kept for the next call): 4 int planes, colour(4), t, normal(3) = 12 words per pixel
    if (c->fh_scratch_pixels < n) {
        cudaStreamSynchronize(c->stream);
        if (c->fh_scratch) cudaFree(c->fh_scratch);
        c->fh_scratch = nullptr;
        c->fh_scratch_pixels = 0;
        CU(cudaMalloc(&c->fh_scratch, n * 12 * sizeof(int)));
        c->fh_scratch_pixels = n;
    }
    int *scratch = c->fh_scratch;
    // planes nobody asked for are not written
    int *d_block = block ? scratch : nullptr, *d_face = face ? scratch + n : nullptr, *d_node = node ? scratch + 2 * n : nullptr,
        *d_kind = kind ? scratch + 3 * n : nullptr;
    float *d_color = color ? reinterpret_cast<float *>(scratch + 4 * n) : nullptr;          // 16-byte aligned: n*16 bytes offset
    float *d_t = t ? reinterpret_cast<float *>(scratch + 8 * n) : nullptr;
    float *d_normal = normal ? reinterpret_cast<float *>(scratch + 9 * n) : nullptr;
    fill_scene(c);
    const int mode = layout_mode(c);
    const unsigned blocks = (unsigned)((n + CCU_FH_THREADS - 1) / CCU_FH_THREADS);
    cudaEventRecord(c->ev0, c->stream);
    const bool fh_bvh = !(c->scene.world_bvh_empty && c->scene.actor_bvh_empty);
    const FirstHitOut fho = {d_block, d_face, d_node, d_kind, d_t, d_normal, d_color};
#define CCU_FH(M, B) k_first_hit<M, B><<<blocks, CCU_FH_THREADS, 0, c->stream>>>(c->scene, seed, (int)n, fho)
    if (fh_bvh) { if (mode == 0) CCU_FH(0, true); else if (mode == 1) CCU_FH(1, true); else CCU_FH(2, true); }
    else { if (mode == 0) CCU_FH(0, false); else if (mode == 1) CCU_FH(1, false); else CCU_FH(2, false); }
#undef CCU_FH
    c->launches++;
    cudaEventRecord(c->ev1, c->stream);
    if (c->projector_type == -1) cudaEventRecord(c->rays_used[c->rays_active], c->stream);
    c->timing_pending = true;
    cudaError_t e = cudaGetLastError();
    auto back = [&](void *dst, const void *src, size_t bytes) {
        if (dst && e == cudaSuccess) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream);
    };
    back(block, d_block, n * 4); back(face, d_face, n * 4); back(node, d_node, n * 4); back(kind, d_kind, n * 4);
    back(t, d_t, n * 4); back(normal, d_normal, n * 12); back(color, d_color, n * 16);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return fail(CCU_ECUDA, "ccu_first_hit: %s", cudaGetErrorString(e));
    stop_timer(c);
    return CCU_OK;
}

int ccu_preview(ccu_ctx *c, int32_t *argb) {
    if (!c || !argb) return fail(CCU_EINVAL, "ccu_preview: null argument");
    std::lock_guard<std::mutex> lk(c->mu);
    int rc = render_ready(c, "ccu_preview");
    if (rc != CCU_OK) return rc;
    DeviceGuard g(c->device);
    size_t n = (size_t)c->width * c->height;
    int *d = nullptr;
    CU(cudaMalloc(&d, n * sizeof(int)));
    fill_scene(c);
    const int mode = layout_mode(c);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    cudaEventRecord(c->ev0, c->stream);
    if (mode == 0) k_preview<0><<<blocks, 256, 0, c->stream>>>(c->scene, (int)n, d);
    else if (mode == 1) k_preview<1><<<blocks, 256, 0, c->stream>>>(c->scene, (int)n, d);
    else k_preview<2><<<blocks, 256, 0, c->stream>>>(c->scene, (int)n, d);
    c->launches++;
    cudaEventRecord(c->ev1, c->stream);
    if (c->projector_type == -1) cudaEventRecord(c->rays_used[c->rays_active], c->stream);
    c->timing_pending = true;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(argb, d, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(CCU_ECUDA, "ccu_preview: %s", cudaGetErrorString(e));
    stop_timer(c);
    return CCU_OK;
}

int ccu_last_kernel_ms(ccu_ctx *c, float *ms) {
    if (!c || !ms) return fail(CCU_EINVAL, "ccu_last_kernel_ms: null argument");
    int rc = wait_render_stream(c);
    if (rc != CCU_OK) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    stop_timer(c);
    *ms = c->last_ms;
    return CCU_OK;
}

int ccu_launch_count(ccu_ctx *c, int64_t *launches) {
    if (!c || !launches) return fail(CCU_EINVAL, "ccu_launch_count: null argument");
    std::lock_guard<std::mutex> lk(c->mu);
    *launches = c->launches;
    return CCU_OK;
}

int ccu_scene_device_bytes(ccu_ctx *c, int64_t *bytes) {
    if (!c || !bytes) return fail(CCU_EINVAL, "ccu_scene_device_bytes: null argument");
    std::lock_guard<std::mutex> lk(c->mu);
    *bytes = (int64_t)(c->tree.bytes() + c->top.bytes() + c->wide.bytes() + c->air_top.bytes() + c->air_wide.bytes() + c->air_bricks.bytes() +
                       c->world_rec.bytes() + c->actor_rec.bytes() + c->tris2.bytes() + c->block_rec.bytes() + c->mat_rec.bytes() + c->quad_rec.bytes() + c->aabb_rec.bytes() + c->block_palette.bytes() +
                       c->quad_models.bytes() + c->aabb_models.bytes() + c->mat_palette.bytes() + c->trigs.bytes() + c->world_bvh.bytes() +
                       c->actor_bvh.bytes() + c->atlas.bytes() + c->sky.bytes());
    return CCU_OK;
}

int ccu_scene_commit_ms(ccu_ctx *c, double *ms) {
    if (!c || !ms) return fail(CCU_EINVAL, "ccu_scene_commit_ms: null argument");
    std::lock_guard<std::mutex> lk(c->mu);
    *ms = c->commit_ms;
    return CCU_OK;
}

int ccu_tonemap(ccu_ctx *c, int32_t width, int32_t height, float exposure, const double *input, int32_t type, int32_t *argb) {
    if (!c || !input || !argb) return fail(CCU_EINVAL, "ccu_tonemap: null argument");
    if (width <= 0 || height <= 0 || (int64_t)width * height > (1ll << 30)) return fail(CCU_EINVAL, "ccu_tonemap: bad canvas %dx%d", width, height);
    if (type < 0 || type > 3) return fail(CCU_EINVAL, "ccu_tonemap: unknown filter %d (0 GAMMA, 1 TONEMAP1, 2 ACES, 3 HABLE)", type);
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    const size_t n = (size_t)width * height;
    double *d_in = nullptr;
    uint32_t *d_out = nullptr;
    CU(cudaMalloc(&d_in, n * 3 * sizeof(double)));
    cudaError_t e = cudaMalloc(&d_out, n * sizeof(uint32_t));
    // the reference uploads the whole double buffer and reads the ARGB image back on every call (GpuPostProcessingFilter.java:43-61)
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, input, n * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        cudaEventRecord(c->ev0, c->stream);
        k_tonemap<<<c->sm_count * 8, 256, 0, c->stream>>>((int)n, exposure, d_in, type, d_out);
        cudaEventRecord(c->ev1, c->stream);
        c->timing_pending = true;
        c->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(argb, d_out, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_in);
    if (d_out) cudaFree(d_out);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? CCU_ENOMEM : CCU_ECUDA, "ccu_tonemap: %s", cudaGetErrorString(e));
    stop_timer(c);
    return CCU_OK;
}

int ccu_bench_gather(ccu_ctx *c, int64_t array_bytes, int32_t dependent, float *gbytes_per_s, float *ns_per_load) {
    if (!c || array_bytes < 4096) return fail(CCU_EINVAL, "ccu_bench_gather: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    size_t sectors = 1;
    while (sectors * 2 * 32 <= (size_t)array_bytes) sectors *= 2;   // power of two number of 32-byte sectors
    uint4 *a = nullptr;
    uint32_t *sink = nullptr;
    CU(cudaMalloc(&a, sectors * 32));
    cudaError_t e = cudaMalloc(&sink, 4);
    if (e != cudaSuccess) { cudaFree(a); return fail(CCU_ENOMEM, "ccu_bench_gather: %s", cudaGetErrorString(e)); }
    k_gather_fill<<<c->sm_count * 8, 256, 0, c->stream>>>(a, sectors, 12345u);
    const int loads = dependent ? 2048 : 4096;
    const int blocks = dependent ? c->sm_count : c->sm_count * 8;
    float best = 1e30f;
    for (int it = 0; it < 4; it++) {   // first iteration warms the caches
        cudaEventRecord(c->ev0, c->stream);
        k_gather<<<blocks, dependent ? 32 : 256, 0, c->stream>>>(a, (uint32_t)(sectors - 1), loads, dependent, (uint32_t)it, sink);
        cudaEventRecord(c->ev1, c->stream);
        e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        if (it > 0) best = std::min(best, ms);
    }
    c->launches += 5;
    cudaFree(a);
    cudaFree(sink);
    if (e != cudaSuccess) return fail(CCU_ECUDA, "ccu_bench_gather: %s", cudaGetErrorString(e));
    const double n_loads = (double)blocks * (dependent ? 1.0 : 256.0) * loads;
    if (gbytes_per_s) *gbytes_per_s = (float)(n_loads * 32.0 / (best * 1e-3) / 1e9);
    if (ns_per_load) *ns_per_load = (float)(best * 1e6 / loads);   // per thread: meaningful for the dependent chain
    return CCU_OK;
}

// Host-only check of the commit-time traversal layouts (no CUDA call): for every voxel of `xyz` (count x 3 ints) returns what
// the value-carrying layout (find_leaf_wide) and the march layout (lean_probe) answer, so that CPU tests can compare both
// with the reference's root descent (octree.h:81-88).
int ccu_debug_bvh_layout(const int32_t *bvh, int64_t n_bvh, const int32_t *trigs, int64_t n_trigs, int32_t *rec, int64_t rec_cap,
                         int32_t *tris, int64_t tris_cap, int64_t *rec_words, int64_t *tris_words, int32_t *root, int32_t *ok) {
    if (!bvh || n_bvh < 0 || n_trigs < 0 || (n_trigs > 0 && !trigs) || !rec_words || !tris_words || !root || !ok)
        return fail(CCU_EINVAL, "ccu_debug_bvh_layout: bad argument");
    const std::vector<int> nodes(bvh, bvh + n_bvh), tp(trigs, trigs + n_trigs);
    BvhLayout b;
    TriRepack tr;
    std::unordered_map<int, int> leaf_map;
    b.root = bvh_is_empty(nodes) ? 0 : bvh_ref(nodes, tp, 0, 0, b, tr, leaf_map);
    if (!b.ok) { b.rec.clear(); tr.tris.clear(); }
    tr.tris.insert(tr.tris.end(), 8, 0);        // as ccu_scene_commit does
    *ok = b.ok ? 1 : 0;
    *root = b.root;
    *rec_words = (int64_t)b.rec.size();
    *tris_words = (int64_t)tr.tris.size();
    if (rec) {
        if (rec_cap < (int64_t)b.rec.size()) return fail(CCU_EINVAL, "ccu_debug_bvh_layout: rec buffer too small");
        std::copy(b.rec.begin(), b.rec.end(), rec);
    }
    if (tris) {
        if (tris_cap < (int64_t)tr.tris.size()) return fail(CCU_EINVAL, "ccu_debug_bvh_layout: tris buffer too small");
        std::copy(tr.tris.begin(), tr.tris.end(), tris);
    }
    return CCU_OK;
}

int ccu_debug_layout_lookup(const int32_t *tree, int64_t n, int32_t depth, const int32_t *xyz, int64_t count, int32_t *wide_value,
                            int32_t *wide_level, int32_t *air_solid, int32_t *air_level) {
    if (!tree || n < 1 || depth < 0 || depth > 30 || (count > 0 && !xyz)) return fail(CCU_EINVAL, "ccu_debug_layout_lookup: bad argument");
    WideLayout wl = build_wide_layout(tree, (size_t)n, depth);
    AirLayout al = build_air_layout(tree, (size_t)n, depth);
    if (!al.ok) return fail(CCU_ESTATE, "air layout could not be built");
    const int cl = wl.cell_level, tl = wl.top_log2;
    for (int64_t i = 0; i < count; i++) {
        const int bx = xyz[3 * i], by = xyz[3 * i + 1], bz = xyz[3 * i + 2];
        if (((bx | by | bz) >> depth) != 0) return fail(CCU_EINVAL, "voxel %lld outside the octree cube", (long long)i);
        if (wl.ok) {
            const size_t cell = ((((size_t)(bx >> cl) << tl) + (size_t)(by >> cl)) << tl) + (size_t)(bz >> cl);
            unsigned e = wl.top[cell];
            int lvl = cl;
            while (!(e & CCU_WIDE_LEAF)) {
                lvl -= 2;
                e = wl.wide[(size_t)e * 64 + ((((bx >> lvl) & 3) << 4) | (((by >> lvl) & 3) << 2) | ((bz >> lvl) & 3))];
            }
            const unsigned v = e & CCU_WIDE_ANY;
            if (wide_value) wide_value[i] = v == CCU_WIDE_ANY ? CCU_ANY_TYPE : (int)v;
            if (wide_level) wide_level[i] = (int)((e >> 26) & 31);
        } else {
            if (wide_value) wide_value[i] = -1;
            if (wide_level) wide_level[i] = -1;
        }
        // the lookup of lean_probe (ccu_march.cuh)
        int lvl = al.cell_level;
        const int atl = al.top_log2;
        unsigned e = al.top[((((size_t)(bx >> lvl) << atl) + (size_t)(by >> lvl)) << atl) + (size_t)(bz >> lvl)];
        while (lvl > 4 && !(e & CCU_WIDE_LEAF)) {
            lvl -= 2;
            e = al.wide[(size_t)e * 64 + ((((bx >> lvl) & 3) << 4) | (((by >> lvl) & 3) << 2) | ((bz >> lvl) & 3))];
        }
        int level;
        if (e & CCU_WIDE_LEAF) {
            level = ((int)(e << 1)) >> 27;
        } else {
            const unsigned wi = (unsigned)(((bx & 12) << 4) | ((by & 12) << 2) | (bz & 12) | (bx & 3));
            const unsigned w = al.bricks[(size_t)e * 256 + wi];
            const int code = (int)((w >> ((((by & 3) << 2) | (bz & 3)) * 2)) & 3u);
            level = w >= CCU_BRICK_UNIFORM ? (int)(w & 31u) : code - 1;
        }
        if (air_solid) air_solid[i] = level < 0 ? 1 : 0;
        if (air_level) air_level[i] = level;
    }
    return CCU_OK;
}

}  // extern "C"
