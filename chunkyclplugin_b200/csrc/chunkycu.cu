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

#include "../../include/chunkycu.h"
#include "ccu_device.cuh"
#include "ccu_wavefront.cuh"
#include "ccu_pool.cuh"
#include "ccu_queue.cuh"
#include "ccu_tonemap.cuh"

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
template <bool WIDE>
__global__ void __launch_bounds__(128) k_render_mega(const __grid_constant__ DScene s, const int *__restrict__ seeds, int n_passes,
                                                     int start_spp, float *__restrict__ res, int n_pixels) {
    for (int gid = blockIdx.x * blockDim.x + threadIdx.x; gid < n_pixels; gid += gridDim.x * blockDim.x) {
        float *px = res + (size_t)gid * 3;
        float3 buf = f3(px[0], px[1], px[2]);
        for (int pass = 0; pass < n_passes; pass++) {
            float3 col = sample_pixel<WIDE>(s, gid, __ldg(seeds + pass));
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

// WIDE: march on the commit-time layout; the treeData index of the hit leaf (reference numbering, ClSceneLoader.java:56-59)
// is then found by one root descent for the hit voxel only.
template <bool WIDE>
__global__ void __launch_bounds__(128) k_first_hit(const __grid_constant__ DScene s, int seed, int n_pixels, int *block, int *face, int *node,
                                                   int *kind, float *t, float *normal, float *color) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_pixels) return;
    uint32_t rng = (uint32_t)seed + (uint32_t)gid;
    rng_next(rng);
    float3 o, d;
    camera_ray<false>(s, gid, rng, o, d);
    Record rec;
    rec.distance = inff_(); rec.material = 0; rec.surf.normal = f3(0, 0, 0); rec.point = f3(0, 0, 0);
    rec.surf.color = make_float4(0, 0, 0, 0); rec.surf.emittance = 0;
    HitInfo hi = {-1, 0, 0, 0, 0};
    bool hit = closest_intersect<WIDE>(s, o, d, rec, hi);
    if (WIDE && hit && hi.kind == 1) {
        int level;
        find_leaf(s, hi.bx, hi.by, hi.bz, level, hi.node);
    }
    if (block) block[gid] = hit ? rec.material : 0;
    if (face) face[gid] = hit ? face_of(rec.surf.normal) : 6;
    if (node) node[gid] = hit ? hi.node : -1;
    if (kind) kind[gid] = hit ? hi.kind : 0;
    if (t) t[gid] = hit ? rec.distance : inff_();
    if (normal) {
        normal[gid * 3 + 0] = hit ? rec.surf.normal.x : 0.0f;
        normal[gid * 3 + 1] = hit ? rec.surf.normal.y : 0.0f;
        normal[gid * 3 + 2] = hit ? rec.surf.normal.z : 0.0f;
    }
    if (color) {
        float4 c = hit ? rec.surf.color : make_float4(0, 0, 0, 0);
        reinterpret_cast<float4 *>(color)[gid] = c;
    }
}

// rayTracer.cl:141-216
template <bool WIDE>
__global__ void __launch_bounds__(128) k_preview(const __grid_constant__ DScene s, int n_pixels, int *res) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_pixels) return;
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
    if (WIDE ? closest_intersect<true>(s, o, d, rec, hi) : closest_intersect<false>(s, o, d, rec, hi)) {
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
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                                     \
    do {                                                                                                             \
        cudaError_t e_ = (call);                                                                                     \
        if (e_ != cudaSuccess) return fail(e_ == cudaErrorMemoryAllocation ? CCU_ENOMEM : CCU_ECUDA, "%s: %s (%s:%d)", #call, \
                                           cudaGetErrorString(e_), __FILE__, __LINE__);                              \
    } while (0)

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t upload(const T *host, size_t count, cudaStream_t st) {
        release();
        n = count;
        size_t alloc = std::max<size_t>(count, 4);   // zero-length arrays become a zero word (ClIntBuffer.java:15-18)
        cudaError_t e = cudaMalloc(&p, alloc * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; return e; }
        e = cudaMemsetAsync(p, 0, alloc * sizeof(T), st);
        if (e != cudaSuccess) return e;
        if (count) e = cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
        return cudaStreamSynchronize(st);   // host memory is not retained past the call
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    size_t bytes() const { return p ? std::max<size_t>(n, 4) * sizeof(T) : 0; }
};

struct ccu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t chunk_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // read-back chunks of ccu_render_merge
    std::mutex mu;
    int sm_count = 0;

    // scene
    DevBuf<int> tree, block_palette, quad_models, aabb_models, mat_palette, trigs, world_bvh, actor_bvh, sun_words;
    std::vector<int> world_head, actor_head;    // first node of each BVH for the emptiness probe
    std::vector<int> world_host, actor_host, trigs_host;   // kept for the commit-time BVH layout
    DevBuf<int> world_rec, actor_rec, tris2;
    int world_root = 0, actor_root = 0, use_bvh2 = 0;
    DevBuf<uchar4> atlas, sky;
    std::vector<int> tree_host;              // kept for the commit-time traversal layout
    DevBuf<unsigned> top, wide;
    DevBuf<unsigned> air_top, air_wide, air_bits;   // march-loop layout (ccu_queue.cuh)
    int use_air = 0;
    int cell_level = 0, top_log2 = 0, use_wide = 0;
    int atlas_w = 0, atlas_h = 0, atlas_layers = 0;
    int depth = 0, sky_res = 0;
    float sky_intensity = 0;
    int sun_host[6] = {0, 0, 0, 0, 0, 0};
    bool have_sun = false;
    bool committed = false;
    DevBuf<float> sun_basis;
    float *unorm = nullptr;

    // camera
    int projector_type = 0;
    float cam[15] = {0};
    DevBuf<float> rays;
    bool have_camera = false;

    // render target
    int width = 0, height = 0;
    float *accum = nullptr;      // running mean float[3*W*H]
    float *pinned = nullptr;     // host staging float[3*W*H]
    int *seeds_dev = nullptr;
    unsigned int *work_counter = nullptr;
    int wait_lanes = 28;
    int refill_min = 4;
    int exit_idle = 8;
    int yield_below = 20;
    int q_refill_min = 8;
    int q_march_bias = 4;
    int q_leaf_min = 12;
    int q_bvh_warps = 23;
    int q_march_warps = 22;
    int blocks_per_sm = CCU_MIN_BLOCKS;
    int seeds_cap = 0;
    int window_spp = 0;
    bool target_live = false;    // between ccu_render_begin and ccu_render_end
    ccu_render_params params = {256, 5, 13.0f, 0};

    float last_ms = 0;
    bool timing_pending = false;
    int64_t launches = 0;

    DScene scene{};
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

bool bvh_is_empty(const std::vector<int> &head) {   // bvh.h:23-32
    if (head.size() < 7 || head[0] != 0) return head.size() < 7;
    for (int i = 1; i < 7; i++) {
        float f;
        memcpy(&f, &head[i], 4);
        if (!std::isnan(f)) return false;
    }
    return true;
}

int fill_scene(ccu_ctx *c) {
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
    s.air_bits = c->air_bits.p;
    s.world_rec = reinterpret_cast<const int4 *>(c->world_rec.p);
    s.actor_rec = reinterpret_cast<const int4 *>(c->actor_rec.p);
    s.tris2 = reinterpret_cast<const int4 *>(c->tris2.p);
    s.world_root = c->world_root;
    s.actor_root = c->actor_root;
    s.block_palette = c->block_palette.p;
    s.block_palette_len = (int)c->block_palette.n;
    s.quad_models = c->quad_models.p;
    s.aabb_models = c->aabb_models.p;
    s.mat_palette = c->mat_palette.p;
    s.world_bvh = c->world_bvh.p;
    s.actor_bvh = c->actor_bvh.p;
    s.trigs = c->trigs.p;
    s.world_bvh_empty = bvh_is_empty(c->world_head);
    s.actor_bvh_empty = bvh_is_empty(c->actor_head);
    s.atlas = c->atlas.p;
    s.atlas_w = c->atlas_w; s.atlas_h = c->atlas_h; s.atlas_layers = c->atlas_layers;
    s.atlas_tiles_x = (c->atlas_w + 15) / 16; s.atlas_tiles_y = (c->atlas_h + 15) / 16;
    s.sky = c->sky.p;
    s.sky_res = c->sky_res;
    s.sky_intensity = c->sky_intensity;
    s.unorm = c->unorm;
    s.sun_flags = c->sun_host[0]; s.sun_tex_size = c->sun_host[1]; s.sun_tex = c->sun_host[2];
    memcpy(&s.sun_intensity, &c->sun_host[3], 4);
    s.projector_type = c->projector_type;
    memcpy(s.cam, c->cam, sizeof s.cam);
    s.rays = c->rays.p;
    s.width = c->width; s.height = c->height;
    if (c->height > 0) {
        s.half_width = (float)(c->width / (2.0 * c->height));   // rayTracer.cl:66
        s.inv_height = (float)(1.0 / c->height);                // rayTracer.cl:67
    }
    s.draw_depth = c->params.draw_depth; s.max_depth = c->params.max_depth; s.emitter_scale = c->params.emitter_scale;
    return CCU_OK;
}

int upload_words(ccu_ctx *c, DevBuf<int> &dst, const int32_t *words, int64_t n, const char *what) {
    if (!c) return fail(CCU_EINVAL, "%s: null context", what);
    if (n < 0 || (n > 0 && !words)) return fail(CCU_EINVAL, "%s: bad array (n=%lld)", what, (long long)n);
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    c->committed = false;
    CU(dst.upload(words, (size_t)n, c->stream));
    return CCU_OK;
}

void stop_timer(ccu_ctx *c) {
    if (c->timing_pending) {
        cudaEventSynchronize(c->ev1);
        cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1);
        c->timing_pending = false;
    }
}

}  // namespace


// ------------------------------------------------------------------------------------------------------
// commit-time traversal layout (see DScene::top / DScene::wide)
// ------------------------------------------------------------------------------------------------------
namespace {
struct WideLayout {
    std::vector<unsigned> top, wide;
    int cell_level = 0, top_log2 = 0;
    bool ok = true;
};

inline unsigned wide_leaf(int word, int level, bool &ok) {
    long long v = -(long long)word;
    unsigned val;
    if (v == 0x7FFFFFFELL) val = CCU_WIDE_ANY;
    else if (v >= 0 && v < (long long)CCU_WIDE_ANY) val = (unsigned)v;
    else { ok = false; val = 0; }
    return CCU_WIDE_LEAF | ((unsigned)level << 26) | val;
}

unsigned wide_node(const int *tree, size_t n, int word, int lvl, WideLayout &b) {
    if (lvl < 2 || (size_t)word + 7 >= n) { b.ok = false; return wide_leaf(0, 0, b.ok); }
    const size_t idx = b.wide.size() / 64;
    if (idx >= 0x7FFFFFFFu) { b.ok = false; return wide_leaf(0, 0, b.ok); }
    b.wide.resize(b.wide.size() + 64);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 4; k++) {
                unsigned e;
                int c1 = tree[(size_t)word + ((((i >> 1) & 1) << 2) | (((j >> 1) & 1) << 1) | ((k >> 1) & 1))];
                if (c1 <= 0) {
                    e = wide_leaf(c1, lvl - 1, b.ok);
                } else if ((size_t)c1 + 7 >= n) {
                    b.ok = false;
                    e = wide_leaf(0, 0, b.ok);
                } else {
                    int c2 = tree[(size_t)c1 + (((i & 1) << 2) | ((j & 1) << 1) | (k & 1))];
                    e = c2 <= 0 ? wide_leaf(c2, lvl - 2, b.ok) : wide_node(tree, n, c2, lvl - 2, b);
                }
                b.wide[idx * 64 + ((i << 4) | (j << 2) | k)] = e;
            }
    return (unsigned)idx;
}

// "Air layout" for the march loop (ccu_queue.cuh lean_probe): the same top table / 64-ary nodes, but an entry only says
// air or not (bit 0 = 1: not air) plus the leaf level, and the finest nodes (4^3 voxels) shrink from 64 words to a
// 128-bit map of 2-bit codes (0 = not air, 1 = air leaf of level 0, 2 = air leaf of level 1).  The march loop needs
// nothing else (octree.h:89-106: an air leaf is left through its cube, anything else goes to the block test), and the
// structure is ~10x smaller than the value-carrying layout, which is what keeps it in L1.
struct AirLayout {
    std::vector<unsigned> top, wide, bits;
    bool ok = true;
};

inline unsigned air_leaf(int word, int level) { return CCU_WIDE_LEAF | ((unsigned)level << 26) | (word == 0 ? 0u : 1u); }

unsigned air_node(const int *tree, size_t n, int word, int lvl, AirLayout &b) {
    if (lvl < 2 || (size_t)word + 7 >= n) { b.ok = false; return air_leaf(1, 0); }
    if (lvl == 2) {
        const size_t idx = b.bits.size() / 4;
        if (idx >= 0x7FFFFFFFu) { b.ok = false; return air_leaf(1, 0); }
        b.bits.resize(b.bits.size() + 4, 0u);
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++)
                for (int k = 0; k < 4; k++) {
                    unsigned code;
                    int c1 = tree[(size_t)word + ((((i >> 1) & 1) << 2) | (((j >> 1) & 1) << 1) | ((k >> 1) & 1))];
                    if (c1 <= 0) {
                        code = c1 == 0 ? 2u : 0u;
                    } else if ((size_t)c1 + 7 >= n) {
                        b.ok = false;
                        code = 0;
                    } else {
                        int c2 = tree[(size_t)c1 + (((i & 1) << 2) | ((j & 1) << 1) | (k & 1))];
                        if (c2 > 0) b.ok = false;          // deeper than the declared depth
                        code = c2 == 0 ? 1u : 0u;
                    }
                    const int v = (i << 4) | (j << 2) | k;
                    b.bits[idx * 4 + (v >> 4)] |= code << ((v & 15) * 2);
                }
        return (unsigned)idx;
    }
    const size_t idx = b.wide.size() / 64;
    if (idx >= 0x7FFFFFFFu) { b.ok = false; return air_leaf(1, 0); }
    b.wide.resize(b.wide.size() + 64);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 4; k++) {
                unsigned e;
                int c1 = tree[(size_t)word + ((((i >> 1) & 1) << 2) | (((j >> 1) & 1) << 1) | ((k >> 1) & 1))];
                if (c1 <= 0) {
                    e = air_leaf(c1, lvl - 1);
                } else if ((size_t)c1 + 7 >= n) {
                    b.ok = false;
                    e = air_leaf(1, 0);
                } else {
                    int c2 = tree[(size_t)c1 + (((i & 1) << 2) | ((j & 1) << 1) | (k & 1))];
                    e = c2 <= 0 ? air_leaf(c2, lvl - 2) : air_node(tree, n, c2, lvl - 2, b);
                }
                b.wide[idx * 64 + ((i << 4) | (j << 2) | k)] = e;
            }
    return (unsigned)idx;
}

// same cell level / table shape as the wide layout
AirLayout build_air_layout(const int *tree, size_t n, int depth, int cl) {
    AirLayout b;
    const int dim = 1 << (depth - cl);
    b.top.assign((size_t)dim * dim * dim, 0u);
    for (int x = 0; x < dim && b.ok; x++)
        for (int y = 0; y < dim; y++)
            for (int z = 0; z < dim; z++) {
                int level = depth;
                int word = tree[0];
                while (word > 0 && level > cl) {
                    level--;
                    int sh = level - cl;
                    size_t at = (size_t)word + ((((x >> sh) & 1) << 2) | (((y >> sh) & 1) << 1) | ((z >> sh) & 1));
                    if (at >= n) { b.ok = false; word = 0; break; }
                    word = tree[at];
                }
                if (word > 0 && cl < 2) { b.ok = false; word = -1; }
                b.top[((size_t)x * dim + y) * dim + z] = word <= 0 ? air_leaf(word, level) : air_node(tree, n, word, cl, b);
            }
    if (b.wide.empty()) b.wide.assign(64, air_leaf(1, 0));
    if (b.bits.empty()) b.bits.assign(4, 0u);
    return b;
}

// Traversal layout of a packed BVH (PackedBvhNode.java:22-31: 7 ints per node, first child at node + 7, second child at
// node[0]) for the BVH stage of ccu_queue.cuh: one 64-byte record per inner node holding BOTH children's boxes
// (bvh.h:73-91 fetches exactly those at every inner node) and a reference per child, and 16-byte aligned triangle
// blocks.  ref >= 0: record index; ref < 0: leaf, -(1 + offset of its block in `tris`, in units of 4 words).
struct BvhLayout {
    std::vector<int> rec;
    int root = 0;
    bool ok = true;
};
struct TriRepack {
    std::vector<int> tris;                       // per leaf: {count, 0, 0, 0} + count x 20 words (PackedTriangle.java:46-78)
    int add(const std::vector<int> &trigs, int prim, bool &ok) {
        if (prim < 0 || (size_t)prim >= trigs.size()) { ok = false; return 0; }
        const int count = trigs[(size_t)prim];
        if (count < 0 || (size_t)prim + 1 + (size_t)count * 20 > trigs.size()) { ok = false; return 0; }
        const int off = (int)(tris.size() / 4);
        tris.push_back(count); tris.push_back(0); tris.push_back(0); tris.push_back(0);
        tris.insert(tris.end(), trigs.begin() + prim + 1, trigs.begin() + prim + 1 + (size_t)count * 20);
        return off;
    }
};

int bvh_ref(const std::vector<int> &bvh, const std::vector<int> &trigs, size_t node, int depth, BvhLayout &b, TriRepack &tr,
            std::unordered_map<int, int> &leaf_map) {
    if (!b.ok) return -1;
    // deeper than the traversal stack of the reference (int nodesToVisit[64], bvh.h:38): undefined there, refused here
    if (node + 6 >= bvh.size() || depth >= 64) { b.ok = false; return -1; }
    const int head = bvh[node];
    if (head <= 0) {
        const int prim = -head;
        auto it = leaf_map.find(prim);   // a leaf block referenced twice (both BVHs share the palette) is stored once
        int off;
        if (it != leaf_map.end()) {
            off = it->second;
        } else {
            off = tr.add(trigs, prim, b.ok);
            leaf_map.emplace(prim, off);
        }
        return -(1 + off);
    }
    const size_t left = node + 7, right = (size_t)head;
    if (left + 6 >= bvh.size() || right + 6 >= bvh.size()) { b.ok = false; return -1; }
    const size_t r = b.rec.size() / 16;
    b.rec.resize(b.rec.size() + 16, 0);
    for (int i = 0; i < 6; i++) {
        b.rec[r * 16 + i] = bvh[left + 1 + i];
        b.rec[r * 16 + 6 + i] = bvh[right + 1 + i];
    }
    const int rl = bvh_ref(bvh, trigs, left, depth + 1, b, tr, leaf_map);
    const int rr = bvh_ref(bvh, trigs, right, depth + 1, b, tr, leaf_map);
    b.rec[r * 16 + 12] = rl;
    b.rec[r * 16 + 13] = rr;
    return (int)r;
}

WideLayout build_wide_layout(const int *tree, size_t n, int depth) {
    WideLayout b;
    int cl = std::max(depth - 7, 4);
    if (const char *e = getenv("CCU_CELL_LEVEL")) cl = std::max(0, atoi(e));   // tuning knob: level of the top table's cells
    if (cl & 1) cl++;
    if (cl > depth) cl = depth & ~1;
    b.cell_level = cl;
    b.top_log2 = depth - cl;
    const int dim = 1 << b.top_log2;
    b.top.assign((size_t)dim * dim * dim, 0u);
    for (int x = 0; x < dim && b.ok; x++)
        for (int y = 0; y < dim; y++)
            for (int z = 0; z < dim; z++) {
                int level = depth;
                int word = tree[0];
                while (word > 0 && level > cl) {
                    level--;
                    int sh = level - cl;
                    size_t at = (size_t)word + ((((x >> sh) & 1) << 2) | (((y >> sh) & 1) << 1) | ((z >> sh) & 1));
                    if (at >= n) { b.ok = false; word = 0; break; }
                    word = tree[at];
                }
                b.top[((size_t)x * dim + y) * dim + z] = word <= 0 ? wide_leaf(word, level, b.ok) : wide_node(tree, n, word, cl, b);
            }
    if (b.wide.empty()) b.wide.assign(64, wide_leaf(0, 0, b.ok));
    return b;
}
}  // namespace

extern "C" {

const char *ccu_last_error(void) { return g_err.c_str(); }
const char *ccu_version(void) { return "chunkycu 0.1 (sm_100a)"; }

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
    if (const char *e = getenv("CCU_WAIT_LANES")) c->wait_lanes = std::max(1, std::min(32, atoi(e)));
    if (const char *e = getenv("CCU_REFILL_MIN")) c->refill_min = std::max(1, std::min(32, atoi(e)));
    if (const char *e = getenv("CCU_EXIT_IDLE")) c->exit_idle = std::max(1, std::min(32, atoi(e)));
    if (const char *e = getenv("CCU_Q_MARCH_WARPS")) c->q_march_warps = std::max(1, std::min(64, atoi(e)));
    if (const char *e = getenv("CCU_Q_BVH_WARPS")) c->q_bvh_warps = std::max(1, std::min(64, atoi(e)));
    if (const char *e = getenv("CCU_Q_LEAF_MIN")) c->q_leaf_min = std::max(1, std::min(32, atoi(e)));
    if (const char *e = getenv("CCU_Q_MARCH_BIAS")) c->q_march_bias = std::max(-32, std::min(32, atoi(e)));
    if (const char *e = getenv("CCU_Q_REFILL_MIN")) c->q_refill_min = std::max(1, std::min(32, atoi(e)));
    if (const char *e = getenv("CCU_YIELD_BELOW")) c->yield_below = std::max(0, std::min(33, atoi(e)));
    if (const char *e = getenv("CCU_BLOCKS_PER_SM")) c->blocks_per_sm = std::max(1, std::min(8, atoi(e)));
    DeviceGuard g(device_index);
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_render_pool<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, POOL_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_render_pool<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, POOL_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_render_pool<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, POOL_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_render_pool<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, POOL_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_render_queue<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, q_smem_bytes(true, true));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_render_queue<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, q_smem_bytes(false, true));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_render_queue<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, q_smem_bytes(true, false));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_render_queue<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, q_smem_bytes(false, false));
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

int ccu_render_end(ccu_ctx *c);

int ccu_ctx_destroy(ccu_ctx *c) {
    if (!c) return CCU_OK;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        DeviceGuard g(c->device);
        cudaStreamSynchronize(c->stream);
        c->tree.release(); c->top.release(); c->wide.release(); c->air_top.release(); c->air_wide.release(); c->air_bits.release(); c->world_rec.release(); c->actor_rec.release(); c->tris2.release(); c->block_palette.release(); c->quad_models.release(); c->aabb_models.release();
        c->mat_palette.release(); c->trigs.release(); c->world_bvh.release(); c->actor_bvh.release();
        c->sun_words.release(); c->atlas.release(); c->sky.release(); c->rays.release(); c->sun_basis.release();
        if (c->accum) cudaFree(c->accum);
        if (c->pinned) cudaFreeHost(c->pinned);
        if (c->seeds_dev) cudaFree(c->seeds_dev);
        if (c->work_counter) cudaFree(c->work_counter);
        if (c->unorm) cudaFree(c->unorm);
        for (auto &e : c->chunk_ev) if (e) cudaEventDestroy(e);
        cudaEventDestroy(c->ev0);
        cudaEventDestroy(c->ev1);
        cudaStreamDestroy(c->stream);
    }
    delete c;
    return CCU_OK;
}

int ccu_scene_begin(ccu_ctx *c) {
    if (!c) return fail(CCU_EINVAL, "ccu_scene_begin: null context");
    std::lock_guard<std::mutex> lk(c->mu);
    c->committed = false;
    return CCU_OK;
}

int ccu_scene_set_octree(ccu_ctx *c, const int32_t *tree, int64_t n, int32_t depth) {
    if (depth < 0 || depth > 30) return fail(CCU_EINVAL, "octree depth %d out of range", depth);
    if (n < 1) return fail(CCU_EINVAL, "octree needs at least the root word");
    int rc = upload_words(c, c->tree, tree, n, "ccu_scene_set_octree");
    if (rc == CCU_OK) {
        c->depth = depth;
        c->tree_host.assign(tree, tree + n);
    }
    return rc;
}
int ccu_scene_set_block_palette(ccu_ctx *c, const int32_t *w, int64_t n) { return upload_words(c, c->block_palette, w, n, "ccu_scene_set_block_palette"); }
int ccu_scene_set_quad_models(ccu_ctx *c, const int32_t *w, int64_t n) { return upload_words(c, c->quad_models, w, n, "ccu_scene_set_quad_models"); }
int ccu_scene_set_aabb_models(ccu_ctx *c, const int32_t *w, int64_t n) { return upload_words(c, c->aabb_models, w, n, "ccu_scene_set_aabb_models"); }
int ccu_scene_set_material_palette(ccu_ctx *c, const int32_t *w, int64_t n) { return upload_words(c, c->mat_palette, w, n, "ccu_scene_set_material_palette"); }
int ccu_scene_set_triangles(ccu_ctx *c, const int32_t *w, int64_t n) {
    int rc = upload_words(c, c->trigs, w, n, "ccu_scene_set_triangles");
    if (rc == CCU_OK) c->trigs_host.assign(w, w + n);
    return rc;
}
int ccu_scene_set_world_bvh(ccu_ctx *c, const int32_t *w, int64_t n) {
    int rc = upload_words(c, c->world_bvh, w, n, "ccu_scene_set_world_bvh");
    if (rc == CCU_OK) { c->world_head.assign(w, w + std::min<int64_t>(n, 7)); c->world_host.assign(w, w + n); }
    return rc;
}
int ccu_scene_set_actor_bvh(ccu_ctx *c, const int32_t *w, int64_t n) {
    int rc = upload_words(c, c->actor_bvh, w, n, "ccu_scene_set_actor_bvh");
    if (rc == CCU_OK) { c->actor_head.assign(w, w + std::min<int64_t>(n, 7)); c->actor_host.assign(w, w + n); }
    return rc;
}

int ccu_scene_atlas_create(ccu_ctx *c, int32_t width, int32_t height, int32_t layers) {
    if (!c) return fail(CCU_EINVAL, "ccu_scene_atlas_create: null context");
    if (width <= 0 || height <= 0 || layers <= 0 || width > 16384 || height > 16384) return fail(CCU_EINVAL, "bad atlas extents %dx%dx%d", width, height, layers);
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    c->committed = false;
    c->atlas.release();
    size_t tiles = (size_t)((width + 15) / 16) * ((height + 15) / 16) * layers;
    c->atlas.n = tiles * 256;
    CU(cudaMalloc(&c->atlas.p, c->atlas.n * sizeof(uchar4)));
    CU(cudaMemsetAsync(c->atlas.p, 0, c->atlas.n * sizeof(uchar4), c->stream));
    c->atlas_w = width; c->atlas_h = height; c->atlas_layers = layers;
    return CCU_OK;
}

int ccu_scene_atlas_write(ccu_ctx *c, int32_t x, int32_t y, int32_t layer, int32_t w, int32_t h, const uint8_t *rgba) {
    if (!c) return fail(CCU_EINVAL, "ccu_scene_atlas_write: null context");
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->atlas.p) return fail(CCU_ESTATE, "atlas not created");
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
    if (!c->tree.p) return fail(CCU_ESTATE, "commit: octree not set");
    if (!c->block_palette.p || !c->mat_palette.p) return fail(CCU_ESTATE, "commit: block/material palette not set");
    if (!c->atlas.p) return fail(CCU_ESTATE, "commit: texture atlas not set");
    if (!c->sky.p) return fail(CCU_ESTATE, "commit: sky not set");
    if (!c->have_sun) return fail(CCU_ESTATE, "commit: sun not set");
    DeviceGuard g(c->device);
    // absent optional palettes behave as the reference's single zero word
    int zero = 0;
    if (!c->quad_models.p) CU(c->quad_models.upload(&zero, 0, c->stream));
    if (!c->aabb_models.p) CU(c->aabb_models.upload(&zero, 0, c->stream));
    if (!c->trigs.p) CU(c->trigs.upload(&zero, 0, c->stream));
    if (!c->trigs.p || c->trigs.n == 0) c->trigs_host.clear();
    if (!c->world_bvh.p) { CU(c->world_bvh.upload(&zero, 0, c->stream)); c->world_head.clear(); c->world_host.clear(); }
    if (!c->actor_bvh.p) { CU(c->actor_bvh.upload(&zero, 0, c->stream)); c->actor_head.clear(); c->actor_host.clear(); }
    // traversal layout: dense top table + 64-ary nodes (two octree levels per load); falls back to the plain
    // reference layout if a leaf value cannot be encoded
    {
        const bool enable = getenv("CCU_NO_WIDE") == nullptr;
        WideLayout wl = build_wide_layout(c->tree_host.data(), c->tree_host.size(), c->depth);
        c->use_wide = (enable && wl.ok) ? 1 : 0;
        c->cell_level = wl.cell_level;
        c->top_log2 = wl.top_log2;
        CU(c->top.upload(wl.top.data(), wl.top.size(), c->stream));
        CU(c->wide.upload(wl.wide.data(), wl.wide.size(), c->stream));
        AirLayout al = build_air_layout(c->tree_host.data(), c->tree_host.size(), c->depth, wl.cell_level);
        c->use_air = al.ok ? 1 : 0;
        CU(c->air_top.upload(al.top.data(), al.top.size(), c->stream));
        CU(c->air_wide.upload(al.wide.data(), al.wide.size(), c->stream));
        CU(c->air_bits.upload(al.bits.data(), al.bits.size(), c->stream));
    }
    // BVH stage layout (pair records + aligned triangle blocks); without it kernel 4 falls back to kernel 3 for BVH scenes
    {
        BvhLayout wb, ab;
        TriRepack tr;
        std::unordered_map<int, int> leaf_map;
        const bool we = bvh_is_empty(c->world_head), ae = bvh_is_empty(c->actor_head);
        if (!we) wb.root = bvh_ref(c->world_host, c->trigs_host, 0, 0, wb, tr, leaf_map);
        if (!ae) ab.root = bvh_ref(c->actor_host, c->trigs_host, 0, 0, ab, tr, leaf_map);
        c->use_bvh2 = (wb.ok && ab.ok) ? 1 : 0;
        c->world_root = wb.root;
        c->actor_root = ab.root;
        if (c->use_bvh2) {
            CU(c->world_rec.upload(wb.rec.data(), wb.rec.size(), c->stream));
            CU(c->actor_rec.upload(ab.rec.data(), ab.rec.size(), c->stream));
            CU(c->tris2.upload(tr.tris.data(), tr.tris.size(), c->stream));
        }
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
    c->committed = true;
    return CCU_OK;
}

int ccu_camera_set(ccu_ctx *c, int32_t projector_type, const float *settings, int64_t n) {
    if (!c || !settings) return fail(CCU_EINVAL, "ccu_camera_set: null argument");
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    if (projector_type == -1) {
        if (n % 6 != 0 || n <= 0) return fail(CCU_EINVAL, "pre-generated rays need 6 floats per pixel (got %lld)", (long long)n);
        cudaStreamSynchronize(c->stream);    // a render in flight may still read the old rays
        CU(c->rays.upload(settings, (size_t)n, c->stream));
    } else {
        if (n < 15) return fail(CCU_EINVAL, "camera settings need 15 floats (got %lld)", (long long)n);
        memcpy(c->cam, settings, sizeof c->cam);
    }
    c->projector_type = projector_type;
    c->have_camera = true;
    return CCU_OK;
}

int ccu_render_end(ccu_ctx *c) {
    if (!c) return fail(CCU_EINVAL, "ccu_render_end: null context");
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    cudaStreamSynchronize(c->stream);
    stop_timer(c);
    // the buffers stay cached for the next render of the same size (freed by ccu_ctx_destroy / a size change)
    c->target_live = false;
    c->window_spp = 0;
    return CCU_OK;
}

int ccu_render_begin(ccu_ctx *c, int32_t width, int32_t height) {
    if (!c) return fail(CCU_EINVAL, "ccu_render_begin: null context");
    if (width <= 0 || height <= 0 || (int64_t)width * height > (1ll << 30)) return fail(CCU_EINVAL, "bad canvas %dx%d", width, height);
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    cudaStreamSynchronize(c->stream);
    size_t n = (size_t)width * height * 3;
    if (c->accum && (size_t)c->width * c->height * 3 != n) {
        cudaFree(c->accum);
        cudaFreeHost(c->pinned);
        c->accum = nullptr;
        c->pinned = nullptr;
    }
    if (!c->accum) {
        CU(cudaMalloc(&c->accum, n * sizeof(float)));
        CU(cudaMallocHost(&c->pinned, n * sizeof(float)));
    }
    CU(cudaMemsetAsync(c->accum, 0, n * sizeof(float), c->stream));
    c->width = width;
    c->height = height;
    c->window_spp = 0;
    c->target_live = true;
    return CCU_OK;
}

int ccu_render_set_params(ccu_ctx *c, const ccu_render_params *p) {
    if (!c || !p) return fail(CCU_EINVAL, "ccu_render_set_params: null argument");
    if (p->draw_depth < 0 || p->draw_depth >= (1 << 24) || p->max_depth < 1 || p->max_depth > 255) return fail(CCU_EINVAL, "bad render params");
    std::lock_guard<std::mutex> lk(c->mu);
    c->params = *p;
    return CCU_OK;
}

static int render_ready(ccu_ctx *c, const char *what) {
    if (!c->committed) return fail(CCU_ESTATE, "%s: scene not committed", what);
    if (!c->have_camera) return fail(CCU_ESTATE, "%s: camera not set", what);
    if (!c->accum || !c->target_live) return fail(CCU_ESTATE, "%s: ccu_render_begin not called", what);
    if (c->projector_type == -1 && c->rays.n != (size_t)c->width * c->height * 6)
        return fail(CCU_ESTATE, "%s: ray buffer holds %zu floats, canvas needs %zu", what, c->rays.n, (size_t)c->width * c->height * 6);
    return CCU_OK;
}

static int render_passes_locked(ccu_ctx *c, const int32_t *seeds, int32_t n_passes, bool first_chunk);

int ccu_render_passes_async(ccu_ctx *c, const int32_t *seeds, int32_t n_passes) {
    if (!c) return fail(CCU_EINVAL, "ccu_render_passes: null context");
    if (n_passes < 0 || (n_passes > 0 && !seeds)) return fail(CCU_EINVAL, "ccu_render_passes: bad seeds");
    std::lock_guard<std::mutex> lk(c->mu);
    int rc = render_ready(c, "ccu_render_passes");
    if (rc != CCU_OK) return rc;
    if (n_passes == 0) return CCU_OK;
    DeviceGuard g(c->device);
    stop_timer(c);
    // one launch covers at most 32768 passes (the default kernel keeps a path's pass index in 16 bits); the reference's
    // windows are at most 1024 passes (OpenClPathTracingRenderer.java:158)
    const int32_t kChunk = 32768;
    for (int32_t off = 0; off < n_passes; off += kChunk) {
        rc = render_passes_locked(c, seeds + off, std::min(kChunk, n_passes - off), off == 0);
        if (rc != CCU_OK) return rc;
    }
    return CCU_OK;
}

static int render_passes_locked(ccu_ctx *c, const int32_t *seeds, int32_t n_passes, bool first_chunk) {
    if (c->seeds_cap < n_passes) {
        cudaStreamSynchronize(c->stream);
        if (c->seeds_dev) cudaFree(c->seeds_dev);
        c->seeds_dev = nullptr;
        c->seeds_cap = 0;
        CU(cudaMalloc(&c->seeds_dev, (size_t)std::max(n_passes, 1024) * sizeof(int)));
        c->seeds_cap = std::max(n_passes, 1024);
    }
    // the seed words are the only per-pass host->device traffic (OpenClPathTracingRenderer.java:106-109)
    CU(cudaMemcpyAsync(c->seeds_dev, seeds, (size_t)n_passes * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));   // seeds[] belongs to the caller again
    fill_scene(c);
    int n_pixels = c->width * c->height;
    if (!c->work_counter) CU(cudaMalloc(&c->work_counter, sizeof(unsigned int)));
    if (first_chunk) CU(cudaEventRecord(c->ev0, c->stream));
    if (c->params.kernel != 1) CU(cudaMemsetAsync(c->work_counter, 0, sizeof(unsigned int), c->stream));
    const bool wide = c->scene.use_wide != 0;
    const bool bvh = !(c->scene.world_bvh_empty && c->scene.actor_bvh_empty);
#define CCU_DISPATCH(KERNEL, GRID, BLOCK, SMEM, ...)                                                              \
    do {                                                                                                          \
        if (bvh && wide) KERNEL<true, true><<<GRID, BLOCK, SMEM, c->stream>>>(__VA_ARGS__);                       \
        else if (bvh) KERNEL<true, false><<<GRID, BLOCK, SMEM, c->stream>>>(__VA_ARGS__);                        \
        else if (wide) KERNEL<false, true><<<GRID, BLOCK, SMEM, c->stream>>>(__VA_ARGS__);                       \
        else KERNEL<false, false><<<GRID, BLOCK, SMEM, c->stream>>>(__VA_ARGS__);                                \
    } while (0)
    if (c->params.kernel == 1) {
        int threads = 128;
        int blocks = (n_pixels + threads - 1) / threads;
        if (wide) k_render_mega<true><<<blocks, threads, 0, c->stream>>>(c->scene, c->seeds_dev, n_passes, c->window_spp, c->accum, n_pixels);
        else k_render_mega<false><<<blocks, threads, 0, c->stream>>>(c->scene, c->seeds_dev, n_passes, c->window_spp, c->accum, n_pixels);
        c->launches++;
    } else {
        // persistent kernels: one resident grid, pixels handed out through a counter
        WaveParams wp;
        wp.seeds = c->seeds_dev;
        wp.n_passes = n_passes;
        wp.start_spp = c->window_spp;
        wp.res = c->accum;
        wp.n_pixels = n_pixels;
        wp.next_pixel = c->work_counter;
        wp.wait_lanes = c->wait_lanes;
        int blocks = c->sm_count * c->blocks_per_sm;
        if (c->params.kernel == 4 || c->params.kernel == 0) {
            // CTA-wide path pool with per-stage work masks (ccu_queue.cuh): one CTA per SM - the fastest, hence the default
            QueueParams qp;
            qp.w = wp;
            qp.yield_below = c->yield_below;
            qp.refill_min = c->q_refill_min;
            qp.march_bias = c->q_march_bias;
            qp.leaf_min = c->q_leaf_min;
            qp.bvh_warps = c->q_bvh_warps;
            qp.march_warps = bvh ? 64 : c->q_march_warps;   // with BVHs the service warps are the ones that do not walk
            const bool tops = c->air_top.n <= (size_t)Q_TOP_WORDS && getenv("CCU_NO_TOPS") == nullptr;
            const int grid = c->sm_count, block = Q_WARPS * 32;
            if (!c->use_air) return fail(CCU_ESTATE, "ccu_render_passes: kernel 4 needs the air layout (malformed octree?)");
            if (bvh && !c->use_bvh2) return fail(CCU_ESTATE, "ccu_render_passes: malformed BVH, or deeper than the 64 levels the reference's traversal stack holds (bvh.h:38)");
            if (tops) {
                if (bvh) k_render_queue<true, true><<<grid, block, q_smem_bytes(true, true), c->stream>>>(c->scene, qp);
                else k_render_queue<false, true><<<grid, block, q_smem_bytes(false, true), c->stream>>>(c->scene, qp);
            } else {
                if (bvh) k_render_queue<true, false><<<grid, block, q_smem_bytes(true, false), c->stream>>>(c->scene, qp);
                else k_render_queue<false, false><<<grid, block, q_smem_bytes(false, false), c->stream>>>(c->scene, qp);
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
                unsigned long long z[32] = {0};
                cudaMemcpyToSymbol(g_qstats, z, sizeof z);
            }
#endif
        } else if (c->params.kernel == 3) {
            // lane-bound state machine (ccu_wavefront.cuh)
            CCU_DISPATCH(k_render_wave, blocks, 256, 0, c->scene, wp);
        } else {
            // per-warp path pool in shared memory (ccu_pool.cuh)
            PoolParams pp;
            pp.w = wp;
            pp.refill_min = c->refill_min;
            pp.exit_idle = c->exit_idle;
            CCU_DISPATCH(k_render_pool, blocks, POOL_WARPS * 32, POOL_SMEM_BYTES, c->scene, pp);
        }
        c->launches++;
    }
#undef CCU_DISPATCH
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev1, c->stream));
    c->timing_pending = true;
    c->window_spp += n_passes;
    return CCU_OK;
}

int ccu_render_sync(ccu_ctx *c) {
    if (!c) return fail(CCU_EINVAL, "ccu_render_sync: null context");
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    CU(cudaStreamSynchronize(c->stream));
    stop_timer(c);
    return CCU_OK;
}

int ccu_render_passes(ccu_ctx *c, const int32_t *seeds, int32_t n_passes) {
    int rc = ccu_render_passes_async(c, seeds, n_passes);
    if (rc != CCU_OK) return rc;
    return ccu_render_sync(c);
}

static int fetch_mean(ccu_ctx *c) {
    size_t n = (size_t)c->width * c->height * 3;
    CU(cudaMemcpyAsync(c->pinned, c->accum, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    stop_timer(c);
    return CCU_OK;
}

int ccu_render_read(ccu_ctx *c, float *mean_rgb, int32_t *window_spp) {
    if (!c || !mean_rgb) return fail(CCU_EINVAL, "ccu_render_read: null argument");
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->accum || !c->target_live) return fail(CCU_ESTATE, "ccu_render_read: no render target");
    DeviceGuard g(c->device);
    int rc = fetch_mean(c);
    if (rc != CCU_OK) return rc;
    memcpy(mean_rgb, c->pinned, (size_t)c->width * c->height * 3 * sizeof(float));
    if (window_spp) *window_spp = c->window_spp;
    return CCU_OK;
}

int ccu_render_merge(ccu_ctx *c, double *sample_buffer, int32_t sample_spp, int32_t *merged_spp) {
    if (!c || !sample_buffer) return fail(CCU_EINVAL, "ccu_render_merge: null argument");
    if (sample_spp < 0) return fail(CCU_EINVAL, "ccu_render_merge: negative spp");
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->accum || !c->target_live) return fail(CCU_ESTATE, "ccu_render_merge: no render target");
    DeviceGuard g(c->device);
    int pass_spp = c->window_spp;
    if (merged_spp) *merged_spp = pass_spp;
    if (pass_spp == 0) {
        CU(cudaStreamSynchronize(c->stream));
        stop_timer(c);
        return CCU_OK;
    }
    // OpenClPathTracingRenderer.java:164-173: blocking read of the float buffer, then the spp-weighted merge on the host.
    // The read-back is cut into chunks so that the merge of chunk k overlaps the copy of chunk k+1.
    const double sinv = 1.0 / (double)(sample_spp + pass_spp);
    const double ds = (double)sample_spp, dp = (double)pass_spp;
    const size_t n = (size_t)c->width * c->height * 3;
    constexpr int NCH = 8;
    if (!c->chunk_ev[0]) {
        for (int k = 0; k < NCH; k++) CU(cudaEventCreateWithFlags(&c->chunk_ev[k], cudaEventDisableTiming));
    }
    const size_t chunk = ((n + NCH - 1) / NCH + 63) & ~(size_t)63;
    for (int k = 0; k < NCH; k++) {
        const size_t lo = std::min(n, k * chunk), hi = std::min(n, (k + 1) * chunk);
        if (hi > lo) CU(cudaMemcpyAsync(c->pinned + lo, c->accum + lo, (hi - lo) * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaEventRecord(c->chunk_ev[k], c->stream));
    }
    unsigned hw = std::thread::hardware_concurrency();
    unsigned nt = std::max(1u, std::min(16u, hw ? hw : 4u));
    if (n < (1u << 20)) nt = 1;
    const float *src = c->pinned;
    cudaEvent_t *evs = c->chunk_ev;
    const int device = c->device;
    auto work = [=](unsigned t) {
        cudaSetDevice(device);
        for (int k = 0; k < NCH; k++) {
            const size_t lo = std::min(n, k * chunk), hi = std::min(n, (k + 1) * chunk);
            cudaEventSynchronize(evs[k]);
            const size_t len = hi - lo, part = (len + nt - 1) / nt;
            const size_t a = lo + std::min(len, t * part), b = lo + std::min(len, (t + 1) * part);
            for (size_t i = a; i < b; i++) sample_buffer[i] = (sample_buffer[i] * ds + (double)src[i] * dp) * sinv;
        }
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; t++) th.emplace_back(work, t);
        for (auto &t : th) t.join();
    }
    CU(cudaStreamSynchronize(c->stream));
    stop_timer(c);
    c->window_spp = 0;   // bufferSppReal = 0 (:170)
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
    if (!c->accum || !c->target_live) return fail(CCU_ESTATE, "ccu_render_device_buffer: no render target");
    *device_ptr = c->accum;
    if (n_floats) *n_floats = (int64_t)c->width * c->height * 3;
    return CCU_OK;
}

int ccu_render_scale(ccu_ctx *c, float factor) {
    if (!c) return fail(CCU_EINVAL, "ccu_render_scale: null context");
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->accum || !c->target_live) return fail(CCU_ESTATE, "ccu_render_scale: no render target");
    DeviceGuard g(c->device);
    size_t n = (size_t)c->width * c->height * 3;
    k_scale<<<c->sm_count * 4, 256, 0, c->stream>>>(c->accum, factor, n);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return CCU_OK;
}

int ccu_stream_handle(ccu_ctx *c, void **cuda_stream) {
    if (!c || !cuda_stream) return fail(CCU_EINVAL, "ccu_stream_handle: null argument");
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
    stop_timer(c);
    size_t n = (size_t)c->width * c->height;
    // one scratch allocation: 4 int planes, t, normal(3), color(4) = 12 words per pixel
    int *scratch = nullptr;
    CU(cudaMalloc(&scratch, n * 12 * sizeof(int)));
    int *d_block = scratch, *d_face = scratch + n, *d_node = scratch + 2 * n, *d_kind = scratch + 3 * n;
    float *d_color = reinterpret_cast<float *>(scratch + 4 * n);          // 16-byte aligned: n*16 bytes offset
    float *d_t = reinterpret_cast<float *>(scratch + 8 * n);
    float *d_normal = reinterpret_cast<float *>(scratch + 9 * n);
    fill_scene(c);
    cudaEventRecord(c->ev0, c->stream);
    if (c->scene.use_wide) k_first_hit<true><<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(c->scene, seed, (int)n, d_block, d_face, d_node, d_kind, d_t, d_normal, d_color);
    else k_first_hit<false><<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(c->scene, seed, (int)n, d_block, d_face, d_node, d_kind, d_t, d_normal, d_color);
    c->launches++;
    cudaEventRecord(c->ev1, c->stream);
    c->timing_pending = true;
    cudaError_t e = cudaGetLastError();
    auto back = [&](void *dst, const void *src, size_t bytes) {
        if (dst && e == cudaSuccess) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream);
    };
    back(block, d_block, n * 4); back(face, d_face, n * 4); back(node, d_node, n * 4); back(kind, d_kind, n * 4);
    back(t, d_t, n * 4); back(normal, d_normal, n * 12); back(color, d_color, n * 16);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(scratch);
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
    stop_timer(c);
    size_t n = (size_t)c->width * c->height;
    int *d = nullptr;
    CU(cudaMalloc(&d, n * sizeof(int)));
    fill_scene(c);
    cudaEventRecord(c->ev0, c->stream);
    if (c->scene.use_wide) k_preview<true><<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(c->scene, (int)n, d);
    else k_preview<false><<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(c->scene, (int)n, d);
    c->launches++;
    cudaEventRecord(c->ev1, c->stream);
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
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    stop_timer(c);
    *ms = c->last_ms;
    return CCU_OK;
}

int ccu_launch_count(ccu_ctx *c, int64_t *launches) {
    if (!c || !launches) return fail(CCU_EINVAL, "ccu_launch_count: null argument");
    *launches = c->launches;
    return CCU_OK;
}

int ccu_scene_device_bytes(ccu_ctx *c, int64_t *bytes) {
    if (!c || !bytes) return fail(CCU_EINVAL, "ccu_scene_device_bytes: null argument");
    std::lock_guard<std::mutex> lk(c->mu);
    *bytes = (int64_t)(c->tree.bytes() + c->top.bytes() + c->wide.bytes() + c->air_top.bytes() + c->air_wide.bytes() + c->air_bits.bytes() + c->world_rec.bytes() + c->actor_rec.bytes() + c->tris2.bytes() + c->block_palette.bytes() + c->quad_models.bytes() + c->aabb_models.bytes() + c->mat_palette.bytes() +
                       c->trigs.bytes() + c->world_bvh.bytes() + c->actor_bvh.bytes() + c->atlas.bytes() + c->sky.bytes());
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
// the value-carrying layout (find_leaf_wide) and the air layout (lean_probe) answer, so that CPU tests can compare both
// with the reference's root descent (octree.h:81-88).
int ccu_debug_layout_lookup(const int32_t *tree, int64_t n, int32_t depth, const int32_t *xyz, int64_t count, int32_t *wide_value,
                            int32_t *wide_level, int32_t *air_solid, int32_t *air_level) {
    if (!tree || n < 1 || depth < 0 || depth > 30 || (count > 0 && !xyz)) return fail(CCU_EINVAL, "ccu_debug_layout_lookup: bad argument");
    WideLayout wl = build_wide_layout(tree, (size_t)n, depth);
    AirLayout al = build_air_layout(tree, (size_t)n, depth, wl.cell_level);
    if (!al.ok) return fail(CCU_ESTATE, "air layout could not be built");
    const int cl = wl.cell_level, tl = wl.top_log2;
    for (int64_t i = 0; i < count; i++) {
        const int bx = xyz[3 * i], by = xyz[3 * i + 1], bz = xyz[3 * i + 2];
        if (((bx | by | bz) >> depth) != 0) return fail(CCU_EINVAL, "voxel %lld outside the octree cube", (long long)i);
        const size_t cell = ((((size_t)(bx >> cl) << tl) + (size_t)(by >> cl)) << tl) + (size_t)(bz >> cl);
        if (wl.ok) {
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
        unsigned e = al.top[cell];
        int lvl = cl;
        while (!(e & CCU_WIDE_LEAF)) {
            if (lvl == 2) {
                const unsigned v = (unsigned)(((bx & 3) << 4) | ((by & 3) << 2) | (bz & 3));
                const unsigned code = (al.bits[(size_t)e * 4 + (v >> 4)] >> ((v & 15u) * 2u)) & 3u;
                e = CCU_WIDE_LEAF | (code == 0 ? 1u : ((code - 1u) << 26));
                break;
            }
            lvl -= 2;
            e = al.wide[(size_t)e * 64 + ((((bx >> lvl) & 3) << 4) | (((by >> lvl) & 3) << 2) | ((bz >> lvl) & 3))];
        }
        if (air_solid) air_solid[i] = (int)(e & 1u);
        if (air_level) air_level[i] = (e & 1u) ? -1 : (int)((e >> 26) & 31);
    }
    return CCU_OK;
}

}  // extern "C"
