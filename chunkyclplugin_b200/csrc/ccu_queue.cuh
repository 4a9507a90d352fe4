// ccu_queue.cuh - persistent wavefront path tracer with a CTA-wide path pool and per-stage work masks.
//
// One CTA per SM owns Q_SLOTS path slots in shared memory (structure of arrays, one 32-bit word per field and
// slot).  A slot is a pixel walking through its passes: running mean, path colour / throughput, RNG state, the
// surface it sits on and its current ray.  Slot (row, column) lives at word row*32 + column of every field, and
// column c is only ever touched by lane c of whichever warp works on it, so every shared-memory access of a
// stage is bank-conflict free without any data movement between stages.
//
// Work is tracked per stage and column as a bit mask over rows (mask[stage][column], bit = row).  A warp picks
// the stage with the most non-empty columns, every lane pops one slot of its column (atomicAnd), the warp runs
// the stage for up to 32 paths *of the same kind*, and every lane pushes its slot to the next stage (atomicOr):
//
//   MARCH   step the ray through air leaves (octree.h:66-107); a lane whose ray ends pops the next ray of its
//           column right away, so the march loop keeps (close to) 32 busy lanes; when too few lanes are busy and
//           another stage has more work the warp parks its rays (t / steps written back) and switches;
//   BLOCK   block / material test of the non-air leaf the ray reached (block.h:30-118); miss -> MARCH;
//           hit -> BVHs, then surface response + sun sampling (kernel.h:33-44, sky.h:68-93) -> MARCH (shadow ray),
//           or BOUNCE when the ray was the shadow ray;
//   EXIT    the ray left the octree: BVHs, sky + sun disc (kernel.h:26-31) -> BOUNCE (shadow ray) or END;
//   BOUNCE  diffuse bounce (kernel.h:46-98) -> MARCH, or END at the depth limit;
//   END     fold the sample into the running mean (rayTracer.cl:109-112), next pass or next pixel (global work
//           counter), camera ray (rayTracer.cl:55-91) -> MARCH.
//
// Scheduling does not touch arithmetic: every path performs exactly the operations of the thread-per-pixel
// kernel in the same order (RNG draw order, per-pixel pass order), so the image is bit-identical.
#pragma once
#include "ccu_wavefront.cuh"

#ifndef CCU_Q_WARPS
#define CCU_Q_WARPS 28
#endif
#ifndef CCU_Q_ROWS
#define CCU_Q_ROWS 32
#endif

namespace ccu {

constexpr int Q_WARPS = CCU_Q_WARPS;       // warps per CTA (one CTA per SM)
constexpr int Q_ROWS = CCU_Q_ROWS;         // slots per lane column
constexpr int Q_SLOTS = Q_ROWS * 32;
constexpr int Q_MW = (Q_ROWS + 31) / 32;   // mask words per stage and column
static_assert(Q_ROWS >= 1 && Q_ROWS <= 64, "at most two mask words per stage and column");

enum QStage : int { QS_MARCH = 0, QS_BLOCK, QS_EXIT, QS_BOUNCE, QS_END, QS_COUNT };

// Slot fields.  The running mean stays in the accumulation buffer (read-modify-write per sample, L2 only); the
// surface point of the current hit is the origin of the shadow ray and lives in QF_O*.
enum QField : int {
    QF_GID = 0, QF_META, QF_RNG, QF_COLX, QF_COLY, QF_COLZ, QF_THRX, QF_THRY, QF_THRZ,
    QF_SNX, QF_SNY, QF_SNZ, QF_SHW,
    QF_OX, QF_OY, QF_OZ, QF_DX, QF_DY, QF_DZ, QF_IX, QF_IY, QF_IZ, QF_T, QF_LIMIT, QF_STEPS,
    QF_COUNT
};
// QF_META: pass (bits 0..15) | ray depth (bits 16..23) | flags
constexpr uint32_t QM_SHADOW = 1u << 24;     // the ray in flight is the sun-sampling shadow ray of the current surface
constexpr uint32_t QM_NEEDPIX = 1u << 25;    // the slot holds no pixel (initial state)

constexpr int Q_TOP_WORDS = 4096;            // top tables up to 16^3 cells are staged in shared memory
constexpr int Q_MASK_WORDS = QS_COUNT * Q_MW * 32;
constexpr int Q_SMEM_WORDS = QF_COUNT * Q_SLOTS + Q_MASK_WORDS + 32;   // + control words: live, tile lock / base / used
constexpr int Q_SMEM_BYTES = Q_SMEM_WORDS * 4;
constexpr int Q_SMEM_BYTES_TOP = (Q_SMEM_WORDS + Q_TOP_WORDS) * 4;

#ifdef CCU_Q_STATS
// debug counters (build with -DCCU_Q_STATS): [2*st] = executions of stage st, [2*st+1] = lanes that had a slot;
// [10] march iterations, [11] lanes in flight summed over iterations, [12] scheduler rounds that found no work,
// [13] march yields, [14] pop attempts, [15] pop retries
__device__ unsigned long long g_qstats[16];
#define QSTAT(i, v) do { const unsigned long long v_ = (unsigned long long)(v); if ((threadIdx.x & 31) == 0) atomicAdd(&g_qstats[i], v_); } while (0)
#define QSTAT_LANE(i, v) atomicAdd(&g_qstats[i], (unsigned long long)(v))
#else
#define QSTAT(i, v) do { } while (0)
#define QSTAT_LANE(i, v) do { } while (0)
#endif

struct QueueParams {
    WaveParams w;
    int yield_below;   // MARCH: consider switching stage once fewer lanes than this are busy
    int refill_min;    // MARCH: hand over / refill once this many lanes hold a finished ray
    int march_bias;    // scheduler: warps of sub-partitions 0..2 count MARCH columns +bias, warps of sub-partition 3 -bias
};

// ------------------------------------------------------------------------------------------------------
// lean march step: the arithmetic of octree.h:66-107 for an air leaf, with the loop invariants hoisted
// ------------------------------------------------------------------------------------------------------
struct LeanRay {
    float3 o, d, inv, doff;   // doff = d * OFFSET (octree.h:73, loop invariant)
    float t, limit;
    int steps;
    int farx, fary, farz;     // 1 when the leaf cube is left through its upper plane on that axis
};

__device__ __forceinline__ void lean_prepare(LeanRay &r) {
    r.doff = r.d * CCU_OFFSET;
    r.farx = r.inv.x < 0.0f ? 0 : 1;
    r.fary = r.inv.y < 0.0f ? 0 : 1;
    r.farz = r.inv.z < 0.0f ? 0 : 1;
}

// AABB_exit (primitives.h:52-61) along one axis for a point q strictly inside [lo, hi): fmax((lo-q)*inv, (hi-q)*inv).
// lo - q <= 0 < hi - q, so the maximum is the (hi-q)*inv term for inv >= 0 and the (lo-q)*inv term for inv < 0;
// the only case where that term is NaN while the reference's fmax is not is inv = -inf with q == lo, where the
// reference yields the other term, -inf: fmaxf(x, -inf) maps exactly that NaN to -inf and leaves every other x alone.
__device__ __forceinline__ float lean_exit_axis(int b, int level, int far, float q, float inv) {
    float plane = (float)(((b >> level) + far) << level);
    return fmaxf((plane - q) * inv, -inff_());
}

// One iteration of octree.h:66-107 on the air layout (DScene::air_*).  Returns 0 = air leaf left (keep marching),
// 1 = non-air leaf reached (ray not advanced; the block stage looks the leaf up), 2 = ray finished without a hit.
// `top` is the air top table (shared or global).
__device__ __forceinline__ int lean_probe(const DScene &s, const unsigned *__restrict__ top, LeanRay &r) {
    if (r.steps >= s.draw_depth || r.t > r.limit) return 2;
    float3 pos = r.o + r.d * r.t;
    float3 q = pos + r.doff;
    int bx = f2i(floorf(q.x)), by = f2i(floorf(q.y)), bz = f2i(floorf(q.z));
    if (((bx | by | bz) >> s.depth) != 0) return 2;
    const int cl = s.cell_level;
    unsigned e = top[((((unsigned)(bx >> cl) << s.top_log2) + (unsigned)(by >> cl)) << s.top_log2) + (unsigned)(bz >> cl)];
    // structured so that the lanes of a warp reconverge before the exit arithmetic: every path ends in a leaf entry
    if (!(e & CCU_WIDE_LEAF)) {
        int lvl = cl;
        while (lvl > 2 && !(e & CCU_WIDE_LEAF)) {
            lvl -= 2;
            e = __ldg(s.air_wide + (e * 64u + (unsigned)((((bx >> lvl) & 3) << 4) | (((by >> lvl) & 3) << 2) | ((bz >> lvl) & 3))));
        }
        if (!(e & CCU_WIDE_LEAF)) {
            // 4^3 voxels, 2 bits each: 0 = not air, 1 = air leaf of level 0, 2 = air leaf of level 1
            const unsigned v = (unsigned)(((bx & 3) << 4) | ((by & 3) << 2) | (bz & 3));
            const unsigned code = (__ldg(s.air_bits + (e * 4u + (v >> 4))) >> ((v & 15u) * 2u)) & 3u;
            e = CCU_WIDE_LEAF | (code == 0 ? 1u : ((code - 1u) << 26));
        }
    }
    if (e & 1u) return 1;
    const int level = (e >> 26) & 31;
    float ex = lean_exit_axis(bx, level, r.farx, q.x, r.inv.x);
    float ey = lean_exit_axis(by, level, r.fary, q.y, r.inv.y);
    float ez = lean_exit_axis(bz, level, r.farz, q.z, r.inv.z);
    r.t += fminf(ex, fminf(ey, ez)) + CCU_OFFSET;
    r.steps++;
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// work masks: mask[(stage * Q_MW + word) * 32 + column], bit = row within the word
// ------------------------------------------------------------------------------------------------------
// Lowest set bit first: warps that pop the same stage at the same time contend for the same rows, and the winner
// takes (most of) a row across all columns.  Rows therefore tend to stay together from stage to stage, which keeps
// the batches full; spreading the warps over different rows (measured) fragments the batches and is slower.
__device__ __forceinline__ int q_pop_word(unsigned *word) {
    unsigned m = *reinterpret_cast<volatile unsigned *>(word);
    while (m) {
        const unsigned bit = m & (0u - m);
        const unsigned old = atomicAnd(word, ~bit);
        QSTAT_LANE(14, 1);
        if (old & bit) return __ffs((int)bit) - 1;
        QSTAT_LANE(15, 1);
        m = old & ~bit;
    }
    return -1;
}
__device__ __forceinline__ int q_pop(unsigned *mask, int stage, int lane) {
    int row = q_pop_word(mask + (stage * Q_MW) * 32 + lane);
    if (Q_MW > 1 && row < 0) {
        row = q_pop_word(mask + (stage * Q_MW + 1) * 32 + lane);
        if (row >= 0) row += 32;
    }
    if (row >= 0) __threadfence_block();
    return row;
}
__device__ __forceinline__ void q_push(unsigned *mask, int stage, int lane, int row) {
    __threadfence_block();
    atomicOr(mask + (stage * Q_MW + (Q_MW > 1 ? (row >> 5) : 0)) * 32 + lane, 1u << (row & 31));
}
__device__ __forceinline__ bool q_has_work(const unsigned *mask, int stage, int lane) {
    unsigned m = *reinterpret_cast<const volatile unsigned *>(mask + (stage * Q_MW) * 32 + lane);
    if (Q_MW > 1) m |= *reinterpret_cast<const volatile unsigned *>(mask + (stage * Q_MW + 1) * 32 + lane);
    return m != 0;
}

#define QI(f) (*reinterpret_cast<int *>(&F[(f) * Q_SLOTS + slot]))
#define QU(f) (F[(f) * Q_SLOTS + slot])
#define QFL(f) (*reinterpret_cast<float *>(&F[(f) * Q_SLOTS + slot]))

// store the ray a stage has just started (march_begin done) and queue it
__device__ __forceinline__ void q_store_ray(uint32_t *F, unsigned *mask, int lane, int row, const March &m, bool entered) {
    const int slot = row * 32 + lane;
    QFL(QF_OX) = m.o.x; QFL(QF_OY) = m.o.y; QFL(QF_OZ) = m.o.z;
    QFL(QF_DX) = m.d.x; QFL(QF_DY) = m.d.y; QFL(QF_DZ) = m.d.z;
    QFL(QF_IX) = m.inv.x; QFL(QF_IY) = m.inv.y; QFL(QF_IZ) = m.inv.z;
    QFL(QF_T) = m.t; QFL(QF_LIMIT) = m.limit; QI(QF_STEPS) = m.steps;
    // a ray that starts outside the octree cube and never enters it is finished already (octree.h:53-64)
    q_push(mask, entered ? QS_MARCH : QS_EXIT, lane, row);
}

// ------------------------------------------------------------------------------------------------------
// stages
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void q_load_lean(const uint32_t *F, int slot, LeanRay &r) {
    r.o = f3(__uint_as_float(F[QF_OX * Q_SLOTS + slot]), __uint_as_float(F[QF_OY * Q_SLOTS + slot]), __uint_as_float(F[QF_OZ * Q_SLOTS + slot]));
    r.d = f3(__uint_as_float(F[QF_DX * Q_SLOTS + slot]), __uint_as_float(F[QF_DY * Q_SLOTS + slot]), __uint_as_float(F[QF_DZ * Q_SLOTS + slot]));
    r.inv = f3(__uint_as_float(F[QF_IX * Q_SLOTS + slot]), __uint_as_float(F[QF_IY * Q_SLOTS + slot]), __uint_as_float(F[QF_IZ * Q_SLOTS + slot]));
    r.t = __uint_as_float(F[QF_T * Q_SLOTS + slot]);
    r.limit = __uint_as_float(F[QF_LIMIT * Q_SLOTS + slot]);
    r.steps = (int)F[QF_STEPS * Q_SLOTS + slot];
    lean_prepare(r);
}

// MARCH
__device__ __forceinline__ void q_stage_march(const DScene &s, const unsigned *__restrict__ top, uint32_t *F, unsigned *mask, int lane,
                                              int yield_below, int refill_min) {
    const unsigned full = 0xffffffffu;
    int cur = -1;   // slot of the ray this lane is marching (row * 32 + column), -1 = none
    LeanRay r;
    r.o = r.d = r.inv = r.doff = f3(0, 0, 0);
    r.t = 0; r.limit = 0; r.steps = 0; r.farx = r.fary = r.farz = 0;
    int done = 0;   // 0 = ray in flight, 1 = reached a non-air leaf, 2 = finished
    int recheck = 0;
    int n_done = 0, n_fly = 0;   // lanes holding a finished ray / a ray in flight (warp-uniform)
    for (;;) {
        // Finished rays are handed over and idle lanes re-filled in batches: once refill_min lanes hold a finished
        // ray, or nothing is in flight.
        if (n_done >= refill_min || n_fly == 0) {
            if (cur < 0 || done != 0) {
                if (cur >= 0) {
                    F[QF_T * Q_SLOTS + cur] = __float_as_uint(r.t);
                    F[QF_STEPS * Q_SLOTS + cur] = (uint32_t)r.steps;
                    q_push(mask, done == 1 ? QS_BLOCK : QS_EXIT, cur & 31, cur >> 5);
                    done = 0;
                }
                const int row = q_pop(mask, QS_MARCH, lane);
                cur = row < 0 ? -1 : row * 32 + lane;
                if (cur >= 0) q_load_lean(F, cur, r);
            }
            const int busy = __popc(__ballot_sync(full, cur >= 0));
            if (busy == 0) return;
            if (busy < yield_below) {
                // few busy lanes: switch when another stage has more lanes' worth of work than this warp is using
                if (recheck == 0) {
                    int best = 0;
#pragma unroll
                    for (int st = QS_BLOCK; st < QS_COUNT; st++) best = max(best, __popc(__ballot_sync(full, q_has_work(mask, st, lane))));
                    if (best > busy) { QSTAT(13, 1); break; }
                    recheck = 4;
                }
                recheck--;
            }
        }
        if (cur >= 0 && done == 0) done = lean_probe(s, top, r);
        const unsigned fly = __ballot_sync(full, cur >= 0 && done == 0);
        n_fly = __popc(fly);
        QSTAT(10, 1); QSTAT(11, n_fly);
        n_done = __popc(__ballot_sync(full, cur >= 0 && done != 0));
    }
    // park the rays still in flight, hand over the finished ones
    if (cur >= 0) {
        F[QF_T * Q_SLOTS + cur] = __float_as_uint(r.t);
        F[QF_STEPS * Q_SLOTS + cur] = (uint32_t)r.steps;
        q_push(mask, done == 0 ? QS_MARCH : (done == 1 ? QS_BLOCK : QS_EXIT), cur & 31, cur >> 5);
    }
}

// BLOCK (kind_block = true) and EXIT (false): the rest of closestIntersect (kernel.h:14-24) and what follows it in
// rayTracer.cl:93-106.  kind_block is warp-uniform.
template <bool HAS_BVH>
__device__ __forceinline__ void q_stage_resolve(const DScene &s, uint32_t *F, unsigned *mask, int lane, const bool kind_block) {
    const int row = q_pop(mask, kind_block ? QS_BLOCK : QS_EXIT, lane);
    QSTAT(2 * (kind_block ? QS_BLOCK : QS_EXIT), 1); QSTAT(2 * (kind_block ? QS_BLOCK : QS_EXIT) + 1, __popc(__ballot_sync(0xffffffffu, row >= 0)));
    if (row < 0) return;
    const int slot = row * 32 + lane;
    March m;
    m.o = f3(QFL(QF_OX), QFL(QF_OY), QFL(QF_OZ));
    m.d = f3(QFL(QF_DX), QFL(QF_DY), QFL(QF_DZ));
    m.inv = f3(QFL(QF_IX), QFL(QF_IY), QFL(QF_IZ));
    m.t = QFL(QF_T); m.limit = QFL(QF_LIMIT);
    m.steps = QI(QF_STEPS);
    bool ray_hit = false;
    float hit_t = 0;
    Surf hit;
    hit.normal = f3(0, 0, 0); hit.color = make_float4(0, 0, 0, 0); hit.emittance = 0;
    if (kind_block) {
        // the leaf under the ray (octree.h:81-88): value and level
        const Cell c = march_cell(m);
        int level, node;
        const int data = s.use_wide ? find_leaf_wide(s, c.bx, c.by, c.bz, level) : find_leaf(s, c.bx, c.by, c.bz, level, node);
        if (!march_block(s, m, data, level, hit, hit_t)) {
            QFL(QF_T) = m.t; QI(QF_STEPS) = m.steps;
            q_push(mask, QS_MARCH, lane, row);
            return;
        }
        ray_hit = true;
    }
    float distance = ray_hit ? hit_t : m.limit;
    if (HAS_BVH) {
        int kind = 0;
        if (bvh_pair(s, m.o, m.d, distance, hit, kind)) ray_hit = true;
    }
    const uint32_t meta = QU(QF_META);
    const bool shadow = (meta & QM_SHADOW) != 0;
    if (!ray_hit) {
        // kernel.h:26-31 with emittance 1 (path segment) or |d.n| (shadow ray)
        float3 throughput = f3(QFL(QF_THRX), QFL(QF_THRY), QFL(QF_THRZ));
        float3 color = f3(QFL(QF_COLX), QFL(QF_COLY), QFL(QF_COLZ));
        float3 sky = sky_radiance(s, m.d);
        color = color + (sky * throughput) * (shadow ? QFL(QF_SHW) : 1.0f);
        QFL(QF_COLX) = color.x; QFL(QF_COLY) = color.y; QFL(QF_COLZ) = color.z;
        q_push(mask, shadow ? QS_BOUNCE : QS_END, lane, row);
    } else if (shadow) {
        q_push(mask, QS_BOUNCE, lane, row);
    } else {
        // kernel.h:20-22 + applyRayColor kernel.h:33-44
        float3 throughput = f3(QFL(QF_THRX), QFL(QF_THRY), QFL(QF_THRZ));
        float3 color = f3(QFL(QF_COLX), QFL(QF_COLY), QFL(QF_COLZ));
        float3 surf_point = m.o + m.d * (distance - CCU_OFFSET);
        float3 col = f3(hit.color.x, hit.color.y, hit.color.z);
        throughput = throughput * col;
        color = color + (col * (hit.emittance * s.emitter_scale)) * throughput;
        QFL(QF_THRX) = throughput.x; QFL(QF_THRY) = throughput.y; QFL(QF_THRZ) = throughput.z;
        QFL(QF_COLX) = color.x; QFL(QF_COLY) = color.y; QFL(QF_COLZ) = color.z;
        QFL(QF_SNX) = hit.normal.x; QFL(QF_SNY) = hit.normal.y; QFL(QF_SNZ) = hit.normal.z;
        if (s.sun_flags & 1) {
            uint32_t rng = QU(QF_RNG);
            float x1 = rng_float(rng);
            float x2 = rng_float(rng);
            QU(QF_RNG) = rng;
            float3 d = sun_sample_direction(s, x1, x2);
            QFL(QF_SHW) = fabsf(dot3(d, hit.normal));
            QU(QF_META) = meta | QM_SHADOW;
            // the shadow ray starts at the surface point and inherits the surface hit's distance as its limit
            const bool entered = march_begin(s, m, surf_point, d, distance);
            q_store_ray(F, mask, lane, row, m, entered);
        } else {
            QFL(QF_OX) = surf_point.x; QFL(QF_OY) = surf_point.y; QFL(QF_OZ) = surf_point.z;
            q_push(mask, QS_BOUNCE, lane, row);
        }
    }
}

// kernel.h:46-98.  QF_O* holds the surface point (the origin of the shadow ray, or written by the resolve stage).
__device__ __forceinline__ void q_stage_bounce(const DScene &s, uint32_t *F, unsigned *mask, int lane) {
    const int row = q_pop(mask, QS_BOUNCE, lane);
    QSTAT(2 * QS_BOUNCE, 1); QSTAT(2 * QS_BOUNCE + 1, __popc(__ballot_sync(0xffffffffu, row >= 0)));
    if (row < 0) return;
    const int slot = row * 32 + lane;
    uint32_t rng = QU(QF_RNG);
    float x1 = rng_float(rng);
    float x2 = rng_float(rng);
    QU(QF_RNG) = rng;
    float3 n = f3(QFL(QF_SNX), QFL(QF_SNY), QFL(QF_SNZ));
    float3 d = diffuse_direction(n, x1, x2);
    float3 o = f3(QFL(QF_OX), QFL(QF_OY), QFL(QF_OZ)) + d * CCU_OFFSET;
    uint32_t meta = QU(QF_META);
    int ray_depth = (int)((meta >> 16) & 0xFF) + 1;
    meta = (meta & ~(0xFFu << 16) & ~QM_SHADOW) | ((uint32_t)ray_depth << 16);
    QU(QF_META) = meta;
    if (ray_depth < s.max_depth) {
        March m;
        const bool entered = march_begin(s, m, o, d, inff_());
        q_store_ray(F, mask, lane, row, m, entered);
    } else {
        q_push(mask, QS_END, lane, row);
    }
}

constexpr int Q_TILE_W = 32, Q_TILE_H = 32;
constexpr unsigned Q_CHUNK = Q_TILE_W * Q_TILE_H;
// k-th pixel in tile order (bands of Q_TILE_H rows, each cut into tiles Q_TILE_W wide, row-major inside a tile;
// the last band / last tile of a band may be smaller): a bijection of [0, W*H) that keeps consecutive k close on screen
__device__ __forceinline__ int tile_order_pixel(unsigned k, int W, int H) {
    const unsigned band_px = (unsigned)W * Q_TILE_H;
    const unsigned band = k / band_px;
    const unsigned kb = k - band * band_px;
    const unsigned bh = min((unsigned)Q_TILE_H, (unsigned)H - band * Q_TILE_H);     // rows in this band
    const unsigned tile_px = Q_TILE_W * bh;
    const unsigned full_tiles = (unsigned)W / Q_TILE_W;
    unsigned tile = kb / tile_px;
    unsigned tw = Q_TILE_W;
    unsigned in = kb - tile * tile_px;
    if (tile >= full_tiles) {            // the narrower remainder tile at the right edge
        tile = full_tiles;
        tw = (unsigned)W - full_tiles * Q_TILE_W;
        in = kb - full_tiles * tile_px;
    }
    const unsigned py = band * Q_TILE_H + in / tw;
    const unsigned px = tile * Q_TILE_W + in % tw;
    return (int)(py * (unsigned)W + px);
}

// rayTracer.cl:109-112, then the next pass / pixel: rayTracer.cl:55-91
__device__ __forceinline__ void q_stage_end(const DScene &s, const WaveParams &w, uint32_t *F, unsigned *mask, int *live, int lane) {
    const unsigned full = 0xffffffffu;
    const int row = q_pop(mask, QS_END, lane);
    QSTAT(2 * QS_END, 1); QSTAT(2 * QS_END + 1, __popc(__ballot_sync(full, row >= 0)));
    const int slot = row < 0 ? lane : row * 32 + lane;
    uint32_t meta = row >= 0 ? QU(QF_META) : 0u;
    int gid = row >= 0 ? QI(QF_GID) : 0;
    int pass = (int)(meta & 0xFFFFu);
    bool need_pixel = row >= 0 && (meta & QM_NEEDPIX) != 0;
    if (row >= 0 && !need_pixel) {
        // the running mean of this pixel is only ever touched by the slot that owns the pixel; L2 accesses keep it
        // coherent between the warps that run the slot's END stages
        float *px = w.res + (size_t)gid * 3;
        float3 mean = f3(__ldcg(px), __ldcg(px + 1), __ldcg(px + 2));
        float3 color = f3(QFL(QF_COLX), QFL(QF_COLY), QFL(QF_COLZ));
        int spp = w.start_spp + pass;
        float fs = (float)spp, fs1 = (float)(spp + 1);
        mean.x = (mean.x * fs + color.x) / fs1;
        mean.y = (mean.y * fs + color.y) / fs1;
        mean.z = (mean.z * fs + color.z) / fs1;
        __stcg(px, mean.x); __stcg(px + 1, mean.y); __stcg(px + 2, mean.z);
        pass++;
        if (pass >= w.n_passes) need_pixel = true;
    }
    // Pixels are handed out in screen tiles: the CTA draws chunks of Q_TILE_W * Q_TILE_H tile-order indices from the
    // global counter and its warps take their pixels from the CTA's current chunk, so the paths resident on one SM
    // stay in a compact screen region (L1 hit rate of the octree walk).  The order pixels are rendered in does not
    // change any pixel's samples.
    const unsigned want = __ballot_sync(full, need_pixel);
    bool alive = row >= 0;
    if (want) {
        const int n = __popc(want);
        const int leader = __ffs((int)want) - 1;
        unsigned base1 = 0, base2 = 0;
        int n1 = 0;
        if (lane == leader) {
            volatile unsigned *ctl = reinterpret_cast<volatile unsigned *>(live);
            while (atomicCAS(reinterpret_cast<unsigned *>(live) + 1, 0u, 1u) != 0u) __nanosleep(20);
            __threadfence_block();
            const unsigned used = ctl[3];
            n1 = min(n, (int)(Q_CHUNK - used));
            base1 = ctl[2] + used;
            if (n1 < n) {
                base2 = atomicAdd(w.next_pixel, (unsigned)Q_CHUNK);
                ctl[2] = base2;
                ctl[3] = (unsigned)(n - n1);
            } else {
                ctl[3] = used + (unsigned)n;
            }
            __threadfence_block();
            atomicExch(reinterpret_cast<unsigned *>(live) + 1, 0u);
        }
        base1 = __shfl_sync(full, base1, leader);
        base2 = __shfl_sync(full, base2, leader);
        n1 = __shfl_sync(full, n1, leader);
        if (need_pixel) {
            const int rank = __popc(want & ((1u << lane) - 1u));
            const unsigned k = rank < n1 ? base1 + (unsigned)rank : base2 + (unsigned)(rank - n1);
            if (k < (unsigned)w.n_pixels) {
                gid = tile_order_pixel(k, s.width, s.height);
                pass = 0;
            } else {
                alive = false;
            }
        }
    }
    const unsigned died = __ballot_sync(full, row >= 0 && !alive);
    if (died && lane == (__ffs((int)died) - 1)) atomicSub(live, __popc(died));
    if (!alive) return;
    // new sample
    QI(QF_GID) = gid;
    QU(QF_META) = (uint32_t)pass;   // ray depth 0, no flags
    QFL(QF_COLX) = 0.0f; QFL(QF_COLY) = 0.0f; QFL(QF_COLZ) = 0.0f;
    QFL(QF_THRX) = 1.0f; QFL(QF_THRY) = 1.0f; QFL(QF_THRZ) = 1.0f;
    uint32_t rng = (uint32_t)__ldg(w.seeds + pass) + (uint32_t)gid;
    rng_next(rng);
    float3 o, d;
    camera_ray<false>(s, gid, rng, o, d);
    QU(QF_RNG) = rng;
    March m;
    const bool entered = march_begin(s, m, o, d, inff_());
    q_store_ray(F, mask, lane, row, m, entered);
}

// TOPS: the top table of the air layout fits Q_TOP_WORDS and is staged in shared memory
template <bool HAS_BVH, bool TOPS>
__global__ void __launch_bounds__(Q_WARPS * 32, 1) k_render_queue(const __grid_constant__ DScene s, const __grid_constant__ QueueParams qp) {
    extern __shared__ uint32_t q_mem[];
    uint32_t *F = q_mem;
    unsigned *mask = q_mem + QF_COUNT * Q_SLOTS;
    int *live = reinterpret_cast<int *>(mask + Q_MASK_WORDS);
    unsigned *top_s = mask + Q_MASK_WORDS + 32;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < Q_SLOTS; i += blockDim.x) F[QF_META * Q_SLOTS + i] = QM_NEEDPIX;
    for (int i = threadIdx.x; i < Q_MASK_WORDS; i += blockDim.x) {
        const int st = i / (Q_MW * 32), word = (i / 32) % Q_MW;
        const int rows = min(32, Q_ROWS - 32 * word);
        mask[i] = st == QS_END ? (rows == 32 ? 0xffffffffu : ((1u << rows) - 1u)) : 0u;
    }
    if (TOPS) {
        const int n = 1 << (3 * s.top_log2);
        for (int i = threadIdx.x; i < n; i += blockDim.x) top_s[i] = __ldg(s.air_top + i);
    }
    if (threadIdx.x == 0) {
        live[0] = Q_SLOTS;
        live[1] = 0;                 // tile lock
        live[2] = 0;                 // base of the CTA's current chunk of tile-order indices
        live[3] = (int)Q_CHUNK;      // indices of the chunk handed out so far (exhausted: the first request draws a chunk)
    }
    __syncthreads();
    const unsigned *top = TOPS ? top_s : s.air_top;

    for (;;) {
        // the stage with the most columns that have work
        int best = -1, best_n = 0;
#pragma unroll
        for (int st = 0; st < QS_COUNT; st++) {
            int n = __popc(__ballot_sync(full, q_has_work(mask, st, lane)));
            // Warps of one SM sub-partition share an instruction cache: three sub-partitions lean towards the (small)
            // march loop, the fourth towards the (large) shading stages.
            if (st == QS_MARCH && n > 0) n = max(1, n + ((threadIdx.x >> 5 & 3) == 3 ? -qp.march_bias : qp.march_bias));
            if (n > best_n) { best_n = n; best = st; }
        }
        if (best < 0) {
            if (*reinterpret_cast<volatile int *>(live) == 0) break;
            QSTAT(12, 1);
            __nanosleep(100);
            continue;
        }
        switch (best) {
            case QS_MARCH: QSTAT(0, 1); q_stage_march(s, top, F, mask, lane, qp.yield_below, qp.refill_min); break;
            case QS_BLOCK:
            case QS_EXIT: q_stage_resolve<HAS_BVH>(s, F, mask, lane, best == QS_BLOCK); break;
            case QS_BOUNCE: q_stage_bounce(s, F, mask, lane); break;
            default: q_stage_end(s, qp.w, F, mask, live, lane); break;
        }
        __syncwarp();
    }
}

#undef QI
#undef QU
#undef QFL

}  // namespace ccu
