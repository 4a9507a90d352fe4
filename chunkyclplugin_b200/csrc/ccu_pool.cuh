// ccu_pool.cuh - persistent wavefront path tracer with a per-warp path pool in shared memory.
//
// Paths are decoupled from lanes.  Every warp owns POOL path slots in shared memory (structure of arrays, one
// 32-bit word per field and slot).  A slot is a pixel walking through its passes: running mean, path colour and
// throughput, RNG state, the surface it sits on, and its current ray.  The warp alternates between
//
//   march phase  - each lane holds ONE ray in registers and steps it through octree leaves (octree.h:66-107,
//                  air leaves only).  When the ray leaves the scene or reaches a non-air leaf the lane writes the
//                  ray back to its slot, marks the slot pending and pops the next ready ray from the warp's
//                  ready ring, so the march loop keeps (close to) 32 busy lanes;
//   shade phase  - pending slots are compacted (ballot + find-nth-set-bit) into batches of up to 32 and run through
//                  the same per-path state machine as ccu_wavefront.cuh (block/material test, BVHs, sky, surface
//                  response, sun sampling, bounce, accumulation, next pass / next pixel, camera ray); slots whose
//                  next ray is ready are pushed to the ready ring.
//
// Scheduling does not touch arithmetic: every path performs exactly the operations of the thread-per-pixel
// kernel in the same order (RNG draw order, per-pixel pass order), so the image is bit-identical.
#pragma once
#include "ccu_wavefront.cuh"

namespace ccu {

constexpr int POOL = 64;          // path slots per warp
constexpr int POOL_WARPS = 8;     // warps per CTA

enum PoolField : int {
    F_STATE = 0, F_GID, F_PASS, F_MEANX, F_MEANY, F_MEANZ, F_COLX, F_COLY, F_COLZ, F_THRX, F_THRY, F_THRZ, F_RNG, F_DEPTH,
    F_SPX, F_SPY, F_SPZ, F_SNX, F_SNY, F_SNZ, F_SHW,
    F_OX, F_OY, F_OZ, F_DX, F_DY, F_DZ, F_IX, F_IY, F_IZ, F_T, F_LIMIT, F_STEPS, F_LEAF, F_LEVEL,
    F_COUNT
};
constexpr int POOL_WORDS_PER_WARP = F_COUNT * POOL + POOL;   // fields + ready ring
constexpr int POOL_SMEM_BYTES = POOL_WARPS * POOL_WORDS_PER_WARP * 4;

enum : int {
    LS_READY = 8,       // the slot's next ray sits in the ready ring
    LS_FLIGHT = 9,      // a lane is marching the slot's ray
};
constexpr int SHADOW_BIT = 1 << 16;   // F_DEPTH: ray_depth | SHADOW_BIT when the ray in flight is the shadow ray

__device__ __forceinline__ bool slot_pending(int st) { return st == LS_BLOCK || st == LS_RAY_DONE || st == LS_NEED_PIXEL; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

struct PoolParams {
    WaveParams w;
    int refill_min;    // refill idle lanes from the ready ring once this many lanes are idle
    int exit_idle;     // leave the march phase once this many lanes are idle and the ring is empty
};

template <bool HAS_BVH, bool WIDE>
__global__ void __launch_bounds__(POOL_WARPS * 32, CCU_MIN_BLOCKS) k_render_pool(const __grid_constant__ DScene s, const __grid_constant__ PoolParams pp) {
    extern __shared__ uint32_t pool_mem[];
    const WaveParams &w = pp.w;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    uint32_t *P = pool_mem + (threadIdx.x >> 5) * POOL_WORDS_PER_WARP;
    uint32_t *ring = P + F_COUNT * POOL;
#define FI(f, slot) (*reinterpret_cast<int *>(&P[(f) * POOL + (slot)]))
#define FF(f, slot) (*reinterpret_cast<float *>(&P[(f) * POOL + (slot)]))

    for (int k = lane; k < POOL; k += 32) FI(F_STATE, k) = LS_NEED_PIXEL;
    __syncwarp();
    int ready_head = 0, ready_count = 0;   // warp-uniform
    // the ray this lane is marching
    int my_slot = -1;
    March m;
    m.o = m.d = m.inv = f3(0, 0, 0);
    m.t = 0; m.limit = 0; m.steps = 0;

    for (;;) {
        // ================================================================ shade phase
        for (;;) {
            unsigned m0 = __ballot_sync(full, slot_pending(FI(F_STATE, lane)));
            unsigned m1 = __ballot_sync(full, slot_pending(FI(F_STATE, lane + 32)));
            const int n0 = __popc(m0), n = n0 + __popc(m1);
            if (n == 0) break;
            int slot = -1;
            if (lane < n) slot = lane < n0 ? (int)__fns(m0, 0, lane + 1) : 32 + (int)__fns(m1, 0, lane - n0 + 1);
            int state = LS_EXHAUSTED;
            // ---- load the slot
            int gid = 0, pass = 0, ray_depth = 0, leaf_data = 0, leaf_level = 0;
            float3 mean = f3(0, 0, 0), color = f3(0, 0, 0), throughput = f3(0, 0, 0), surf_point = f3(0, 0, 0), surf_normal = f3(0, 0, 0);
            float shadow_weight = 0;
            uint32_t rng = 0;
            bool shadow = false, ray_hit = false;
            float hit_t = 0;
            Surf hit;
            hit.normal = f3(0, 0, 0); hit.color = make_float4(0, 0, 0, 0); hit.emittance = 0;
            March r;
            r.o = r.d = r.inv = f3(0, 0, 0);
            r.t = 0; r.limit = 0; r.steps = 0;
            if (slot >= 0) {
                state = FI(F_STATE, slot);
                gid = FI(F_GID, slot); pass = FI(F_PASS, slot);
                mean = f3(FF(F_MEANX, slot), FF(F_MEANY, slot), FF(F_MEANZ, slot));
                color = f3(FF(F_COLX, slot), FF(F_COLY, slot), FF(F_COLZ, slot));
                throughput = f3(FF(F_THRX, slot), FF(F_THRY, slot), FF(F_THRZ, slot));
                rng = P[F_RNG * POOL + slot];
                int dw = FI(F_DEPTH, slot);
                ray_depth = dw & 0xFFFF;
                shadow = (dw & SHADOW_BIT) != 0;
                surf_point = f3(FF(F_SPX, slot), FF(F_SPY, slot), FF(F_SPZ, slot));
                surf_normal = f3(FF(F_SNX, slot), FF(F_SNY, slot), FF(F_SNZ, slot));
                shadow_weight = FF(F_SHW, slot);
                r.o = f3(FF(F_OX, slot), FF(F_OY, slot), FF(F_OZ, slot));
                r.d = f3(FF(F_DX, slot), FF(F_DY, slot), FF(F_DZ, slot));
                r.inv = f3(FF(F_IX, slot), FF(F_IY, slot), FF(F_IZ, slot));
                r.t = FF(F_T, slot); r.limit = FF(F_LIMIT, slot); r.steps = FI(F_STEPS, slot);
                leaf_data = FI(F_LEAF, slot); leaf_level = FI(F_LEVEL, slot);
            }
            // ---- per-path state machine (same transitions as k_render_wave)
            for (;;) {
                bool pending = state != LS_MARCH && state != LS_EXHAUSTED;
                if (!__any_sync(full, pending)) break;
                if (state == LS_BLOCK) {
                    if (march_block(s, r, leaf_data, leaf_level, hit, hit_t)) {
                        ray_hit = true;
                        state = LS_RAY_DONE;
                    } else {
                        state = LS_MARCH;
                    }
                }
                if (state == LS_RAY_DONE) {
                    float distance = ray_hit ? hit_t : r.limit;
                    if (HAS_BVH) {
                        int kind = 0;
                        if (bvh_pair(s, r.o, r.d, distance, hit, kind)) ray_hit = true;
                    }
                    if (!ray_hit) {
                        float3 sky = sky_radiance(s, r.d);
                        color = color + (sky * throughput) * (shadow ? shadow_weight : 1.0f);
                        state = shadow ? LS_BOUNCE : LS_SAMPLE_END;
                    } else if (shadow) {
                        state = LS_BOUNCE;
                    } else {
                        surf_point = r.o + r.d * (distance - CCU_OFFSET);
                        surf_normal = hit.normal;
                        float3 col = f3(hit.color.x, hit.color.y, hit.color.z);
                        throughput = throughput * col;
                        color = color + (col * (hit.emittance * s.emitter_scale)) * throughput;
                        if (s.sun_flags & 1) {
                            float x1 = rng_float(rng);
                            float x2 = rng_float(rng);
                            float3 d = sun_sample_direction(s, x1, x2);
                            shadow_weight = fabsf(dot3(d, surf_normal));
                            shadow = true;
                            ray_hit = false;
                            state = march_begin(s, r, surf_point, d, distance) ? LS_MARCH : LS_RAY_DONE;
                        } else {
                            state = LS_BOUNCE;
                        }
                    }
                }
                if (state == LS_BOUNCE) {
                    float x1 = rng_float(rng);
                    float x2 = rng_float(rng);
                    float3 d = diffuse_direction(surf_normal, x1, x2);
                    float3 o = surf_point + d * CCU_OFFSET;
                    ray_depth += 1;
                    if (ray_depth < s.max_depth) {
                        shadow = false;
                        ray_hit = false;
                        state = march_begin(s, r, o, d, inff_()) ? LS_MARCH : LS_RAY_DONE;
                    } else {
                        state = LS_SAMPLE_END;
                    }
                }
                if (state == LS_SAMPLE_END) {
                    int spp = w.start_spp + pass;
                    float fs = (float)spp, fs1 = (float)(spp + 1);
                    mean.x = (mean.x * fs + color.x) / fs1;
                    mean.y = (mean.y * fs + color.y) / fs1;
                    mean.z = (mean.z * fs + color.z) / fs1;
                    pass++;
                    if (pass < w.n_passes) {
                        state = LS_NEW_SAMPLE;
                    } else {
                        float *px = w.res + (size_t)gid * 3;
                        px[0] = mean.x; px[1] = mean.y; px[2] = mean.z;
                        state = LS_NEED_PIXEL;
                    }
                }
                if (state == LS_NEED_PIXEL) {
                    unsigned int k = atomicAdd(w.next_pixel, 1u);
                    if (k < (unsigned int)w.n_pixels) {
                        gid = (int)k;
                        const float *px = w.res + (size_t)gid * 3;
                        mean = f3(px[0], px[1], px[2]);
                        pass = 0;
                        state = LS_NEW_SAMPLE;
                    } else {
                        state = LS_EXHAUSTED;
                    }
                }
                if (state == LS_NEW_SAMPLE) {
                    color = f3(0, 0, 0);
                    throughput = f3(1, 1, 1);
                    ray_depth = 0;
                    rng = (uint32_t)__ldg(w.seeds + pass) + (uint32_t)gid;
                    rng_next(rng);
                    float3 o, d;
                    camera_ray<false>(s, gid, rng, o, d);
                    shadow = false;
                    ray_hit = false;
                    state = march_begin(s, r, o, d, inff_()) ? LS_MARCH : LS_RAY_DONE;
                }
            }
            // ---- store the slot, queue its ray
            const bool ready = slot >= 0 && state == LS_MARCH;
            if (slot >= 0) {
                FI(F_STATE, slot) = ready ? LS_READY : LS_EXHAUSTED;
                FI(F_GID, slot) = gid; FI(F_PASS, slot) = pass;
                FF(F_MEANX, slot) = mean.x; FF(F_MEANY, slot) = mean.y; FF(F_MEANZ, slot) = mean.z;
                FF(F_COLX, slot) = color.x; FF(F_COLY, slot) = color.y; FF(F_COLZ, slot) = color.z;
                FF(F_THRX, slot) = throughput.x; FF(F_THRY, slot) = throughput.y; FF(F_THRZ, slot) = throughput.z;
                P[F_RNG * POOL + slot] = rng;
                FI(F_DEPTH, slot) = ray_depth | (shadow ? SHADOW_BIT : 0);
                FF(F_SPX, slot) = surf_point.x; FF(F_SPY, slot) = surf_point.y; FF(F_SPZ, slot) = surf_point.z;
                FF(F_SNX, slot) = surf_normal.x; FF(F_SNY, slot) = surf_normal.y; FF(F_SNZ, slot) = surf_normal.z;
                FF(F_SHW, slot) = shadow_weight;
                FF(F_OX, slot) = r.o.x; FF(F_OY, slot) = r.o.y; FF(F_OZ, slot) = r.o.z;
                FF(F_DX, slot) = r.d.x; FF(F_DY, slot) = r.d.y; FF(F_DZ, slot) = r.d.z;
                FF(F_IX, slot) = r.inv.x; FF(F_IY, slot) = r.inv.y; FF(F_IZ, slot) = r.inv.z;
                FF(F_T, slot) = r.t; FF(F_LIMIT, slot) = r.limit; FI(F_STEPS, slot) = r.steps;
            }
            const unsigned rm = __ballot_sync(full, ready);
            if (ready) ring[(ready_head + ready_count + __popc(rm & lanemask_lt())) & (POOL - 1)] = (uint32_t)slot;
            ready_count += __popc(rm);
            __syncwarp();
        }
        // ================================================================ march phase
        bool any_work = false;
        int finished = 0;                  // rays that ended in this phase (warp-uniform)
        for (;;) {
            const unsigned idle = __ballot_sync(full, my_slot < 0);
            const int n_idle = __popc(idle);
            if (ready_count > 0 && (n_idle >= pp.refill_min || n_idle == 32)) {
                const int rank = __popc(idle & lanemask_lt());
                if (my_slot < 0 && rank < ready_count) {
                    my_slot = (int)ring[(ready_head + rank) & (POOL - 1)];
                    FI(F_STATE, my_slot) = LS_FLIGHT;
                    m.o = f3(FF(F_OX, my_slot), FF(F_OY, my_slot), FF(F_OZ, my_slot));
                    m.d = f3(FF(F_DX, my_slot), FF(F_DY, my_slot), FF(F_DZ, my_slot));
                    m.inv = f3(FF(F_IX, my_slot), FF(F_IY, my_slot), FF(F_IZ, my_slot));
                    m.t = FF(F_T, my_slot); m.limit = FF(F_LIMIT, my_slot); m.steps = FI(F_STEPS, my_slot);
                }
                const int taken = min(n_idle, ready_count);
                ready_head = (ready_head + taken) & (POOL - 1);
                ready_count -= taken;
            }
            if (__ballot_sync(full, my_slot >= 0) == 0) break;
            any_work = true;
            bool done = false;
            if (my_slot >= 0) {
                int data, level, node;
                int rc = march_probe<WIDE>(s, m, data, level, node);
                if (rc != 0) {
                    FF(F_T, my_slot) = m.t;
                    FI(F_STEPS, my_slot) = m.steps;
                    FI(F_LEAF, my_slot) = data;
                    FI(F_LEVEL, my_slot) = level;
                    FI(F_STATE, my_slot) = rc == 1 ? LS_BLOCK : LS_RAY_DONE;
                    my_slot = -1;
                    done = true;
                }
            }
            finished += __popc(__ballot_sync(full, done));
            // hand over to the shade phase once the ring is dry and enough lanes idle on finished rays
            if (ready_count == 0 && finished > 0 && 32 - __popc(__ballot_sync(full, my_slot >= 0)) >= pp.exit_idle) break;
        }
        __syncwarp();
        if (!any_work) {
            // nothing in flight and nothing ready: finished unless the shade phase can still produce work
            unsigned p0 = __ballot_sync(full, slot_pending(FI(F_STATE, lane)));
            unsigned p1 = __ballot_sync(full, slot_pending(FI(F_STATE, lane + 32)));
            if ((p0 | p1) == 0) break;
        }
    }
#undef FI
#undef FF
}

}  // namespace ccu
