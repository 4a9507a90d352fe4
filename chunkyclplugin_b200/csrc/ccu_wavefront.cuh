// ccu_wavefront.cuh - persistent-thread path tracer with warp-level ray re-filling.
//
// The reference is a one-thread-per-pixel mega-kernel (rayTracer.cl:11-113): lanes of a warp idle while the
// longest ray of the warp marches (measured on B200: 9 of 32 lanes active per instruction).  Here every lane
// is a small state machine.  All lanes of a warp execute the same octree march step; a lane whose ray ends
// or reaches a solid block waits until enough lanes of its warp are waiting too, then the waiting lanes are
// processed together (block / material test, sky lookup, surface response, sun sampling, bounce, camera ray, next pixel) and re-enter the march loop with a new ray.
// A lane owns one pixel at a time and walks its passes in order with the running mean in registers, so the RNG
// draw order and the accumulation order (rayTracer.cl:109-112) are exactly the reference's and
// the image is bit-identical to the thread-per-pixel kernel; pixels are handed out through a global counter.
#pragma once
#include "ccu_device.cuh"

#ifndef CCU_MIN_BLOCKS
#define CCU_MIN_BLOCKS 3
#endif

namespace ccu {

enum LaneState : int {
    LS_MARCH = 0,       // a ray is being marched through the octree
    LS_RAY_DONE = 1,    // octree part of closestIntersect finished (hit or left): BVHs, then shading
    LS_BOUNCE = 2,      // nextPath (kernel.h:46-98)
    LS_SAMPLE_END = 3,  // fold the sample into the running mean, next pass
    LS_NEED_PIXEL = 4,  // fetch the next pixel
    LS_NEW_SAMPLE = 5,  // seed the RNG, build the camera ray
    LS_BLOCK = 6,       // a non-air leaf was reached: block / material test (block.h:30-118)
    LS_EXHAUSTED = 7,
};

struct WaveParams {
    const int *seeds;
    int n_passes;
    int start_spp;
    float *res;
    int n_pixels;
    unsigned int *next_pixel;   // global work counter (zeroed before the launch)
    int wait_lanes;             // leave the march loop once this many lanes of the warp are waiting
};

template <bool HAS_BVH, bool WIDE>
__global__ void __launch_bounds__(256, CCU_MIN_BLOCKS) k_render_wave(const __grid_constant__ DScene s, const __grid_constant__ WaveParams w) {
    const unsigned full = 0xffffffffu;
    // pixel / pass
    int gid = -1, pass = 0;
    float3 mean = f3(0, 0, 0);
    // path
    float3 color = f3(0, 0, 0), throughput = f3(1, 1, 1);
    uint32_t rng = 0;
    int ray_depth = 0;
    // current ray
    March m;
    m.o = m.d = m.inv = f3(0, 0, 0);
    m.t = 0; m.limit = 0; m.steps = 0;
    bool shadow = false;        // the ray in flight is the sun-sampling shadow ray of the current surface
    bool ray_hit = false;
    float hit_t = 0;
    Surf hit;                   // surface data of the ray in flight (valid when ray_hit)
    hit.normal = f3(0, 0, 0); hit.color = make_float4(0, 0, 0, 0); hit.emittance = 0;
    // the surface the path currently sits on (needed across the shadow ray for the bounce)
    float3 surf_point = f3(0, 0, 0), surf_normal = f3(0, 0, 0);
    float shadow_weight = 0;
    int leaf_data = 0, leaf_level = 0;   // the non-air leaf awaiting its block test
    int state = LS_NEED_PIXEL;

    for (;;) {
        // ---------------------------------------------------------------- shade / refill phase
        for (;;) {
            bool pending = state != LS_MARCH && state != LS_EXHAUSTED;
            if (!__any_sync(full, pending)) break;
            if (state == LS_BLOCK) {
                if (march_block(s, m, leaf_data, leaf_level, hit, hit_t)) {
                    ray_hit = true;
                    state = LS_RAY_DONE;
                } else {
                    state = LS_MARCH;
                }
            }
            if (state == LS_RAY_DONE) {
                float distance = ray_hit ? hit_t : m.limit;
                if (HAS_BVH) {
                    int kind = 0;
                    if (bvh_pair(s, m.o, m.d, distance, hit, kind)) ray_hit = true;
                }
                if (!ray_hit) {
                    // kernel.h:26-31 with emittance 1 (path segment) or |d.n| (shadow ray)
                    float3 sky = sky_radiance(s, m.d);
                    color = color + (sky * throughput) * (shadow ? shadow_weight : 1.0f);
                    state = shadow ? LS_BOUNCE : LS_SAMPLE_END;
                } else if (shadow) {
                    state = LS_BOUNCE;
                } else {
                    // kernel.h:20-22 + applyRayColor kernel.h:33-44
                    surf_point = m.o + m.d * (distance - CCU_OFFSET);
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
                        // the shadow ray inherits the surface hit's distance as its limit
                        state = march_begin(s, m, surf_point, d, distance) ? LS_MARCH : LS_RAY_DONE;
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
                    state = march_begin(s, m, o, d, inff_()) ? LS_MARCH : LS_RAY_DONE;
                } else {
                    state = LS_SAMPLE_END;
                }
            }
            if (state == LS_SAMPLE_END) {
                // rayTracer.cl:109-112
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
                state = march_begin(s, m, o, d, inff_()) ? LS_MARCH : LS_RAY_DONE;
            }
        }
        unsigned alive = __ballot_sync(full, state == LS_MARCH);
        if (alive == 0) break;
        // ---------------------------------------------------------------- march phase
        const int n_alive = __popc(alive);
        const int min_active = max(n_alive - w.wait_lanes, 0);
        for (;;) {
            if (state == LS_MARCH) {
                int node;
                int r = march_probe<WIDE>(s, m, leaf_data, leaf_level, node);
                if (r == 1) state = LS_BLOCK;
                else if (r == 2) { ray_hit = false; state = LS_RAY_DONE; }
            }
            if (__popc(__ballot_sync(full, state == LS_MARCH)) <= min_active) break;
        }
    }
}

}  // namespace ccu
