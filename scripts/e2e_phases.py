#!/usr/bin/env python3
"""Where the end-to-end step time goes: render_begin / render_passes / render_merge / render_end (host clock)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chunkyclplugin_b200 import native, scenes as S
from chunkyclplugin_b200.javarandom import pass_seeds
p = S.terrain_scene(256, 1920, 1080)
ctx = native.Context(0)
ctx.scene_begin(); ctx.set_atlas(p.atlas); ctx.set_block_palette(p.block_palette); ctx.set_material_palette(p.mat_palette)
ctx.set_aabb_models(p.aabb_models); ctx.set_quad_models(p.quad_models); ctx.set_triangles(p.bvh_trigs)
ctx.set_world_bvh(p.world_bvh); ctx.set_actor_bvh(p.actor_bvh); ctx.set_sun(p.sun); ctx.set_sky(p.sky, p.sky_intensity)
ctx.set_octree(p.octree, p.octree_depth); ctx.scene_commit()
ctx.camera_set(p.projector_type, p.camera)
sb = np.zeros(1920 * 1080 * 3, dtype=np.float64)
seeds = np.asarray(pass_seeds(16), dtype=np.int32)
for it in range(6):
    t0 = time.perf_counter(); ctx.render_begin(1920, 1080)
    t1 = time.perf_counter(); ctx.render_passes(seeds)
    t2 = time.perf_counter(); ctx.render_merge(sb, 0 if it == 0 else 16)
    t3 = time.perf_counter(); ctx.render_end()
    t4 = time.perf_counter()
    print(f"begin {1e3*(t1-t0):.2f} ms  passes {1e3*(t2-t1):.2f} ms (kernel {ctx.last_kernel_ms():.2f})  merge {1e3*(t3-t2):.2f} ms  end {1e3*(t4-t3):.2f} ms  total {1e3*(t4-t0):.2f} ms")
