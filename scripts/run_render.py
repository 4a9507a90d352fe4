#!/usr/bin/env python3
"""Minimal driver for profiling: loads a scene, renders WINDOWS windows of PASSES passes at WxH."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chunkyclplugin_b200 import native, scenes as S
from chunkyclplugin_b200.javarandom import pass_seeds

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="terrain256")
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--passes", type=int, default=4)
ap.add_argument("--windows", type=int, default=2)
ap.add_argument("--kernel", type=int, default=0)
ap.add_argument("--first-hit", action="store_true")
a = ap.parse_args()
if a.scene == "terrain256":
    p = S.terrain_scene(256, a.width, a.height)
elif a.scene == "indoor":
    p = S.indoor_scene(256, a.width, a.height)
elif a.scene == "entities":
    p = S.entity_scene(256, a.width, a.height)
elif a.scene == "large":
    p = S.large_world_scene(width=a.width, height=a.height)
elif a.scene == "benchmark":
    # the reference's benchmark scene (benchmark/OpenCL_test), packed by oracle/clref/make_ref.py where /root/reference exists
    import dataclasses
    z = np.load(os.path.join(ROOT, "oracle", "_ref", "benchmark_scene.npz"))
    base = S.terrain_scene(8, a.width, a.height)
    p = dataclasses.replace(base, octree=z["octree"], octree_depth=int(z["octree_depth"]), block_palette=z["block_palette"],
                            mat_palette=z["mat_palette"], camera=z["camera"], name="benchmark")
else:
    raise SystemExit("unknown scene")
ctx = native.Context(0)
ctx.scene_begin(); ctx.set_atlas(p.atlas); ctx.set_block_palette(p.block_palette); ctx.set_material_palette(p.mat_palette)
ctx.set_aabb_models(p.aabb_models); ctx.set_quad_models(p.quad_models); ctx.set_triangles(p.bvh_trigs)
ctx.set_world_bvh(p.world_bvh); ctx.set_actor_bvh(p.actor_bvh); ctx.set_sun(p.sun); ctx.set_sky(p.sky, p.sky_intensity)
ctx.set_octree(p.octree, p.octree_depth); ctx.scene_commit()
ctx.camera_set(p.projector_type, p.camera); ctx.render_begin(p.width, p.height)
ctx.render_set_params(kernel=a.kernel)
seeds = pass_seeds(a.passes * a.windows)
for w in range(a.windows):
    ctx.render_reset_window()
    ctx.render_passes(seeds[w * a.passes:(w + 1) * a.passes])
    ms = ctx.last_kernel_ms()
    print(f"window {w}: {ms:.3f} ms, {ms / a.passes:.3f} ms/pass, {p.width * p.height * a.passes / ms / 1e3:.1f} Msamples/s")
if a.first_hit:
    ctx.first_hit(seeds[0]); print("first hit ms", ctx.last_kernel_ms())
img, spp = ctx.render_read()
print("mean", img.mean(), "spp", spp)
