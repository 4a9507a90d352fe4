#!/usr/bin/env python3
"""Quick kernel timing for tuning sweeps: ms per pass of a 16-pass window (device events), first-hit ms, and an md5 of the
accumulation buffer (every variant must print the same hash - scheduling never changes arithmetic).

  CHUNKYCU_LIB=/path/to/variant.so python scripts/qbench.py --workloads config1,indoor --reps 5
"""
import argparse, hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chunkyclplugin_b200 import native  # noqa: E402
from chunkyclplugin_b200.javarandom import pass_seeds  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workloads", default="config1")
ap.add_argument("--passes", type=int, default=16)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--kernel", type=int, default=0)
ap.add_argument("--tag", default=os.environ.get("CHUNKYCU_LIB", "default"))
a = ap.parse_args()
ctx = native.Context(0)
for wl in a.workloads.split(","):
    p = bench.build_scene(wl)
    _, w, h = bench.WORKLOADS[wl]
    t0 = time.perf_counter()
    ctx.scene_begin(); ctx.set_atlas(p.atlas); ctx.set_block_palette(p.block_palette); ctx.set_material_palette(p.mat_palette)
    ctx.set_aabb_models(p.aabb_models); ctx.set_quad_models(p.quad_models); ctx.set_triangles(p.bvh_trigs)
    ctx.set_world_bvh(p.world_bvh); ctx.set_actor_bvh(p.actor_bvh); ctx.set_sun(p.sun); ctx.set_sky(p.sky, p.sky_intensity)
    ctx.set_octree(p.octree, p.octree_depth); ctx.scene_commit()
    load_s = time.perf_counter() - t0
    ctx.camera_set(p.projector_type, p.camera); ctx.render_begin(w, h); ctx.render_set_params(kernel=a.kernel)
    seeds = np.asarray(pass_seeds(a.passes), np.int32)
    ms = []
    for i in range(a.reps + 1):
        ctx.render_reset_window()
        ctx.render_passes(seeds)
        ms.append(ctx.last_kernel_ms())
    img, _ = ctx.render_read()
    fh = []
    lib = native.load()
    for i in range(4):
        native.check(lib.ccu_first_hit(ctx._h, 12345, None, None, None, None, None, None, None))
        fh.append(ctx.last_kernel_ms())
    ctx.render_end()
    best = min(ms[1:])
    print(json.dumps({"tag": a.tag, "workload": wl, "ms_per_pass": best / a.passes, "median_ms_per_pass": float(np.median(ms[1:])) / a.passes,
                      "gsamples_s": w * h * a.passes / (best * 1e-3) / 1e9, "first_hit_ms": min(fh[1:]), "md5": hashlib.md5(img.tobytes()).hexdigest()[:12],
                      "commit_ms": ctx.scene_commit_ms(), "upload_s": load_s, "scene_mb": ctx.scene_device_bytes() / 1e6}), flush=True)
ctx.close()
