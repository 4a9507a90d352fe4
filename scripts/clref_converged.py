#!/usr/bin/env python3
"""North-star acceptance numbers against the reference's UNMODIFIED OpenCL kernel on the same B200 (BASELINE configs 1 / 2):

  * first-hit buffers of the full 1920x1080 frame: block id / hit flag bit-exact?  distance / normal within how many ulp?
  * converged radiance: N passes at 1920x1080 by the reference kernel (one launch per pass, as the Java host does) and by
    libchunkycu through the C ABI, same seeds -> per-pixel relative RMSE.

Writes gpurun_out/clref_converged.json (copy to profiles/).  Test infrastructure: uses oracle/clref."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chunkyclplugin_b200 import native, scenes as S
from chunkyclplugin_b200.javarandom import pass_seeds
from oracle import clref

N_PASSES = int(os.environ.get("PASSES", "256"))


def ulp_diff(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


def main():
    why = clref.available()
    if why is not None:
        print("reference kernel cannot run here:", why)
        return 1
    out = {"passes": N_PASSES, "scenes": {}}
    which = os.environ.get("SCENES", "config1,config3").split(",")
    scenes = []
    if "config1" in which: scenes.append(("config1_terrain256", S.terrain_scene(256, 1920, 1080)))
    if "config3" in which: scenes.append(("config3_indoor256", S.indoor_scene(256, 1920, 1080)))
    if "config4" in which: scenes.append(("config4_entities256", S.entity_scene(256, 1920, 1080)))
    builds = os.environ.get("BUILDS", "stock,strict").split(",")
    for name, p in scenes:
        ctx = native.Context(0)
        ctx.scene_begin(); ctx.set_atlas(p.atlas); ctx.set_block_palette(p.block_palette); ctx.set_material_palette(p.mat_palette)
        ctx.set_aabb_models(p.aabb_models); ctx.set_quad_models(p.quad_models); ctx.set_triangles(p.bvh_trigs)
        ctx.set_world_bvh(p.world_bvh); ctx.set_actor_bvh(p.actor_bvh); ctx.set_sun(p.sun); ctx.set_sky(p.sky, p.sky_intensity)
        ctx.set_octree(p.octree, p.octree_depth); ctx.scene_commit()
        ctx.camera_set(p.projector_type, p.camera); ctx.render_begin(p.width, p.height)
        seeds = pass_seeds(N_PASSES)
        res = {}
        for build in builds:
            ref = clref.ClReference(p, strict=(build == "strict"))
            # ---- config 2: first-hit buffers of the full frame
            fh_ref = ref.first_hit(seeds[0])
            fh = ctx.first_hit(seeds[0])
            hit_ref = fh_ref["hit"] != 0
            hit = fh["kind"] > 0
            both = hit & hit_ref
            r = {"pixels": int(hit.size), "hit_flag_mismatches": int((hit != hit_ref).sum()),
                 "block_id_mismatches": int((fh["block"][both] != fh_ref["block"][both]).sum()),
                 "normal_mismatches": int((fh["normal"].reshape(-1, 3)[both] != fh_ref["normal"].reshape(-1, 3)[both]).any(axis=1).sum()),
                 "distance_max_ulp": int(ulp_diff(fh["t"][both], fh_ref["t"][both]).max()),
                 "distance_bit_exact_pixels": int((ulp_diff(fh["t"][both], fh_ref["t"][both]) == 0).sum()), "hit_pixels": int(both.sum())}
            # ---- config 1 / 3: converged radiance, same seeds
            t0 = time.perf_counter()
            img_ref, times = ref.render(seeds)
            t_ref = time.perf_counter() - t0
            ctx.render_reset_window()
            ctx.render_begin(p.width, p.height)
            ctx.render_passes(seeds)
            img, spp = ctx.render_read()
            assert spp == N_PASSES
            a, b = img.astype(np.float64).reshape(-1, 3), img_ref.astype(np.float64).reshape(-1, 3)
            rmse = float(np.sqrt(np.mean((a - b) ** 2)))
            r.update({"radiance_rmse": rmse, "radiance_rmse_relative_to_mean": rmse / float(b.mean()),
                      "radiance_mean_ours": float(a.mean()), "radiance_mean_reference": float(b.mean()),
                      "pixels_differing_by_more_than_1pct": int((np.abs(a - b).max(axis=1) > 0.01 * np.maximum(b.max(axis=1), 1e-3)).sum()),
                      "bit_identical_pixels": int((img.reshape(-1, 3).view(np.uint32) == img_ref.reshape(-1, 3).view(np.uint32)).all(axis=1).sum()),
                      "reference_ms_per_pass": float(np.median(times)), "ours_ms_per_pass": ctx.last_kernel_ms() / N_PASSES,
                      "device": ref.device_name()})
            res[build] = r
            print(name, build, json.dumps(r))
            ref.close()
        out["scenes"][name] = res
        ctx.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", os.environ.get("OUT", "clref_converged.json")), "w"), indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
