#!/usr/bin/env python3
"""Runs the reference's own OpenCL kernels on the GPU box and compares them with the oracle and the CUDA path.

Writes gpurun_out/clref_report.json, the PTX the NVIDIA JIT produced (gpurun_out/clref_*.ptx) and golden
fixtures (gpurun_out/clref_golden_*.npz) that are then committed under tests/golden/.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle                                    # noqa: E402
from oracle import clref                         # noqa: E402
from chunkyclplugin_b200 import scenes as S      # noqa: E402
from chunkyclplugin_b200.javarandom import pass_seeds   # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
report = {}

why = clref.available()
print("clref available:", why or "yes")
if why:
    json.dump({"unavailable": why}, open(os.path.join(OUT, "clref_report.json"), "w"))
    sys.exit(0)


def ulp_diff(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


def compare_first_hit(name, p, ref, fh):
    hit_o = (fh["kind"] > 0).astype(np.int32)
    r = {}
    r["pixels"] = int(hit_o.size)
    r["hit_mismatch"] = int((ref["hit"] != hit_o).sum())
    both = (ref["hit"] == 1) & (hit_o == 1)
    r["block_mismatch"] = int((ref["block"][both] != fh["block"][both]).sum())
    nr = ref["normal"].reshape(-1, 3)[both]
    no = fh["normal"].reshape(-1, 3)[both]
    r["normal_mismatch"] = int((nr != no).any(axis=1).sum())
    same = both.copy()
    same[both] &= (ref["block"][both] == fh["block"][both]) & ~(nr != no).any(axis=1)
    ud = ulp_diff(ref["t"][same], fh["t"][same])
    r["t_max_ulp"] = int(ud.max()) if ud.size else 0
    r["t_bit_exact_frac"] = float((ud == 0).mean()) if ud.size else 1.0
    cd = np.abs(ref["color"].reshape(-1, 4)[same] - fh["color"].reshape(-1, 4)[same])
    r["color_max_abs"] = float(cd.max()) if cd.size else 0.0
    r["color_bit_exact_frac"] = float((cd == 0).all(axis=1).mean()) if cd.size else 1.0
    ray_o = oracle.Oracle(p).camera_rays(pass_seeds(1)[0])
    r["ray_max_ulp"] = int(ulp_diff(ref["ray"], ray_o.reshape(-1)).max())
    print(name, r)
    return r


seed0 = pass_seeds(1)[0]
cases = {
    "terrain256": S.terrain_scene(256, 480, 270),
    "terrain64": S.terrain_scene(64, 160, 90, seed=7),
    "indoor": S.indoor_scene(64, 128, 72),
    "entities": S.entity_scene(128, 160, 90, n_world=96, n_actor=8, subdiv=1),
    "decorated_quads": S.terrain_scene(64, 160, 90, seed=11, decorate=True),
}

for strict in (True, False):
    tag = "strict" if strict else "stock"
    for name, p in cases.items():
        t0 = time.time()
        ref = clref.ClReference(p, strict=strict)
        if name == "terrain256":
            print(ref.device_name(), "| build options:", ref.options, "| build+upload s:", round(time.time() - t0, 2))
            open(os.path.join(OUT, f"clref_{tag}.ptx"), "wb").write(ref.binary())
            report["device"] = ref.device_name()
        o = oracle.Oracle(p)
        fh_o = o.first_hit(seed0)
        fh_r = ref.first_hit(seed0)
        report[f"first_hit/{tag}/{name}"] = compare_first_hit(f"[{tag}] {name}", p, fh_r, fh_o)
        # radiance: 64 passes, reference kernel vs oracle (running mean both)
        seeds = pass_seeds(64)
        img_r, times = ref.render(seeds)
        img_o = o.render(seeds)
        a, b = img_r.reshape(-1, 3).astype(np.float64), img_o.reshape(-1, 3).astype(np.float64)
        rmse = float(np.sqrt(((a - b) ** 2).mean()))
        rel = rmse / float(np.sqrt((b ** 2).mean()))
        lum_r, lum_o = a.mean(), b.mean()
        report[f"radiance64/{tag}/{name}"] = {"rmse": rmse, "rel_rmse": rel, "mean_ref": lum_r, "mean_oracle": lum_o,
                                             "bit_exact_frac": float((img_r == img_o).mean()),
                                             "kernel_ms_per_pass": float(np.median(times))}
        print(f"[{tag}] {name} radiance64:", report[f"radiance64/{tag}/{name}"])
        pv_r, _ = ref.preview()
        pv_o = o.preview()
        d = np.abs(((pv_r.view(np.uint32)[:, None] >> np.array([16, 8, 0], dtype=np.uint32)) & 255).astype(np.int32) -
                   ((pv_o.view(np.uint32)[:, None] >> np.array([16, 8, 0], dtype=np.uint32)) & 255).astype(np.int32))
        report[f"preview/{tag}/{name}"] = {"exact_frac": float((pv_r == pv_o).mean()), "max_channel_diff": int(d.max()),
                                          "frac_gt1": float((d.max(axis=1) > 1).mean())}
        print(f"[{tag}] {name} preview:", report[f"preview/{tag}/{name}"])
        if name in ("terrain64", "indoor", "entities", "decorated_quads"):
            np.savez_compressed(os.path.join(OUT, f"clref_golden_{tag}_{name}.npz"),
                                seed=np.int32(seed0), hit=fh_r["hit"].astype(np.int8), block=fh_r["block"], t=fh_r["t"],
                                normal=fh_r["normal"].astype(np.int8), color=fh_r["color"], ray=fh_r["ray"],
                                radiance64=img_r, preview=pv_r)
        if name == "terrain256":
            # math builtins vs detmath
            x = np.linspace(-7, 7, 100001).astype(np.float32)
            u = np.linspace(-1, 1, 100001).astype(np.float32)
            for fn, code, xs in (("sin", 0, x), ("cos", 1, x), ("asin", 3, u), ("acos", 4, u)):
                got = ref.math(code, xs)
                mine = oracle.math_fn(fn, xs)
                report[f"math/{tag}/{fn}_max_ulp_vs_detmath"] = int(ulp_diff(got, mine)[np.isfinite(got)].max())
            rng = np.random.default_rng(1)
            a_, b_ = rng.normal(size=100001).astype(np.float32), rng.normal(size=100001).astype(np.float32)
            report[f"math/{tag}/atan2_max_ulp_vs_detmath"] = int(ulp_diff(ref.math(2, a_, b_), oracle.math_fn("atan2", a_, b_)).max())
            pos = np.abs(a_) + 1e-3
            report[f"math/{tag}/recip_exact"] = bool(np.array_equal(ref.math(5, pos), (np.float32(1) / pos).astype(np.float32)))
            report[f"math/{tag}/sqrt_exact"] = bool(np.array_equal(ref.math(6, pos), np.sqrt(pos).astype(np.float32)))
            report[f"math/{tag}/div_exact"] = bool(np.array_equal(ref.math(7, a_, pos), (a_ / pos).astype(np.float32)))
            print({k: v for k, v in report.items() if k.startswith(f"math/{tag}")})
        ref.close()

# same-box performance bar: the reference kernel at 1080p, one launch per pass as its host does
p = S.terrain_scene(256, 1920, 1080)
ref = clref.ClReference(p, strict=False)
_, times = ref.render(pass_seeds(20))
fh = ref.first_hit(seed0)
report["perf/stock/terrain256_1080p"] = {"render_ms_per_pass_median": float(np.median(times[4:])), "render_ms_all": [round(t, 3) for t in times],
                                         "first_hit_ms": fh["ms"],
                                         "samples_per_s": 1920 * 1080 / (float(np.median(times[4:])) * 1e-3)}
print(report["perf/stock/terrain256_1080p"])
ref.close()
json.dump(report, open(os.path.join(OUT, "clref_report.json"), "w"), indent=1)
