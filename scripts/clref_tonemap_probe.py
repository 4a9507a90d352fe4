#!/usr/bin/env python3
"""Runs the reference's UNMODIFIED tonemap kernel (OpenCL, NVIDIA runtime) on the GPU box and writes the golden fixture
tests/golden/clref_tonemap.npz (+ gpurun_out copy): a fixed sample buffer, and for both builds (stock / strict) and all four
filters the ARGB image the reference kernel produced.  Also prints how the oracle compares."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from oracle import clref


def sample_buffer():
    """Deterministic HDR-ish buffer: smooth ramps, a noise block and the special values of double.h / rgba.h."""
    rng = np.random.default_rng(20260117)
    w, h = 96, 64
    buf = np.zeros((h, w, 3))
    ramp = np.linspace(0.0, 1.0, w)
    buf[:16, :, :] = (ramp ** 2.2)[None, :, None] * np.array([1.0, 0.5, 0.25])
    buf[16:32] = (rng.random((16, w, 3)) ** 3) * 8.0
    buf[32:48] = 10.0 ** rng.uniform(-6, 2, (16, w, 3))
    buf[48:] = rng.random((16, w, 3))
    special = [0.0, -0.0, 1.0, 0.004, 0.0039999, 1e-300, 5e-324, 1e300, -1.0, -1e-3, 0.5, 2.0 ** -24, 254.5 / 255.0, 255.0 / 255.0]
    for i, v in enumerate(special):
        buf[48, i, :] = v
    buf[49, 0, :] = [np.inf, -np.inf, np.nan]
    return w, h, buf.reshape(-1)


def main():
    why = clref.available()
    if why is not None or not os.path.exists(clref.TONEMAP_CL):
        print("cannot run the reference tonemap kernel here:", why or "oracle/_ref/chunkycl_tonemap.cl missing")
        return 1
    w, h, buf = sample_buffer()
    out = {"width": np.int32(w), "height": np.int32(h), "input": buf, "exposures": np.array([1.0, 0.37], dtype=np.float32)}
    sh = np.arange(0, 24, 8)[:, None]
    for strict in (False, True):
        tm = clref.ClTonemap(strict=strict)
        for ei, exposure in enumerate(out["exposures"]):
            for t in range(4):
                img, ms = tm.filter(w, h, float(exposure), buf, t)
                key = f"{'strict' if strict else 'stock'}_e{ei}_t{t}"
                out[key] = img
                ref = oracle.tonemap(w, h, float(exposure), buf, t)
                d = np.abs(((img.view(np.uint32) >> sh) & 255).astype(int) - ((ref.view(np.uint32) >> sh) & 255).astype(int))
                print(f"{key}: kernel {ms:.3f} ms; oracle differs in {(img != ref).sum()} of {img.size} pixels, max channel diff {d.max()}")
    for path in (os.path.join(ROOT, "tests", "golden", "clref_tonemap.npz"), os.path.join(ROOT, "gpurun_out", "clref_tonemap.npz")):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez_compressed(path, **out)
    print("wrote tests/golden/clref_tonemap.npz")
    return 0


if __name__ == "__main__":
    sys.exit(main())
