"""Post-processing filters (post_processing_filter.cl:5-51): oracle vs the reference kernel's own output (golden fixture
produced by scripts/clref_tonemap_probe.py on a B200), and the CUDA path vs the oracle."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "clref_tonemap.npz")
FILTERS = {0: "GAMMA", 1: "TONEMAP1", 2: "TONEMAP2 (ACES)", 3: "TONEMAP3 (HABLE)"}


def _channels(argb):
    a = np.asarray(argb).view(np.uint32)
    return np.stack([(a >> s) & 255 for s in (24, 16, 8, 0)], -1).astype(np.int64)


def _special_buffer():
    rng = np.random.default_rng(9)
    w, h = 64, 48
    buf = (rng.random((h, w, 3)) ** 3) * 6.0
    buf[0, :14, :] = np.array([0.0, -0.0, 1.0, 0.004, 0.0039999, 1e-300, 5e-324, 1e300, -1.0, -1e-3, 0.5, 2.0 ** -24, 254.5 / 255.0, 1.0])[:, None]
    buf[1, 0, :] = [np.inf, -np.inf, np.nan]
    return w, h, buf.reshape(-1)


@pytest.mark.parametrize("build", ["strict", "stock"])
@pytest.mark.parametrize("ftype", sorted(FILTERS))
def test_oracle_reproduces_the_reference_kernel(build, ftype):
    """Every pixel of every filter, both builds of the unmodified kernel, incl. inf / nan / negative / denormal inputs."""
    import oracle
    g = np.load(GOLDEN)
    w, h, buf = int(g["width"]), int(g["height"]), g["input"]
    for ei, exposure in enumerate(g["exposures"]):
        got = oracle.tonemap(w, h, float(exposure), buf, ftype)
        want = g[f"{build}_e{ei}_t{ftype}"]
        assert np.array_equal(got, want), f"{FILTERS[ftype]} {build}: {(got != want).sum()} pixels differ"


@pytest.mark.parametrize("ftype", sorted(FILTERS))
def test_oracle_against_a_float64_restatement(ftype):
    """Independent check of the formulas: numpy float64 evaluation of post_processing_filter.cl agrees within 1 LSB."""
    import oracle
    rng = np.random.default_rng(3)
    w, h = 80, 40
    buf = (rng.random(w * h * 3) ** 2) * 5.0
    exposure = 0.8
    c = buf.astype(np.float32).astype(np.float64) * exposure
    if ftype == 0:
        c = c ** (1 / 2.2)
    elif ftype == 1:
        c = np.maximum(0, c - 0.004)
        c = (c * (6.2 * c + 0.5)) / (c * (6.2 * c + 1.7) + 0.06)
    elif ftype == 2:
        c = np.clip((c * (2.51 * c + 0.03)) / (c * (2.43 * c + 0.59) + 0.14), 0, 1) ** (1 / 2.2)
    else:
        f = lambda x: ((x * (0.15 * x + 0.05) + 0.004) / (x * (0.15 * x + 0.5) + 0.06)) - 0.02 / 0.30
        c = f(c * 16) / f(11.2)
    want = np.clip(np.floor(c * 255 + 0.5), 0, 255).reshape(-1, 3)
    got = _channels(oracle.tonemap(w, h, exposure, buf, ftype))
    assert (got[:, 0] == 255).all()
    assert np.abs(got[:, 1:] - want).max() <= 1
    assert (got[:, 1:] != want).mean() < 0.01


def test_oracle_rejects_bad_arguments():
    import oracle
    with pytest.raises(ValueError):
        oracle.tonemap(4, 4, 1.0, np.zeros(48), 7)


@pytest.mark.gpu
@pytest.mark.parametrize("ftype", sorted(FILTERS))
def test_cuda_filter_bit_exact(ftype, cuda_ctx):
    import oracle
    w, h, buf = _special_buffer()
    for exposure in (1.0, 0.37, 4.0):
        got = cuda_ctx.tonemap(w, h, exposure, buf, ftype)
        assert np.array_equal(got, oracle.tonemap(w, h, exposure, buf, ftype)), FILTERS[ftype]
    g = np.load(GOLDEN)
    got = cuda_ctx.tonemap(int(g["width"]), int(g["height"]), float(g["exposures"][1]), g["input"], ftype)
    assert np.array_equal(got, g[f"strict_e1_t{ftype}"])          # == the reference kernel's own output


@pytest.mark.gpu
def test_cuda_filter_full_frame_and_errors(cuda_ctx):
    import oracle
    from chunkyclplugin_b200 import native
    rng = np.random.default_rng(4)
    w, h = 1920, 1080
    buf = rng.random(w * h * 3) * 2.0
    got = cuda_ctx.tonemap(w, h, 1.0, buf, 2)
    assert np.array_equal(got, oracle.tonemap(w, h, 1.0, buf, 2))
    with pytest.raises(native.ChunkyCuError):
        cuda_ctx.tonemap(4, 4, 1.0, np.zeros(48), 9)
    with pytest.raises(ValueError):
        cuda_ctx.tonemap(4, 4, 1.0, np.zeros(47), 0)


@pytest.mark.gpu
def test_filter_class_mirrors_the_reference_filter(cuda_ctx):
    """GpuPostProcessingFilter.processFrame (GpuPostProcessingFilter.java:40-65) through the host mirror."""
    import oracle
    from chunkyclplugin_b200.renderer import BitmapImage, GpuPostProcessingFilter, RendererInstance
    rng = np.random.default_rng(8)
    w, h = 48, 20
    buf = rng.random(w * h * 3) * 3
    for fid, ftype in GpuPostProcessingFilter.IMPOSTERS.items():
        f = GpuPostProcessingFilter(fid, RendererInstance.get(0))
        assert f.getId() == fid
        img = BitmapImage(w, h)
        f.processFrame(w, h, buf, img, 1.5)
        assert np.array_equal(img.data, oracle.tonemap(w, h, 1.5, buf, ftype))
    with pytest.raises(KeyError):
        GpuPostProcessingFilter("NONE")
