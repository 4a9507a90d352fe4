"""CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar: bit-exact for every integer buffer AND every float buffer - both sides follow the same arithmetic
contract: IEEE fp32 ops in source order, no contraction, deterministic transcendentals.
"""
import numpy as np
import pytest

from chunkyclplugin_b200.javarandom import pass_seeds
from conftest import load_scene

pytestmark = pytest.mark.gpu

ALL = ["tiny8", "terrain64", "terrain128", "terrain256", "terrain256_nosun", "decorated", "mixed", "indoor", "entities", "large512"]


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("name", ALL)
def test_first_hit_bit_exact(name, scenes, cuda_ctx):
    import oracle
    p = scenes(name)
    load_scene(cuda_ctx, p)
    seed = pass_seeds(1)[0]
    got = cuda_ctx.first_hit(seed)
    ref = oracle.Oracle(p).first_hit(seed)
    for k in ("block", "face", "node", "kind"):
        assert np.array_equal(got[k], ref[k]), f"{name}: {k} differs in {(got[k] != ref[k]).sum()} pixels"
    for k in ("t", "normal", "color"):
        assert np.array_equal(_bits(got[k]), _bits(ref[k])), f"{name}: {k} differs"
    assert (ref["kind"] > 0).any()


@pytest.mark.parametrize("kernel", [4, 1], ids=["wavefront", "thread_per_pixel"])
@pytest.mark.parametrize("name", ALL)
def test_render_bit_exact(name, kernel, scenes, cuda_ctx):
    """Both render kernels (persistent wavefront = default, thread-per-pixel cross-check) against the oracle."""
    import oracle
    p = scenes(name)
    load_scene(cuda_ctx, p)
    cuda_ctx.render_set_params(kernel=kernel)
    try:
        seeds = pass_seeds(6)
        cuda_ctx.render_passes(seeds)
        got, spp = cuda_ctx.render_read()
    finally:
        cuda_ctx.render_set_params()
    assert spp == 6
    ref = oracle.Oracle(p).render(seeds)
    bad = _bits(got) != _bits(ref)
    assert not bad.any(), f"{name}: {bad.sum()} of {bad.size} floats differ; max abs diff {np.abs(got - ref).max()}"
    assert np.isfinite(got).all() and got.max() > 0


def test_render_depth11_world(scenes, cuda_ctx):
    """BASELINE config 5 shape: a depth-11 octree (top table in global memory, three node levels below it)."""
    import oracle
    p = scenes("large2048")
    load_scene(cuda_ctx, p)
    seeds = pass_seeds(3)
    cuda_ctx.render_passes(seeds)
    got, _ = cuda_ctx.render_read()
    ref = oracle.Oracle(p).render(seeds)
    assert np.array_equal(_bits(got), _bits(ref))
    assert got.max() > 0


@pytest.mark.parametrize("wh", [(75, 41), (33, 70), (1, 1), (31, 5)])
def test_render_odd_canvas_sizes(wh, cuda_ctx):
    """Canvas sizes that are not multiples of the 32x32 pixel tiles the default kernel hands pixels out in."""
    import dataclasses
    import oracle
    from chunkyclplugin_b200 import scenes as S
    p = dataclasses.replace(S.terrain_scene(64, 160, 90, seed=7), width=wh[0], height=wh[1])
    load_scene(cuda_ctx, p)
    seeds = pass_seeds(3)
    cuda_ctx.render_passes(seeds)
    got, _ = cuda_ctx.render_read()
    assert np.array_equal(_bits(got), _bits(oracle.Oracle(p).render(seeds)))


def test_render_top_table_in_global_memory(scenes, cuda_ctx, monkeypatch):
    """The default kernel with its shared-memory staging of the top table switched off gives the same image."""
    import oracle
    p = scenes("terrain64")
    seeds = pass_seeds(4)
    ref = oracle.Oracle(p).render(seeds)
    monkeypatch.setenv("CCU_NO_TOPS", "1")
    load_scene(cuda_ctx, p)
    cuda_ctx.render_passes(seeds)
    got, _ = cuda_ctx.render_read()
    assert np.array_equal(_bits(got), _bits(ref))


def test_render_split_batches_equals_one_batch(scenes, cuda_ctx):
    """Passes issued 1 + 2 + 3 at a time accumulate to the same running mean as one call of 6 (rayTracer.cl:109-112)."""
    p = scenes("terrain64")
    seeds = pass_seeds(6)
    load_scene(cuda_ctx, p)
    cuda_ctx.render_passes(seeds)
    one, _ = cuda_ctx.render_read()
    load_scene(cuda_ctx, p)
    cuda_ctx.render_passes(seeds[:1]); cuda_ctx.render_passes(seeds[1:3]); cuda_ctx.render_passes(seeds[3:])
    many, spp = cuda_ctx.render_read()
    assert spp == 6
    assert np.array_equal(_bits(one), _bits(many))


def test_window_reset_keeps_buffer(scenes, cuda_ctx):
    """bufferSpp = 0 restarts the running mean without clearing the buffer (OpenClPathTracingRenderer.java:170)."""
    import oracle
    p = scenes("terrain64")
    seeds = pass_seeds(5)
    load_scene(cuda_ctx, p)
    cuda_ctx.render_passes(seeds[:3])
    cuda_ctx.render_reset_window()
    cuda_ctx.render_passes(seeds[3:])
    got, spp = cuda_ctx.render_read()
    assert spp == 2
    o = oracle.Oracle(p)
    ref = o.render(seeds[:3])
    ref = o.render(seeds[3:], start_spp=0, res=ref)
    assert np.array_equal(_bits(got), _bits(ref))


def test_pregenerated_rays_camera(scenes, cuda_ctx):
    """projectorType -1: rays come from the host (camera.h:8-11, ClCamera.java:72-105)."""
    import dataclasses
    import oracle
    from chunkyclplugin_b200.scenes import pregenerated_rays
    p = scenes("terrain64")
    q = dataclasses.replace(p, projector_type=-1, camera=pregenerated_rays(p.camera, p.width, p.height))
    load_scene(cuda_ctx, q)
    seeds = pass_seeds(3)
    cuda_ctx.render_passes(seeds)
    got, _ = cuda_ctx.render_read()
    ref = oracle.Oracle(q).render(seeds)
    assert np.array_equal(_bits(got), _bits(ref))
    fh, rf = cuda_ctx.first_hit(seeds[0]), oracle.Oracle(q).first_hit(seeds[0])
    assert np.array_equal(fh["node"], rf["node"]) and np.array_equal(_bits(fh["t"]), _bits(rf["t"]))


def test_preview_bit_exact(scenes, cuda_ctx):
    import oracle
    for name in ("terrain64", "mixed"):
        p = scenes(name)
        load_scene(cuda_ctx, p)
        assert np.array_equal(cuda_ctx.preview(), oracle.Oracle(p).preview())


def test_render_params(scenes, cuda_ctx):
    """drawDepth / maxDepth / emitterScale are launch parameters with the reference constants as defaults."""
    import oracle
    p = scenes("indoor")
    load_scene(cuda_ctx, p)
    cuda_ctx.render_set_params(draw_depth=64, max_depth=3, emitter_scale=7.0)
    try:
        seeds = pass_seeds(3)
        cuda_ctx.render_passes(seeds)
        got, _ = cuda_ctx.render_read()
        ref = oracle.Oracle(p, draw_depth=64, max_depth=3, emitter_scale=7.0).render(seeds)
        assert np.array_equal(_bits(got), _bits(ref))
    finally:
        cuda_ctx.render_set_params()


def test_merge_into_sample_buffer(scenes, cuda_ctx):
    """ccu_render_merge == OpenClPathTracingRenderer.java:167-173 on the host."""
    p = scenes("terrain64")
    load_scene(cuda_ctx, p)
    seeds = pass_seeds(7)
    cuda_ctx.render_passes(seeds[:4])
    mean1, _ = cuda_ctx.render_read()
    sample = np.zeros(p.width * p.height * 3, dtype=np.float64)
    assert cuda_ctx.render_merge(sample, 0) == 4
    assert np.array_equal(sample, (sample * 0 + mean1.astype(np.float64) * 4) * (1.0 / 4))
    cuda_ctx.render_passes(seeds[4:])
    mean2, spp = cuda_ctx.render_read()
    assert spp == 3
    expect = (sample * 4 + mean2.astype(np.float64) * 3) * (1.0 / 7)
    assert cuda_ctx.render_merge(sample, 4) == 3
    assert np.array_equal(sample, expect)


def test_errors_are_loud(cuda_ctx):
    from chunkyclplugin_b200 import native
    c = native.Context(0)
    try:
        with pytest.raises(native.ChunkyCuError):
            c.scene_commit()                       # nothing uploaded
        with pytest.raises(native.ChunkyCuError):
            c._wh = (4, 4); c.render_passes([1])   # no scene / target
        with pytest.raises(native.ChunkyCuError):
            c.render_begin(0, 10)
    finally:
        c.close()
    with pytest.raises(native.ChunkyCuError):
        native.Context(999)


def test_reference_layout_descent_equals_wide_layout(scenes):
    """CCU_NO_WIDE=1 keeps the reference's root-descent layout on the device; images must not change."""
    import os
    from chunkyclplugin_b200 import native
    p = scenes("decorated")
    seeds = pass_seeds(3)
    imgs = []
    for flag in ("1", None):
        if flag:
            os.environ["CCU_NO_WIDE"] = flag
        else:
            os.environ.pop("CCU_NO_WIDE", None)
        c = native.Context(0)
        try:
            load_scene(c, p)
            c.render_passes(seeds)
            imgs.append(c.render_read()[0])
        finally:
            c.close()
    assert np.array_equal(_bits(imgs[0]), _bits(imgs[1]))


def test_two_contexts_render_concurrently_from_two_threads(scenes):
    """Every export is thread-safe per context (SURVEY 8b threading row): two host threads, one context each, at once."""
    import threading
    import oracle
    from chunkyclplugin_b200 import native
    names = ["terrain64", "indoor"]
    seeds = pass_seeds(5)
    out, errs = {}, []

    def worker(name):
        try:
            ctx = native.Context(0)
            p = scenes(name)
            for _ in range(3):
                load_scene(ctx, p)
                ctx.render_passes(seeds)
                out[name], _ = ctx.render_read()
            ctx.close()
        except Exception as e:      # surfaced in the main thread
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(n,)) for n in names]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for n in names:
        assert np.array_equal(_bits(out[n]), _bits(oracle.Oracle(scenes(n)).render(seeds))), n


def test_one_context_shared_by_threads(scenes, cuda_ctx):
    """The camera upload / preview path may be called from another thread than the pass loop (renderLock in the
    reference, OpenClPathTracingRenderer.java:56,103,142): calls serialise on the context's mutex."""
    import threading
    import oracle
    p = scenes("terrain64")
    load_scene(cuda_ctx, p)
    seeds = pass_seeds(8)
    stop = threading.Event()
    errs = []

    def camera_thread():
        try:
            while not stop.is_set():
                cuda_ctx.camera_set(p.projector_type, p.camera)
        except Exception as e:
            errs.append(e)

    t = threading.Thread(target=camera_thread)
    t.start()
    try:
        for i in range(0, 8, 2):
            cuda_ctx.render_passes(seeds[i:i + 2])
    finally:
        stop.set()
        t.join()
    assert not errs, errs
    got, spp = cuda_ctx.render_read()
    assert spp == 8 and np.array_equal(_bits(got), _bits(oracle.Oracle(p).render(seeds)))


def test_batch_longer_than_one_launch(cuda_ctx):
    """A single call with more passes than one launch covers (32768) is split internally; same running mean."""
    import dataclasses
    import oracle
    from chunkyclplugin_b200 import scenes as S
    p = dataclasses.replace(S.terrain_scene(64, 160, 90, seed=7), width=8, height=6)
    seeds = pass_seeds(40000)
    load_scene(cuda_ctx, p)
    cuda_ctx.render_passes(seeds)
    got, spp = cuda_ctx.render_read()
    assert spp == 40000
    assert np.array_equal(_bits(got), _bits(oracle.Oracle(p).render(seeds)))
