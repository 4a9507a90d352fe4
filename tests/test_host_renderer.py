"""The host-side mirror of the reference's renderer classes, driven end to end on the GPU."""
import numpy as np
import pytest

from chunkyclplugin_b200.javarandom import pass_seeds

pytestmark = pytest.mark.gpu


def _manager(packed, target_spp, dump_frequency=0):
    from chunkyclplugin_b200.renderer import DefaultRenderManager, Scene, SnapshotControl
    return DefaultRenderManager(Scene(packed, target_spp=target_spp), SnapshotControl(dump_frequency))


def test_path_tracing_renderer_fills_sample_buffer(scenes):
    """render() = OpenClPathTracingRenderer.render: seeds from Random(0), one merge at the target spp."""
    import oracle
    from chunkyclplugin_b200.renderer import CudaPathTracingRenderer, CudaSceneLoader, RendererInstance
    p = scenes("terrain64")
    r = CudaPathTracingRenderer(CudaSceneLoader(RendererInstance.get(0)))
    assert r.getId() == "ChunkyClRenderer" and r.autoPostProcess() is False
    mgr = _manager(p, 5)
    r.render(mgr)
    s = mgr.bufferedScene
    assert s.spp == 5 and mgr.redraws == 1 and s.post_process_calls == 1
    ref = oracle.Oracle(p).render(pass_seeds(5)).astype(np.float64)
    assert np.array_equal(s.sample_buffer, (0 * 0 + ref * 5) * (1.0 / 5))


def test_render_dump_events_split_windows_and_resume(scenes):
    """A save event every 2 spp forces merges (OpenClPathTracingRenderer.java:150-182); windows are merged with
    spp weights (:167-173), and a second render() call resumes from scene.spp (:91-92)."""
    import oracle
    from chunkyclplugin_b200.renderer import CudaPathTracingRenderer, CudaSceneLoader, RendererInstance
    p = scenes("terrain64")
    r = CudaPathTracingRenderer(CudaSceneLoader(RendererInstance.get(0)))
    mgr = _manager(p, 6, dump_frequency=2)
    r.render(mgr)
    s = mgr.bufferedScene
    assert s.spp == 6 and mgr.redraws == 3
    o = oracle.Oracle(p)
    seeds = pass_seeds(6)
    expect = np.zeros(p.width * p.height * 3)
    done = 0
    for k in range(3):
        w = o.render(seeds[2 * k:2 * k + 2]).astype(np.float64)
        expect = (expect * done + w * 2) * (1.0 / (done + 2))
        done += 2
    assert np.array_equal(s.sample_buffer, expect)
    # resume: the reference restarts Random(0) on every render() call (OpenClPathTracingRenderer.java:95)
    s.target_spp = 8
    r.render(mgr)
    w = o.render(seeds[:2]).astype(np.float64)
    assert s.spp == 8 and np.array_equal(s.sample_buffer, (expect * 6 + w * 2) * (1.0 / 8))


def test_post_render_callback_stops_rendering(scenes):
    from chunkyclplugin_b200.renderer import CudaPathTracingRenderer, CudaSceneLoader, RendererInstance
    p = scenes("terrain64")
    r = CudaPathTracingRenderer(CudaSceneLoader(RendererInstance.get(0)), passes_per_call=1)
    r.CALLBACK_MS = 0.0            # poll after every call (the reference polls at most every 100 ms, OpenClPathTracingRenderer.java:153-157)
    calls = []
    r.setPostRender(lambda: calls.append(1) or len(calls) >= 3)
    mgr = _manager(p, 100)
    r.render(mgr)
    assert mgr.bufferedScene.spp == 3          # stopped by the callback after 3 passes, nothing merged (as the reference)
    assert mgr.redraws == 0


def test_preview_renderer(scenes):
    import oracle
    from chunkyclplugin_b200.renderer import CudaPreviewRenderer, CudaSceneLoader, RendererInstance
    p = scenes("mixed")
    r = CudaPreviewRenderer(CudaSceneLoader(RendererInstance.get(0)))
    assert r.getId() == "ChunkyClPreviewRenderer"
    mgr = _manager(p, 1)
    r.render(mgr)
    assert np.array_equal(mgr.bufferedScene.back_buffer, oracle.Oracle(p).preview())


def test_scene_reload_on_scene_change(scenes):
    """ensureLoad re-uploads when the scene object changes (AbstractSceneLoader.java:46-55)."""
    import oracle
    from chunkyclplugin_b200.renderer import CudaPathTracingRenderer, CudaSceneLoader, RendererInstance
    r = CudaPathTracingRenderer(CudaSceneLoader(RendererInstance.get(0)))
    for name in ("terrain64", "indoor", "terrain64"):
        p = scenes(name)
        mgr = _manager(p, 2)
        r.render(mgr)
        ref = oracle.Oracle(p).render(pass_seeds(2)).astype(np.float64)
        assert np.array_equal(mgr.bufferedScene.sample_buffer, ref)
