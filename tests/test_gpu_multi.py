"""Sample-parallel rendering through the library's group API (ccu_group_*): N-GPU image == 1-GPU image up to fp32
summation order.  The 1-member group runs on any GPU box; the 2-GPU cases skip (loudly) below 2 devices."""
import os
import socket

import numpy as np
import pytest

from chunkyclplugin_b200.javarandom import pass_seeds
from conftest import load_scene

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _n_gpus():
    from chunkyclplugin_b200 import native
    return native.device_count()


def _one_gpu_merge(cuda_ctx, p, n_passes, sample_spp, base):
    load_scene(cuda_ctx, p)
    cuda_ctx.render_passes(pass_seeds(n_passes))
    one, _ = cuda_ctx.render_read()
    return (base * sample_spp + one.astype(np.float64) * n_passes) / (sample_spp + n_passes)


def _load_group_scene(group, p):
    m0 = group.member(0)
    load_scene(m0, p)                       # upload + commit on member 0 only ...
    m0.render_end()
    group.replicate_scene()                 # ... device-to-device copies for the others
    group.camera_set(p.projector_type, p.camera)
    group.render_begin(p.width, p.height)


def test_group_of_one_is_the_single_gpu_path(scenes, cuda_ctx):
    """world = 1: no NCCL, same window merge as ccu_render_merge - and bit-identical to it."""
    from chunkyclplugin_b200 import native
    p = scenes("terrain64")
    want = _one_gpu_merge(cuda_ctx, p, 5, 3, 0.25)
    group = native.Group(devices=[0])
    try:
        assert (group.world, group.local_members) == (1, 1)
        _load_group_scene(group, p)
        sb = np.full(p.width * p.height * 3, 0.25, dtype=np.float64)
        seeds = pass_seeds(5)
        group.render_passes(seeds[:2])
        group.render_passes(seeds[2:])       # a window may be rendered in several batches
        assert group.render_merge(sb, 3) == 5
        assert np.array_equal(sb, want)
        assert group.render_merge(sb, 8) == 0   # empty window: nothing merged
        group.render_end()
    finally:
        group.close()


@pytest.mark.parametrize("n_passes", [6, 7])
def test_two_gpu_group_matches_one_gpu(n_passes, scenes, cuda_ctx):
    """One process, two GPUs (ncclCommInitAll): replicated scene, striped passes, reduce-scatter, per-GPU share merge.
    7 passes = unequal per-GPU counts (window sums are scaled on the device), 6 = equal (folded into the merge weight)."""
    if _n_gpus() < 2:
        pytest.skip("NEEDS 2 GPUs - the 2-GPU group path is not exercised on this box")
    from chunkyclplugin_b200 import native
    p = scenes("terrain64")
    want = _one_gpu_merge(cuda_ctx, p, n_passes, 3, 0.25)
    group = native.Group(devices=[0, 1])
    try:
        assert (group.world, group.local_members) == (2, 2)
        _load_group_scene(group, p)
        sb = np.full(p.width * p.height * 3, 0.25, dtype=np.float64)
        group.render_passes(pass_seeds(n_passes))
        group.render_sync()
        assert group.render_merge(sb, 3) == n_passes
        render_ms, reduce_ms = group.last_ms()
        assert render_ms > 0 and reduce_ms > 0
        assert np.allclose(sb, want, rtol=2e-6, atol=1e-7)
        # second window into the same buffer, merged while a third one renders (ccu_group_render_merge_async)
        more = pass_seeds(n_passes + 10)[n_passes:]
        group.render_passes(more[:4])
        assert group.render_merge_async(sb, 3 + n_passes) == 4
        group.render_passes(more[4:])
        group.render_merge_wait()
        assert group.reduce_only() == 6          # device part only: the sums stay on the GPUs ...
        assert group.render_merge(sb, 7 + n_passes) == 6     # ... until the next merge reads them back
        group.render_end()
    finally:
        group.close()
    assert np.isfinite(sb).all() and sb.max() > 0


def _worker(rank, world, port, n_passes, shm_name, n_doubles):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)       # only carries the NCCL id and the barriers
    try:
        from chunkyclplugin_b200 import native, scenes as S
        from chunkyclplugin_b200.multigpu import SampleParallelRenderer, SharedSampleBuffer, join_process_group
        p = S.terrain_scene(64, 160, 90, seed=7)
        ctx = native.Context(rank)
        load_scene(ctx, p)
        ctx.render_end()
        group = join_process_group(ctx, rank, world)
        group.camera_set(p.projector_type, p.camera)
        group.render_begin(p.width, p.height)
        sb = SharedSampleBuffer(n_doubles, name=shm_name, create=False)
        spr = SampleParallelRenderer(group)
        total = spr.render_and_merge(pass_seeds(n_passes), sb.array, 3)   # every rank merges its share into shared memory
        assert total == n_passes
        dist.barrier()
        group.render_end()
        group.close()
        ctx.close()
        sb.close()
    finally:
        dist.destroy_process_group()


def test_one_process_per_gpu_merges_into_shared_memory(scenes, cuda_ctx):
    """torchrun shape: ccu_group_join on every rank, the sample buffer in shared memory."""
    if _n_gpus() < 2:
        pytest.skip("NEEDS 2 GPUs - the one-process-per-GPU group path is not exercised on this box")
    import torch.multiprocessing as mp
    from chunkyclplugin_b200.multigpu import SharedSampleBuffer
    p = scenes("terrain64")
    n_passes = 6
    want = _one_gpu_merge(cuda_ctx, p, n_passes, 3, 0.25)
    sb = SharedSampleBuffer(p.width * p.height * 3)
    try:
        sb.array.fill(0.25)
        mp.spawn(_worker, args=(2, _free_port(), n_passes, sb.name, sb.array.size), nprocs=2, join=True)
        assert np.allclose(sb.array, want, rtol=2e-6, atol=1e-7)
    finally:
        sb.close()
