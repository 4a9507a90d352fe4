"""Sample-parallel rendering across 2 GPUs (NCCL): N-GPU image == 1-GPU image up to fp32 summation order."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_passes, out_path):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from chunkyclplugin_b200 import native, scenes as S
        from chunkyclplugin_b200.javarandom import pass_seeds
        from chunkyclplugin_b200.multigpu import SampleParallelRenderer
        from conftest import load_scene
        p = S.terrain_scene(64, 160, 90, seed=7)
        ctx = native.Context(rank)
        load_scene(ctx, p)
        spr = SampleParallelRenderer(ctx, rank, world)
        n_local = spr.render_window(pass_seeds(n_passes))
        out, n = spr.reduce_window(n_local)
        if rank == 0:
            assert n == n_passes
            np.save(out_path, out.cpu().numpy())
        ctx.close()
    finally:
        dist.destroy_process_group()


def test_two_gpu_reduce_matches_one_gpu(tmp_path, scenes, cuda_ctx):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from chunkyclplugin_b200.javarandom import pass_seeds
    from conftest import load_scene
    n_passes = 7
    out_path = str(tmp_path / "combined.npy")
    mp.spawn(_worker, args=(2, _free_port(), n_passes, out_path), nprocs=2, join=True)
    combined = np.load(out_path)
    p = scenes("terrain64")
    load_scene(cuda_ctx, p)
    cuda_ctx.render_passes(pass_seeds(n_passes))
    one, _ = cuda_ctx.render_read()
    assert np.allclose(combined, one, rtol=2e-6, atol=1e-7)


def _worker_merge(rank, world, port, n_passes, out_path):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from chunkyclplugin_b200 import native, scenes as S
        from chunkyclplugin_b200.javarandom import pass_seeds
        from chunkyclplugin_b200.multigpu import SampleParallelRenderer
        from conftest import load_scene
        p = S.terrain_scene(64, 160, 90, seed=7)
        ctx = native.Context(rank)
        load_scene(ctx, p)
        spr = SampleParallelRenderer(ctx, rank, world)
        sb = np.full(p.width * p.height * 3, 0.25, dtype=np.float64) if rank == 0 else None
        total = spr.render_and_merge(pass_seeds(n_passes), sb, sample_spp=3)
        assert total == n_passes
        if rank == 0:
            np.save(out_path, sb)
        ctx.close()
    finally:
        dist.destroy_process_group()


def test_two_gpu_window_merged_into_the_host_sample_buffer(tmp_path, scenes, cuda_ctx):
    """render_and_merge: 2 ranks render, NCCL reduce, rank 0 merges with passSpp = all passes (java :167-173)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from chunkyclplugin_b200.javarandom import pass_seeds
    from conftest import load_scene
    n_passes = 6
    out_path = str(tmp_path / "merged.npy")
    mp.spawn(_worker_merge, args=(2, _free_port(), n_passes, out_path), nprocs=2, join=True)
    merged = np.load(out_path)
    p = scenes("terrain64")
    load_scene(cuda_ctx, p)
    cuda_ctx.render_passes(pass_seeds(n_passes))
    one, _ = cuda_ctx.render_read()
    want = (0.25 * 3 + one.astype(np.float64) * n_passes) / (3 + n_passes)
    assert np.allclose(merged, want, rtol=2e-6, atol=1e-7)
