"""Host-side logic of the sample-parallel multi-GPU path, on CPU with the gloo backend (world_size 2).

The per-rank buffers come from the oracle (each rank renders its share of the passes), so the test checks the real
invariant: striped passes + sum-reduce == the single-rank render of all passes, up to fp32 summation order.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chunkyclplugin_b200.javarandom import pass_seeds
from chunkyclplugin_b200.multigpu import combine_windows, partition_passes, share_bounds


def test_partition_is_a_round_robin_cover():
    seeds = pass_seeds(11)
    for world in (1, 2, 4, 8):
        parts = [partition_passes(seeds, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == sorted(seeds)
        assert parts[0][:2] == [seeds[0], seeds[world]] if len(seeds) > world else True
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_partition_continues_through_a_window():
    """A window rendered in several batches keeps the global pass numbering (ccu_group_render_passes)."""
    seeds = pass_seeds(10)
    for world in (2, 3, 8):
        for r in range(world):
            whole = partition_passes(seeds, r, world)
            split = partition_passes(seeds[:3], r, world) + partition_passes(seeds[3:], r, world, first_pass=3)
            assert whole == split


def test_share_bounds_tile_the_buffer():
    for n in (1, 3 * 75 * 41, 3 * 1920 * 1080, 3 * 3840 * 2160):
        for world in (1, 2, 4, 8):
            b = [share_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert all((hi - lo) % 256 == 0 for lo, hi in b[:-1] if hi < n)


def test_combine_single_process_is_identity_mean():
    m = torch.tensor([1.0, 2.0, 3.0])
    out, n = combine_windows(m.clone(), 4)
    assert n == 4 and torch.allclose(out, m)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_passes, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from chunkyclplugin_b200 import scenes as S
        p = S.terrain_scene(64, 48, 27, seed=7)
        seeds = pass_seeds(n_passes)
        mine = partition_passes(seeds, rank, world)
        local = oracle.Oracle(p).render(mine, threads=2) if mine else np.zeros(48 * 27 * 3, np.float32)
        out, n = combine_windows(torch.from_numpy(local), len(mine), dst=0)
        if rank == 0:
            assert n == n_passes
            np.save(out_path, out.numpy())
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_passes", [6, 1])
def test_two_rank_reduce_matches_single_rank(tmp_path, n_passes):
    import oracle
    from chunkyclplugin_b200 import scenes as S
    out_path = str(tmp_path / "combined.npy")
    mp.spawn(_worker, args=(2, _free_port(), n_passes, out_path), nprocs=2, join=True)
    combined = np.load(out_path)
    p = S.terrain_scene(64, 48, 27, seed=7)
    full = oracle.Oracle(p).render(pass_seeds(n_passes))
    # same samples, different fp32 summation order (per-rank running means, then a weighted sum)
    assert np.allclose(combined, full, rtol=2e-6, atol=1e-7)
    assert combined.max() > 0
