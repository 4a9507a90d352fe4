"""Sample-parallel multi-GPU rendering on top of the library's group API (``ccu_group_*``, include/chunkycu.h).

The reference is single-device (RendererInstance.java:81-101).  Passes are independent given the scene
(``state = seed_p + gid``, rayTracer.cl:55), so pass ``p`` of a window goes to GPU ``p mod N`` with the same seed it
would have had on one GPU: the union of samples is identical to the 1-GPU run and the N-GPU image equals the 1-GPU
image up to fp32 summation order.  The only exchange step is the sum of the per-GPU window buffers:
an NCCL reduce-scatter inside ``libchunkycu.so``; every GPU then reads its share back over its own PCIe link and the
shares are merged into the host sample buffer in parallel.  Nothing here touches device memory - this module is the
thin host-side caller:

* ``SampleParallelRenderer(native.Group(devices=[...]))``  one process drives all GPUs (the JVM plugin's shape);
* ``join_process_group(ctx, rank, world)``                 one process per GPU (``torchrun``): the NCCL id travels over
  the host's own channel (``torch.distributed`` broadcast, any backend) and the sample buffer lives in shared memory
  (``SharedSampleBuffer``) so that every rank can merge its share into it.

``partition_passes`` / ``combine_windows`` restate the partition and the reduction on the host; the CPU test-suite runs
them with the gloo backend (world_size 2) to check the logic the library implements on the device.
"""
from __future__ import annotations

from multiprocessing import shared_memory
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import native


def partition_passes(seeds: Sequence[int], rank: int, world: int, first_pass: int = 0) -> List[int]:
    """Seeds of the passes this rank renders: pass p (numbered through the window) -> rank p mod world."""
    return [int(s) for i, s in enumerate(seeds) if (first_pass + i) % world == rank]


def combine_windows(local_mean, local_spp: int, dst: int = 0, group=None):
    """Host restatement of the reduction (torch tensors, any backend): per-rank window means -> mean over all passes
    on ``dst``.  sum_r mean_r * n_r / sum_r n_r; with equal pass counts this is the sum of means / world, which is what
    ``ccu_group_render_merge`` folds into its merge weight."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    local_mean.mul_(float(local_spp))
    total = torch.tensor([local_spp], dtype=torch.int64, device=local_mean.device)
    if world > 1:
        dist.reduce(local_mean, dst=dst, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    n = int(total.item())
    if rank == dst:
        if n > 0:
            local_mean.div_(float(n))
        return local_mean, n
    return None, n


def share_bounds(n_floats: int, rank: int, world: int) -> Tuple[int, int]:
    """Float range of the sample buffer that ``rank`` merges (what the reduce-scatter hands it): equal shares of the
    buffer length rounded up to a multiple of 256 * world (ccu_group_render_begin)."""
    align = 256 * world
    padded = (n_floats + align - 1) // align * align
    share = padded // world
    return min(n_floats, rank * share), min(n_floats, (rank + 1) * share)


class SharedSampleBuffer:
    """Chunky's double sample buffer in POSIX shared memory, so that the ranks of a one-process-per-GPU job can each merge
    their share into the same buffer.  Rank 0 creates it, the others attach by name."""

    def __init__(self, n_doubles: int, name: Optional[str] = None, create: bool = True):
        if create:
            self.shm = shared_memory.SharedMemory(create=True, size=n_doubles * 8, name=name)
        else:
            self.shm = shared_memory.SharedMemory(name=name)
            try:        # the creator unlinks it; keep this process's resource tracker from reporting it as leaked
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.owner = create
        self.array = np.ndarray((n_doubles,), dtype=np.float64, buffer=self.shm.buf)
        if create:
            self.array.fill(0.0)

    @property
    def name(self) -> str:
        return self.shm.name

    def close(self):
        self.array = None
        try:
            self.shm.close()
            if self.owner:
                self.shm.unlink()
        except Exception:
            pass


def join_process_group(ctx: native.Context, rank: int, world: int) -> native.Group:
    """One process per GPU: rank 0 draws the NCCL id, ``torch.distributed`` (already initialised by the host, any
    backend) carries it to the others, every rank joins with its own context."""
    uid = None
    if world > 1:
        import torch.distributed as dist
        box = [native.group_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    return native.Group(ctx=ctx, unique_id=uid, rank=rank, world=world)


class SampleParallelRenderer:
    """One window end to end on a group: all passes of the window (1-GPU seed order) in, merged sample buffer out."""

    def __init__(self, group: native.Group):
        self.group = group
        self.world = group.world

    def render_window(self, seeds: Sequence[int], sync: bool = True) -> int:
        """Queue the passes on their GPUs (asynchronously on every member) and, unless ``sync`` is off, wait for them."""
        self.group.render_passes(np.asarray(seeds, dtype=np.int32))
        if sync:
            self.group.render_sync()
        return len(seeds)

    def merge(self, sample_buffer: np.ndarray, sample_spp: int) -> int:
        """Reduce-scatter + per-GPU read-back + merge (OpenClPathTracingRenderer.java:164-173 with passSpp = all GPUs'
        passes); returns the number of passes merged.  Collective in a one-process-per-GPU job."""
        return self.group.render_merge(sample_buffer, sample_spp)

    def render_and_merge(self, seeds: Sequence[int], sample_buffer: np.ndarray, sample_spp: int, overlap: bool = False) -> int:
        """``overlap``: return once the reduce-scatter and the share read-backs are queued - the next window renders while the
        shares are merged (call ``finish()`` before reading the buffer)."""
        self.render_window(seeds, sync=False)
        if overlap:
            return self.group.render_merge_async(sample_buffer, sample_spp)
        return self.merge(sample_buffer, sample_spp)

    def finish(self):
        self.group.render_merge_wait()
