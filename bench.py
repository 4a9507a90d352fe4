#!/usr/bin/env python3
"""Benchmark of the ChunkyCL render path on B200 (contract: see the task statement / DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one window of SPP_PER_STEP path-tracing passes over the full 1920x1080 canvas of the BASELINE config-1
scene (synthetic 256^3 terrain, sun + sky, no entities) on every rank, followed - for N > 1 - by the NCCL sum-reduce
of the per-GPU window buffers (weak scaling: per-GPU work is fixed, the N-GPU job renders N x the samples).
`value` is device-timed (CUDA events on the launching stream, max over ranks) with the scene resident in HBM;
`e2e` goes through the host renderer class with host buffers (seed upload + readback/merge into the double
sample buffer inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 1920, 1080
SPP_PER_STEP = 16
METRIC = "path samples/sec"
UNIT = "samples/s"
WORKLOAD = "config1: synthetic 256^3 terrain octree, 1920x1080, 16 spp per step, sun+sky, no entities (path tracing, max depth 5)"
# The contract line is config1 (the configuration the metric is quoted on that fits one GPU).  --workload selects one of the
# other BASELINE configurations for an extra measurement (same JSON shape, named in config.workload); they are parity-test
# cases first (tests/test_gpu_parity.py), bench lines second.
WORKLOADS = {
    "config1": WORKLOAD,
    "indoor": "config3: 256^3 carved rooms with emissive blocks, sunlight disabled, 1920x1080, 16 spp per step (max depth 5 = 4 bounces)",
    "entities": "config4: config1 terrain + synthetic triangle meshes in world/actor BVHs, 1920x1080, 16 spp per step",
    "large": "config5: 2048x256x2048 world in a depth-11 octree, 3840x2160, 16 spp per step",
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake", nv.nvmlClocksThrottleReasonSyncBoost: "sync_boost",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def build_scene(workload: str = "config1"):
    from chunkyclplugin_b200 import scenes as S
    if workload == "indoor":
        return S.indoor_scene(256, WIDTH, HEIGHT)
    if workload == "entities":
        return S.entity_scene(256, WIDTH, HEIGHT)
    if workload == "large":
        return S.large_world_scene(width=WIDTH, height=HEIGHT)
    return S.terrain_scene(256, WIDTH, HEIGHT)


def cpu_port_rate(scene, seeds, stride: int, threads: int = 0):
    """The oracle (CPU port of the same algorithm, OpenMP over pixels) on a strided pixel subset; returns
    (samples/s, counters, n_samples, cores)."""
    import oracle
    o = oracle.Oracle(scene)
    gids = np.arange(0, scene.width * scene.height, stride, dtype=np.int32)
    t0 = time.perf_counter()
    threads = threads or (os.cpu_count() or 1)          # torchrun exports OMP_NUM_THREADS=1; use every host core anyway
    o.render(seeds, gids=gids, threads=threads)
    dt = time.perf_counter() - t0
    n = gids.size * len(seeds)
    return n / dt, o.last_counters, n, threads


def run_reference(args):
    """--impl reference: the CPU arm.  Chunky's own Java path tracer cannot run here (no JVM, chunky-core not
    vendored), so the reference arm is the oracle port of the same algorithm on all host cores (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from chunkyclplugin_b200.javarandom import pass_seeds
    scene = build_scene(args.workload)
    seeds = pass_seeds(SPP_PER_STEP)
    stride = 16                       # each step = every 16th pixel of the 1080p frame x 16 spp = 2.07 M samples
    rates = []
    for i in range(args.warmup + args.steps):
        rate, counters, n, cores = cpu_port_rate(scene, seeds, stride)
        if i >= args.warmup:
            rates.append((rate, n))
    total = sum(n for _, n in rates)
    secs = sum(n / r for r, n in rates)
    value = total / secs
    sample = f"every {stride}th pixel of the 1920x1080 frame x {SPP_PER_STEP} spp per step ({rates[0][1]} samples/step), OpenMP over pixels"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / len(rates), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "Chunky's Java CPU path tracer (the reference's CPU path) needs a JVM + chunky-core, neither available; "
                "this arm times the C port of the reference kernel's algorithm (oracle/) on all host cores",
    }
    # same-box GPU bar: the reference's own OpenCL kernel JIT-compiled by the NVIDIA OpenCL runtime, if present
    try:
        from oracle import clref
        why = clref.available()
        if why is None:
            ref = clref.ClReference(scene, strict=False)
            _, times = ref.render(pass_seeds(SPP_PER_STEP + 4))
            ms = float(np.median(times[4:]))
            line["opencl_reference_same_gpu"] = {"value": WIDTH * HEIGHT / (ms * 1e-3), "unit": UNIT, "ms_per_pass": ms,
                                                 "device": ref.device_name(), "how": "unmodified reference kernel, one launch per pass, cl_event profiling"}
            ref.close()
        else:
            line["opencl_reference_same_gpu"] = {"unavailable": why}
    except Exception as e:          # never let the extra evidence break the contract line
        line["opencl_reference_same_gpu"] = {"unavailable": repr(e)}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernel", type=int, default=0, help="0 auto (4), 1 megakernel, 2 per-warp pool, 3 lane-bound wavefront, 4 CTA-wide wavefront")
    ap.add_argument("--workload", default="config1", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    global WORKLOAD, WIDTH, HEIGHT
    WORKLOAD = WORKLOADS[args.workload]
    if args.workload == "large":
        WIDTH, HEIGHT = 3840, 2160
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from chunkyclplugin_b200 import native
    from chunkyclplugin_b200.javarandom import JavaRandom
    from chunkyclplugin_b200.multigpu import SampleParallelRenderer
    from chunkyclplugin_b200.renderer import (CudaPathTracingRenderer, CudaSceneLoader, DefaultRenderManager, RendererInstance, Scene)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            print(f"bench.py: --gpus {args.gpus} needs torchrun with {args.gpus} ranks", file=sys.stderr)
            return 2
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene = build_scene(args.workload)
    inst = RendererInstance.get(local_rank)
    ctx = inst.context
    loader = CudaSceneLoader(inst)
    chunky_scene = Scene(scene, target_spp=SPP_PER_STEP)
    loader.ensureLoad(chunky_scene)
    ctx.camera_set(scene.projector_type, scene.camera)
    ctx.render_begin(WIDTH, HEIGHT)
    ctx.render_set_params(kernel=args.kernel)
    spr = SampleParallelRenderer(ctx, rank, world)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    rand = JavaRandom(0)

    def step_seeds():
        # weak scaling: every rank renders SPP_PER_STEP passes; rank r takes passes r, r+N, ... of the global window
        s = [rand.next_int() for _ in range(SPP_PER_STEP * world)]
        return s

    def one_step(timed: bool):
        seeds = step_seeds()
        flush.fill_(rank + 1)                         # L2 flush between iterations (outside the timed events)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        n_local = spr.render_window(seeds)            # blocking; device time from CUDA events on the ctx stream
        ms = ctx.last_kernel_ms()
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            spr.reduce_window(n_local)
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        wall = (time.perf_counter() - t0) * 1e3
        return ms, wall

    launches0 = ctx.launch_count()
    for _ in range(args.warmup):
        one_step(False)
    launches_w = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dev_ms, wall_ms = [], []
    for _ in range(args.steps):
        ms, wall = one_step(True)
        dev_ms.append(ms)
        wall_ms.append(wall)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    sampler.join()
    launches = ctx.launch_count() - launches_w + (args.steps if world > 1 else 0)   # + one NCCL reduce per step
    total_ms = torch.tensor([sum(dev_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    samples_per_step = WIDTH * HEIGHT * SPP_PER_STEP * world
    value = samples_per_step * args.steps / (total_ms * 1e-3)

    # ---- first-hit pass (BASELINE config 2): primary-ray Mrays/s on the same scene ----------------------------
    fh_ms = None
    if rank == 0:
        ts = []
        for i in range(6):
            flush.fill_(i)
            torch.cuda.synchronize()
            lib = native.load()
            import ctypes as C
            native.check(lib.ccu_first_hit(ctx._h, 12345 + i, None, None, None, None, None, None, None))
            ts.append(ctx.last_kernel_ms())
        fh_ms = float(np.median(ts[2:]))

    # ---- memory-system denominators of this path: random 32-byte-sector gathers, L2- and HBM-resident --------
    memsys = None
    if rank == 0:
        l2_bw, _ = ctx.bench_gather(4 << 20)
        hbm_bw, _ = ctx.bench_gather(1 << 30)
        _, l2_ns = ctx.bench_gather(4 << 20, dependent=True)
        _, hbm_ns = ctx.bench_gather(1 << 30, dependent=True)
        memsys = {"l2_random_sector_gbs": l2_bw, "hbm_random_sector_gbs": hbm_bw, "l2_dependent_gather_ns": l2_ns,
                  "hbm_dependent_gather_ns": hbm_ns,
                  "how": "ccu_bench_gather: random 16-byte ld.global.cg per 32-byte sector over a 4 MiB / 1 GiB array, all SMs"}

    # ---- end to end through the host renderer (public API): seeds H2D + readback/merge D2H per step ------------
    e2e = None
    if True:
        ctx.render_end()
        renderer = CudaPathTracingRenderer(loader)
        e_samples, e_secs = 0, 0.0
        # one Chunky scene object for all steps, as in Chunky (the double sample buffer is allocated once per scene and
        # stays resident); every step is a fresh render of SPP_PER_STEP passes into it
        cs = Scene(scene, target_spp=SPP_PER_STEP)
        cs.packed = scene                                  # same scene object: no re-upload
        mgr = DefaultRenderManager(cs)
        for i in range(3 + max(3, args.steps // 2)):
            cs.spp = 0
            cs.sample_buffer.fill(0.0)
            flush.fill_(i)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            if world == 1:
                renderer.render(mgr)                       # render_begin, 16 passes, merge into the double buffer, render_end
            else:
                # N GPUs: every rank renders its 16 passes of the 16*N-pass window, one NCCL reduce, rank 0 alone merges the
                # window into the host sample buffer (what a multi-GPU plugin host does; the other ranks hold no host buffer)
                ctx.camera_set(scene.projector_type, scene.camera)
                ctx.render_begin(WIDTH, HEIGHT)
                spr.render_and_merge(step_seeds(), cs.sample_buffer if rank == 0 else None, 0)
                ctx.render_end()
                cs.spp = SPP_PER_STEP
                if rank != 0:
                    cs.sample_buffer[0] = 1.0
            dt = time.perf_counter() - t0
            if i >= 3:
                e_samples += WIDTH * HEIGHT * SPP_PER_STEP
                e_secs += dt
            assert cs.spp == SPP_PER_STEP and cs.sample_buffer.max() > 0
        t = torch.tensor([e_secs], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": e_samples * world / float(t.item()), "unit": UNIT, "h2d_bytes_per_step": 4 * SPP_PER_STEP,
               "d2h_bytes_per_step": 4 * 3 * WIDTH * HEIGHT,
               "how": "CudaPathTracingRenderer.render(): render_begin + 16 passes + ccu_render_merge into the host double buffer" if world == 1 else
                      "per rank: camera + render_begin + 16 passes; NCCL reduce; rank 0: ccu_render_merge of the 16*N-pass window into the host double buffer"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- CPU baseline + algorithmic bytes (bounded sample of the same workload) ---------------------------------
    peaks, peak_src = measured_peaks()
    cpu = None
    bytes_per_sample = None
    if not args.no_cpu_baseline:
        import oracle
        from chunkyclplugin_b200.javarandom import pass_seeds
        stride = 8
        rate, counters, n, cores = cpu_port_rate(scene, pass_seeds(SPP_PER_STEP), stride)
        bytes_per_sample = oracle.algorithmic_bytes(counters) / counters["samples"]
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"every {stride}th pixel of the 1920x1080 frame x {SPP_PER_STEP} spp ({n} samples), oracle C port, OpenMP"}
    else:
        alg = os.path.join(ROOT, "profiles", "alg_bytes.json")
        if os.path.exists(alg):
            bytes_per_sample = json.load(open(alg)).get("terrain256_1080p_bytes_per_sample")
    kernel_ms = float(np.mean(dev_ms))
    roofline = None
    if bytes_per_sample:
        achieved = bytes_per_sample * WIDTH * HEIGHT * SPP_PER_STEP / (kernel_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                    "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_sample": bytes_per_sample,
                    "kernel": "render kernel, one launch per 16-pass window", "kernel_ms": kernel_ms,
                    "note": "bound is nominal: the scene is L1/L2 resident, the kernel is limited by instruction issue and the latency of "
                            "dependent 32-byte-sector gathers (see memory_system and profiles/), not by HBM streaming"}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if roofline and os.path.exists(traffic_file):
        roofline["traffic"] = json.load(open(traffic_file)).get("render_dram_bytes_per_launch")

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "spp_per_step_per_gpu": SPP_PER_STEP, "global_spp_per_step": SPP_PER_STEP * world,
                   "parallelism": f"sample-parallel x{world}" + (" + NCCL reduce per step" if world > 1 else ""),
                   "l2": "flushed between timed steps (256 MB write); the 2 MB scene itself is L2-resident by nature"},
        "clocks": sampler.summary(),
        "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu,
        "primary_rays": {"value": WIDTH * HEIGHT / (fh_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms": fh_ms,
                         "workload": "config2: first-hit pass at 1080p, same scene"} if fh_ms else None,
        "memory_system": memsys,
        "wall_ms_per_step": float(np.mean(wall_ms)),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
