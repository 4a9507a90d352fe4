#!/usr/bin/env python3
"""Recipe: build the reference's OpenCL translation unit from the sources where they lie.

Runs the C preprocessor over /root/reference/src/main/opencl/kernel/include/rayTracer.cl (resolving
its #include "..." lines exactly as KernelLoader.java:17-68 does with header programs; opencl.h is the
IDE-only stub that the loader replaces by an empty header, KernelLoader.java:26) and writes the single
preprocessed translation unit to oracle/_ref/chunkycl_kernel.cl.  That file is a build output: it is
git-ignored (never committed) but travels to the GPU box, where the NVIDIA OpenCL runtime JIT-compiles
it with the reference's own flags.  Nothing is copied into the tracked tree.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CHUNKYCL_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "src/main/opencl/kernel/include/rayTracer.cl")
TONE = os.path.join(REF, "src/main/opencl/tonemap/include/post_processing_filter.cl")
OUT_DIR = os.path.join(os.path.dirname(HERE), "_ref")


def preprocess(src: str, out: str) -> None:
    # -undef/-nostdinc: no host macros or headers leak in; -P: no line markers; -C is NOT used (comments dropped)
    cmd = ["gcc", "-E", "-P", "-undef", "-nostdinc", "-x", "c", "-w", src]
    text = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
    with open(out, "w") as f:
        f.write(text)


def main() -> int:
    if not os.path.exists(SRC):
        print(f"reference sources not found at {SRC}; keeping any existing oracle/_ref", file=sys.stderr)
        return 0
    os.makedirs(OUT_DIR, exist_ok=True)
    preprocess(SRC, os.path.join(OUT_DIR, "chunkycl_kernel.cl"))
    if os.path.exists(TONE):
        preprocess(TONE, os.path.join(OUT_DIR, "chunkycl_tonemap.cl"))
    # the reference's benchmark scene (benchmark/OpenCL_test/), loaded where it lies and packed into the arrays the C ABI
    # takes: a build output like the .cl files (git-ignored, travels to the GPU box) for scripts/run_render.py --scene benchmark
    bench = os.path.join(REF, "benchmark", "OpenCL_test")
    if os.path.exists(os.path.join(bench, "OpenCL_test.octree2")):
        sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
        import numpy as np
        from chunkyclplugin_b200 import octree2
        sc = octree2.load_scene(os.path.join(bench, "OpenCL_test.octree2"), os.path.join(bench, "OpenCL_test.json"), 1920, 1080)
        np.savez_compressed(os.path.join(OUT_DIR, "benchmark_scene.npz"), octree=sc.octree, octree_depth=sc.octree_depth,
                            block_palette=sc.block_palette, mat_palette=sc.mat_palette, camera=sc.camera)
    print("wrote", OUT_DIR)
    return 0


if __name__ == "__main__":
    sys.exit(main())
