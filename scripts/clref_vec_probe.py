#!/usr/bin/env python3
"""Dump how the NVIDIA OpenCL runtime expands dot/cross/normalize (PTX + numeric samples)."""
import ctypes as C, os, sys, re
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import clref
from chunkyclplugin_b200 import scenes as S
OUT = os.path.join(ROOT, "gpurun_out"); os.makedirs(OUT, exist_ok=True)
p = S.terrain_scene(64, 32, 18)
rng = np.random.default_rng(5)
n = 4096
a = (rng.normal(size=(n, 3)) * np.exp(rng.normal(size=(n, 1)) * 3)).astype(np.float32)
b = (rng.normal(size=(n, 3)) * np.exp(rng.normal(size=(n, 1)) * 3)).astype(np.float32)
for strict in (True, False):
    ref = clref.ClReference(p, strict=strict)
    out = np.zeros(n * 9, np.float32)
    ba, bb = ref._buffer(a), ref._buffer(b)
    bo = ref._buffer(out, clref.CL_MEM_READ_WRITE | clref.CL_MEM_COPY_HOST_PTR)
    k = ref._kernel("probe_vec")
    for i, m in enumerate((ba, bb, bo)):
        ref.cl.clSetKernelArg(k, i, 8, C.byref(m))
    ref._launch(k, n)
    ref._read(bo, out)
    tag = "strict" if strict else "stock"
    np.savez(os.path.join(OUT, f"clref_vec_{tag}.npz"), a=a, b=b, out=out.reshape(n, 9))
    ptx = ref.binary().decode(errors="replace")
    i0 = ptx.index(".entry probe_vec(")
    open(os.path.join(OUT, f"clref_vec_{tag}.ptx"), "w").write(ptx[i0:])
    ref.close()
print("done")
