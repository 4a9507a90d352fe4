"""Runs the reference's UNMODIFIED OpenCL kernels on the GPU box.  TEST INFRASTRUCTURE ONLY.

The GPU boxes ship the NVIDIA OpenCL driver (libnvidia-opencl.so.1) but no ICD registration and
no CL headers, so this module (a) points the Khronos loader at the vendor library through
OCL_ICD_FILENAMES and (b) hand-declares the few OpenCL 1.2 entry points through ctypes.  It drives
the kernels the way the reference's Java host does:

* program build:  same options as KernelLoader.java:52 ("-cl-std=CL1.2 -Werror");
* buffers:        READ_ONLY | COPY_HOST_PTR int buffers (ClIntBuffer.java:22-24), RGBA8 UNORM
                  image2d_array atlas (ClTextureLoader.java:46-58), RGBA8 image2d sky (ClSky.java:33-62);
* pass loop:      per pass: write seed, write bufferSpp, bind 20 args, enqueue W*H work-items,
                  wait (OpenClPathTracingRenderer.java:102-144).

The kernel source is oracle/_ref/chunkycl_kernel.cl, produced by oracle/clref/make_ref.py from
/root/reference (git-ignored build output).  ``strict=True`` prepends ``#pragma OPENCL FP_CONTRACT OFF``
and adds -cl-fp32-correctly-rounded-divide-sqrt: the same source under IEEE arithmetic, used to pin the
oracle's restatement bit-for-bit; ``strict=False`` is the stock reference build.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_CL = os.path.join(os.path.dirname(_HERE), "_ref", "chunkycl_kernel.cl")
PROBE_CL = os.path.join(_HERE, "probe_kernels.cl")
TONEMAP_CL = os.path.join(os.path.dirname(_HERE), "_ref", "chunkycl_tonemap.cl")

CL_DEVICE_TYPE_GPU = 1 << 2
CL_MEM_READ_WRITE, CL_MEM_READ_ONLY, CL_MEM_COPY_HOST_PTR = 1 << 0, 1 << 2, 1 << 5
CL_RGBA, CL_UNORM_INT8 = 0x10B5, 0x10D2
CL_MEM_OBJECT_IMAGE2D, CL_MEM_OBJECT_IMAGE2D_ARRAY = 0x10F1, 0x10F3
CL_PROGRAM_BUILD_LOG, CL_PROGRAM_BINARY_SIZES, CL_PROGRAM_BINARIES = 0x1183, 0x1165, 0x1166
CL_QUEUE_PROFILING_ENABLE = 1 << 1
CL_PROFILING_COMMAND_START, CL_PROFILING_COMMAND_END = 0x1282, 0x1283
CL_DEVICE_NAME, CL_DEVICE_VERSION, CL_DRIVER_VERSION = 0x102B, 0x102F, 0x102D

REFERENCE_BUILD_OPTIONS = "-cl-std=CL1.2 -Werror"       # KernelLoader.java:52


class _ImageFormat(C.Structure):
    _fields_ = [("order", C.c_uint32), ("dtype", C.c_uint32)]


class _ImageDesc(C.Structure):
    _fields_ = [("image_type", C.c_uint32), ("width", C.c_size_t), ("height", C.c_size_t), ("depth", C.c_size_t),
                ("array_size", C.c_size_t), ("row_pitch", C.c_size_t), ("slice_pitch", C.c_size_t),
                ("num_mip_levels", C.c_uint32), ("num_samples", C.c_uint32), ("buffer", C.c_void_p)]


class ClError(RuntimeError):
    pass


_cl = None


def _load():
    global _cl
    if _cl is not None:
        return _cl
    if "OCL_ICD_FILENAMES" not in os.environ and not os.path.isdir("/etc/OpenCL/vendors"):
        os.environ["OCL_ICD_FILENAMES"] = "libnvidia-opencl.so.1"
    cl = C.CDLL("libOpenCL.so.1")
    vp = C.c_void_p
    cl.clCreateContext.restype = vp
    cl.clCreateCommandQueue.restype = vp
    cl.clCreateBuffer.restype = vp
    cl.clCreateImage.restype = vp
    cl.clCreateProgramWithSource.restype = vp
    cl.clCreateKernel.restype = vp
    cl.clCreateContext.argtypes = [vp, C.c_uint32, C.POINTER(vp), vp, vp, C.POINTER(C.c_int32)]
    cl.clCreateCommandQueue.argtypes = [vp, vp, C.c_uint64, C.POINTER(C.c_int32)]
    cl.clCreateBuffer.argtypes = [vp, C.c_uint64, C.c_size_t, vp, C.POINTER(C.c_int32)]
    cl.clCreateImage.argtypes = [vp, C.c_uint64, C.POINTER(_ImageFormat), C.POINTER(_ImageDesc), vp, C.POINTER(C.c_int32)]
    cl.clCreateProgramWithSource.argtypes = [vp, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int32)]
    cl.clBuildProgram.argtypes = [vp, C.c_uint32, C.POINTER(vp), C.c_char_p, vp, vp]
    cl.clGetProgramBuildInfo.argtypes = [vp, vp, C.c_uint32, C.c_size_t, vp, C.POINTER(C.c_size_t)]
    cl.clGetProgramInfo.argtypes = [vp, C.c_uint32, C.c_size_t, vp, C.POINTER(C.c_size_t)]
    cl.clCreateKernel.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int32)]
    cl.clSetKernelArg.argtypes = [vp, C.c_uint32, C.c_size_t, vp]
    cl.clEnqueueNDRangeKernel.argtypes = [vp, vp, C.c_uint32, vp, C.POINTER(C.c_size_t), vp, C.c_uint32, vp, C.POINTER(vp)]
    cl.clEnqueueReadBuffer.argtypes = [vp, vp, C.c_uint32, C.c_size_t, C.c_size_t, vp, C.c_uint32, vp, vp]
    cl.clEnqueueWriteBuffer.argtypes = [vp, vp, C.c_uint32, C.c_size_t, C.c_size_t, vp, C.c_uint32, vp, vp]
    cl.clGetEventProfilingInfo.argtypes = [vp, C.c_uint32, C.c_size_t, vp, vp]
    cl.clWaitForEvents.argtypes = [C.c_uint32, C.POINTER(vp)]
    cl.clReleaseEvent.argtypes = [vp]
    cl.clReleaseMemObject.argtypes = [vp]
    cl.clReleaseKernel.argtypes = [vp]
    cl.clFinish.argtypes = [vp]
    cl.clGetDeviceIDs.argtypes = [vp, C.c_uint64, C.c_uint32, C.POINTER(vp), C.POINTER(C.c_uint32)]
    cl.clGetPlatformIDs.argtypes = [C.c_uint32, C.POINTER(vp), C.POINTER(C.c_uint32)]
    cl.clGetDeviceInfo.argtypes = [vp, C.c_uint32, C.c_size_t, vp, C.POINTER(C.c_size_t)]
    _cl = cl
    return cl


def _chk(err: int, what: str):
    if err != 0:
        raise ClError(f"{what} failed with OpenCL error {err}")


def available() -> Optional[str]:
    """None if the reference kernel can run here, else the reason it cannot."""
    if not os.path.exists(REF_CL):
        return f"{REF_CL} missing (run oracle/clref/make_ref.py where /root/reference exists)"
    try:
        cl = _load()
    except OSError as e:
        return f"libOpenCL.so.1 not loadable: {e}"
    n = C.c_uint32()
    err = cl.clGetPlatformIDs(0, None, C.byref(n))
    if err != 0 or n.value == 0:
        return f"no OpenCL platform (clGetPlatformIDs -> {err}, {n.value})"
    plats = (C.c_void_p * n.value)()
    cl.clGetPlatformIDs(n.value, plats, None)
    nd = C.c_uint32()
    err = cl.clGetDeviceIDs(plats[0], CL_DEVICE_TYPE_GPU, 0, None, C.byref(nd))
    if err != 0 or nd.value == 0:
        return f"no OpenCL GPU device (clGetDeviceIDs -> {err})"
    return None


class ClReference:
    """One OpenCL context + the reference program, with a scene uploaded the reference's way."""

    def __init__(self, scene, strict: bool = False, device_index: int = 0, extra_options: str = ""):
        cl = self.cl = _load()
        n = C.c_uint32()
        _chk(cl.clGetPlatformIDs(0, None, C.byref(n)), "clGetPlatformIDs")
        plats = (C.c_void_p * n.value)()
        cl.clGetPlatformIDs(n.value, plats, None)
        nd = C.c_uint32()
        _chk(cl.clGetDeviceIDs(plats[0], CL_DEVICE_TYPE_GPU, 0, None, C.byref(nd)), "clGetDeviceIDs")
        devs = (C.c_void_p * nd.value)()
        cl.clGetDeviceIDs(plats[0], CL_DEVICE_TYPE_GPU, nd.value, devs, None)
        self.device = C.c_void_p(devs[device_index])
        err = C.c_int32()
        dev_arr = (C.c_void_p * 1)(self.device)
        self.ctx = cl.clCreateContext(None, 1, dev_arr, None, None, C.byref(err))
        _chk(err.value, "clCreateContext")
        self.queue = cl.clCreateCommandQueue(self.ctx, self.device, CL_QUEUE_PROFILING_ENABLE, C.byref(err))
        _chk(err.value, "clCreateCommandQueue")
        self.strict = strict
        src = open(REF_CL).read() + "\n" + open(PROBE_CL).read()
        opts = REFERENCE_BUILD_OPTIONS
        if strict:
            src = "#pragma OPENCL FP_CONTRACT OFF\n" + src
            opts += " -cl-fp32-correctly-rounded-divide-sqrt"
        if extra_options:
            opts += " " + extra_options
        self.options = opts
        sb = src.encode()
        strs = (C.c_char_p * 1)(sb)
        lens = (C.c_size_t * 1)(len(sb))
        self.program = cl.clCreateProgramWithSource(self.ctx, 1, strs, lens, C.byref(err))
        _chk(err.value, "clCreateProgramWithSource")
        berr = cl.clBuildProgram(self.program, 1, dev_arr, opts.encode(), None, None)
        if berr != 0:
            raise ClError(f"clBuildProgram failed ({berr}):\n{self.build_log()}")
        self._mem = []
        self.scene = scene
        self._upload(scene)

    # -- helpers ---------------------------------------------------------------------------
    def device_name(self) -> str:
        buf = C.create_string_buffer(256)
        self.cl.clGetDeviceInfo(self.device, CL_DEVICE_NAME, 256, buf, None)
        drv = C.create_string_buffer(256)
        self.cl.clGetDeviceInfo(self.device, CL_DRIVER_VERSION, 256, drv, None)
        return f"{buf.value.decode()} (OpenCL driver {drv.value.decode()})"

    def build_log(self) -> str:
        sz = C.c_size_t()
        self.cl.clGetProgramBuildInfo(self.program, self.device, CL_PROGRAM_BUILD_LOG, 0, None, C.byref(sz))
        buf = C.create_string_buffer(sz.value + 1)
        self.cl.clGetProgramBuildInfo(self.program, self.device, CL_PROGRAM_BUILD_LOG, sz.value, buf, None)
        return buf.value.decode(errors="replace")

    def binary(self) -> bytes:
        """CL_PROGRAM_BINARIES - PTX text on NVIDIA; shows the JIT's fma/div/sqrt policy."""
        sz = C.c_size_t()
        self.cl.clGetProgramInfo(self.program, CL_PROGRAM_BINARY_SIZES, C.sizeof(C.c_size_t), C.byref(sz), None)
        buf = C.create_string_buffer(sz.value)
        ptrs = (C.c_void_p * 1)(C.addressof(buf))
        self.cl.clGetProgramInfo(self.program, CL_PROGRAM_BINARIES, C.sizeof(C.c_void_p), ptrs, None)
        return buf.raw

    def _buffer(self, arr: np.ndarray, flags=CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR):
        a = np.ascontiguousarray(arr)
        if a.size == 0:                                   # ClIntBuffer.java:15-18
            a = np.zeros(1, dtype=a.dtype)
        err = C.c_int32()
        m = self.cl.clCreateBuffer(self.ctx, flags, a.nbytes, a.ctypes.data_as(C.c_void_p), C.byref(err))
        _chk(err.value, "clCreateBuffer")
        self._mem.append(m)
        return C.c_void_p(m)

    def _image(self, rgba: np.ndarray, array: bool):
        a = np.ascontiguousarray(rgba, dtype=np.uint8)
        fmt = _ImageFormat(CL_RGBA, CL_UNORM_INT8)
        desc = _ImageDesc()
        if array:
            desc.image_type = CL_MEM_OBJECT_IMAGE2D_ARRAY
            desc.array_size, desc.height, desc.width = a.shape[0], a.shape[1], a.shape[2]
        else:
            desc.image_type = CL_MEM_OBJECT_IMAGE2D
            desc.height, desc.width = a.shape[0], a.shape[1]
        err = C.c_int32()
        m = self.cl.clCreateImage(self.ctx, CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR, C.byref(fmt), C.byref(desc),
                                  a.ctypes.data_as(C.c_void_p), C.byref(err))
        _chk(err.value, "clCreateImage")
        self._mem.append(m)
        return C.c_void_p(m)

    def _upload(self, s):
        i32 = lambda v: np.array([v], dtype=np.int32)
        self.b = dict(
            projectorType=self._buffer(i32(s.projector_type)),
            cameraSettings=self._buffer(s.camera.astype(np.float32)),
            octreeDepth=self._buffer(i32(s.octree_depth)),
            octreeData=self._buffer(s.octree.astype(np.int32)),
            bPalette=self._buffer(s.block_palette), quadModels=self._buffer(s.quad_models),
            aabbModels=self._buffer(s.aabb_models), worldBvhData=self._buffer(s.world_bvh),
            actorBvhData=self._buffer(s.actor_bvh), bvhTrigs=self._buffer(s.bvh_trigs),
            textureAtlas=self._image(s.atlas, array=True), matPalette=self._buffer(s.mat_palette),
            skyTexture=self._image(s.sky, array=False),
            skyIntensity=self._buffer(np.array([s.sky_intensity], dtype=np.float32)),
            sunData=self._buffer(s.sun),
            randomSeed=self._buffer(i32(0), CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR),
            bufferSpp=self._buffer(i32(0), CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR),
            width=self._buffer(i32(s.width)), height=self._buffer(i32(s.height)),
        )

    def _kernel(self, name: str):
        err = C.c_int32()
        k = self.cl.clCreateKernel(self.program, name.encode(), C.byref(err))
        _chk(err.value, f"clCreateKernel({name})")
        return C.c_void_p(k)

    def _set_args(self, k, names):
        for i, n in enumerate(names):
            m = n if isinstance(n, C.c_void_p) else self.b[n]
            _chk(self.cl.clSetKernelArg(k, i, C.sizeof(C.c_void_p), C.byref(m)), f"clSetKernelArg {i}")

    def _launch(self, k, n: int) -> float:
        """Enqueue n work-items, wait; returns device time in ms (event profiling)."""
        ev = C.c_void_p()
        g = (C.c_size_t * 1)(n)
        _chk(self.cl.clEnqueueNDRangeKernel(self.queue, k, 1, None, g, None, 0, None, C.byref(ev)), "clEnqueueNDRangeKernel")
        evs = (C.c_void_p * 1)(ev)
        _chk(self.cl.clWaitForEvents(1, evs), "clWaitForEvents")
        t0, t1 = C.c_uint64(), C.c_uint64()
        self.cl.clGetEventProfilingInfo(ev, CL_PROFILING_COMMAND_START, 8, C.byref(t0), None)
        self.cl.clGetEventProfilingInfo(ev, CL_PROFILING_COMMAND_END, 8, C.byref(t1), None)
        self.cl.clReleaseEvent(ev)
        return (t1.value - t0.value) * 1e-6

    def _write_i32(self, mem, v: int):
        a = np.array([v], dtype=np.int32)
        _chk(self.cl.clEnqueueWriteBuffer(self.queue, mem, 1, 0, 4, a.ctypes.data_as(C.c_void_p), 0, None, None), "clEnqueueWriteBuffer")

    def _read(self, mem, arr: np.ndarray):
        _chk(self.cl.clEnqueueReadBuffer(self.queue, mem, 1, 0, arr.nbytes, arr.ctypes.data_as(C.c_void_p), 0, None, None), "clEnqueueReadBuffer")

    # -- the reference's render loop ---------------------------------------------------------
    RENDER_ARGS = ("projectorType", "cameraSettings", "octreeDepth", "octreeData", "bPalette", "quadModels",
                   "aabbModels", "worldBvhData", "actorBvhData", "bvhTrigs", "textureAtlas", "matPalette",
                   "skyTexture", "skyIntensity", "sunData", "randomSeed", "bufferSpp", "width", "height")

    def render(self, seeds: Sequence[int], start_spp: int = 0, res: Optional[np.ndarray] = None):
        """One launch per pass exactly as OpenClPathTracingRenderer.java:102-144.  Returns (mean RGB, kernel ms list)."""
        s = self.scene
        n = s.width * s.height
        if res is None:
            res = np.zeros(n * 3, dtype=np.float32)
        buf = self._buffer(res, CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR)
        k = self._kernel("render")
        times = []
        for p, seed in enumerate(seeds):
            self._write_i32(self.b["randomSeed"], int(seed))
            self._write_i32(self.b["bufferSpp"], start_spp + p)
            self._set_args(k, list(self.RENDER_ARGS) + [buf])
            times.append(self._launch(k, n))
        self._read(buf, res)
        self.cl.clReleaseKernel(k)
        self.cl.clReleaseMemObject(buf)
        self._mem.remove(buf.value)
        return res, times

    def preview(self):
        s = self.scene
        n = s.width * s.height
        res = np.zeros(n, dtype=np.int32)
        buf = self._buffer(res, CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR)
        k = self._kernel("preview")
        names = [a for a in self.RENDER_ARGS if a not in ("randomSeed", "bufferSpp")] + [buf]
        self._set_args(k, names)
        ms = self._launch(k, n)
        self._read(buf, res)
        self.cl.clReleaseKernel(k)
        return res, ms

    def first_hit(self, seed: int) -> Dict[str, np.ndarray]:
        s = self.scene
        n = s.width * s.height
        out = dict(hit=np.zeros(n, np.int32), block=np.zeros(n, np.int32), t=np.zeros(n, np.float32),
                   normal=np.zeros(n * 3, np.float32), color=np.zeros(n * 4, np.float32), ray=np.zeros(n * 6, np.float32))
        bufs = {k_: self._buffer(v, CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR) for k_, v in out.items()}
        k = self._kernel("probe_first_hit")
        self._write_i32(self.b["randomSeed"], int(seed))
        names = ["projectorType", "cameraSettings", "octreeDepth", "octreeData", "bPalette", "quadModels", "aabbModels",
                 "worldBvhData", "actorBvhData", "bvhTrigs", "textureAtlas", "matPalette", "randomSeed", "width", "height",
                 bufs["hit"], bufs["block"], bufs["t"], bufs["normal"], bufs["color"], bufs["ray"]]
        self._set_args(k, names)
        out["ms"] = self._launch(k, n)
        for k_ in bufs:
            self._read(bufs[k_], out[k_])
        self.cl.clReleaseKernel(k)
        return out

    def math(self, fn: int, x: np.ndarray, y: Optional[np.ndarray] = None) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(y if y is not None else np.zeros_like(x), np.float32)
        out = np.zeros_like(x)
        bx, by = self._buffer(x), self._buffer(y)
        bo = self._buffer(out, CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR)
        k = self._kernel("probe_math")
        for i, m in enumerate((bx, by, bo)):
            _chk(self.cl.clSetKernelArg(k, i, C.sizeof(C.c_void_p), C.byref(m)), "clSetKernelArg")
        f = C.c_int32(fn)
        _chk(self.cl.clSetKernelArg(k, 3, 4, C.byref(f)), "clSetKernelArg")
        self._launch(k, x.size)
        self._read(bo, out)
        self.cl.clReleaseKernel(k)
        return out

    def close(self):
        for m in self._mem:
            self.cl.clReleaseMemObject(m)
        self._mem = []


class ClTonemap:
    """The reference's tonemap program (tonemap/include/post_processing_filter.cl), built and driven the way
    GpuPostProcessingFilter.java:35-65 does: upload the double sample buffer, 6 scalar / buffer args, W*H work-items,
    read the ARGB image back."""

    def __init__(self, strict: bool = False, device_index: int = 0):
        cl = self.cl = _load()
        n = C.c_uint32()
        _chk(cl.clGetPlatformIDs(0, None, C.byref(n)), "clGetPlatformIDs")
        plats = (C.c_void_p * n.value)()
        cl.clGetPlatformIDs(n.value, plats, None)
        nd = C.c_uint32()
        _chk(cl.clGetDeviceIDs(plats[0], CL_DEVICE_TYPE_GPU, 0, None, C.byref(nd)), "clGetDeviceIDs")
        devs = (C.c_void_p * nd.value)()
        cl.clGetDeviceIDs(plats[0], CL_DEVICE_TYPE_GPU, nd.value, devs, None)
        self.device = C.c_void_p(devs[device_index])
        err = C.c_int32()
        dev_arr = (C.c_void_p * 1)(self.device)
        self.ctx = cl.clCreateContext(None, 1, dev_arr, None, None, C.byref(err))
        _chk(err.value, "clCreateContext")
        self.queue = cl.clCreateCommandQueue(self.ctx, self.device, CL_QUEUE_PROFILING_ENABLE, C.byref(err))
        _chk(err.value, "clCreateCommandQueue")
        src = open(TONEMAP_CL).read()
        opts = REFERENCE_BUILD_OPTIONS
        if strict:
            src = "#pragma OPENCL FP_CONTRACT OFF\n" + src
            opts += " -cl-fp32-correctly-rounded-divide-sqrt"
        sb = src.encode()
        strs = (C.c_char_p * 1)(sb)
        lens = (C.c_size_t * 1)(len(sb))
        self.program = cl.clCreateProgramWithSource(self.ctx, 1, strs, lens, C.byref(err))
        _chk(err.value, "clCreateProgramWithSource")
        berr = cl.clBuildProgram(self.program, 1, dev_arr, opts.encode(), None, None)
        if berr != 0:
            sz = C.c_size_t()
            cl.clGetProgramBuildInfo(self.program, self.device, CL_PROGRAM_BUILD_LOG, 0, None, C.byref(sz))
            buf = C.create_string_buffer(sz.value + 1)
            cl.clGetProgramBuildInfo(self.program, self.device, CL_PROGRAM_BUILD_LOG, sz.value, buf, None)
            raise ClError(f"clBuildProgram(tonemap) failed ({berr}):\n{buf.value.decode(errors='replace')}")

    def filter(self, width: int, height: int, exposure: float, sample_buffer: np.ndarray, filter_type: int):
        cl = self.cl
        inp = np.ascontiguousarray(sample_buffer, dtype=np.float64).reshape(-1)
        n = width * height
        assert inp.size == 3 * n
        out = np.zeros(n, dtype=np.int32)
        err = C.c_int32()
        m_in = C.c_void_p(cl.clCreateBuffer(self.ctx, CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR, inp.nbytes, inp.ctypes.data_as(C.c_void_p), C.byref(err)))
        _chk(err.value, "clCreateBuffer(input)")
        m_out = C.c_void_p(cl.clCreateBuffer(self.ctx, CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR, out.nbytes, out.ctypes.data_as(C.c_void_p), C.byref(err)))
        _chk(err.value, "clCreateBuffer(output)")
        k = C.c_void_p(cl.clCreateKernel(self.program, b"filter", C.byref(err)))
        _chk(err.value, "clCreateKernel(filter)")
        w, h, e, t = C.c_int32(width), C.c_int32(height), C.c_float(exposure), C.c_int32(filter_type)
        _chk(cl.clSetKernelArg(k, 0, 4, C.byref(w)), "arg0")
        _chk(cl.clSetKernelArg(k, 1, 4, C.byref(h)), "arg1")
        _chk(cl.clSetKernelArg(k, 2, 4, C.byref(e)), "arg2")
        _chk(cl.clSetKernelArg(k, 3, C.sizeof(C.c_void_p), C.byref(m_in)), "arg3")
        _chk(cl.clSetKernelArg(k, 4, C.sizeof(C.c_void_p), C.byref(m_out)), "arg4")
        _chk(cl.clSetKernelArg(k, 5, 4, C.byref(t)), "arg5")
        ev = C.c_void_p()
        g = (C.c_size_t * 1)(n)
        _chk(cl.clEnqueueNDRangeKernel(self.queue, k, 1, None, g, None, 0, None, C.byref(ev)), "clEnqueueNDRangeKernel")
        evs = (C.c_void_p * 1)(ev)
        _chk(cl.clWaitForEvents(1, evs), "clWaitForEvents")
        t0, t1 = C.c_uint64(), C.c_uint64()
        cl.clGetEventProfilingInfo(ev, CL_PROFILING_COMMAND_START, 8, C.byref(t0), None)
        cl.clGetEventProfilingInfo(ev, CL_PROFILING_COMMAND_END, 8, C.byref(t1), None)
        cl.clReleaseEvent(ev)
        _chk(cl.clEnqueueReadBuffer(self.queue, m_out, 1, 0, out.nbytes, out.ctypes.data_as(C.c_void_p), 0, None, None), "clEnqueueReadBuffer")
        cl.clReleaseKernel(k)
        cl.clReleaseMemObject(m_in)
        cl.clReleaseMemObject(m_out)
        return out, (t1.value - t0.value) * 1e-6
