"""The C-ABI shared library: builds, loads, exports every symbol the header declares, and fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "chunkycu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ccu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from chunkyclplugin_b200 import build, native
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/chunkycu.h but not exported"
    # the Python binding covers the same set, no more, no less
    assert sorted(native.SYMBOLS) == names


def test_version_and_error_string():
    from chunkyclplugin_b200 import native
    lib = native.load()
    assert b"sm_100a" in lib.ccu_version()
    assert isinstance(lib.ccu_last_error(), bytes)


def test_no_gpu_means_loud_failure_not_fallback():
    """Without a CUDA device every entry point reports CCU_ENODEVICE; nothing routes to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path is exercised on the CPU box")
    from chunkyclplugin_b200 import native
    with pytest.raises(native.ChunkyCuError) as e:
        native.device_count()
    assert e.value.code == native.CCU_ENODEVICE
    with pytest.raises(native.ChunkyCuError) as e:
        native.Context(0)
    assert e.value.code == native.CCU_ENODEVICE
    from chunkyclplugin_b200.renderer import RendererInstance
    with pytest.raises(native.ChunkyCuError):
        RendererInstance.get()


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under chunkyclplugin_b200/ or include/ may reference it."""
    bad = []
    for base in ("chunkyclplugin_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    src = open(os.path.join(dirpath, f), errors="replace").read()
                    if re.search(r"^\s*(import|from)\s+oracle\b|#include\s+\"[^\"]*oracle|liboracle", src, flags=re.M):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
