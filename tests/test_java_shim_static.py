"""The Java FFM shim cannot be compiled here (no JDK), so it is checked statically: every downcall handle in
java/.../ChunkyCu.java must name a function include/chunkycu.h declares, with the same arity and argument kinds
(pointer / int / long / float).  A mismatch here would be a crash at the first call inside Chunky."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JAVA = os.path.join(ROOT, "java", "dev", "thatredox", "chunkynative", "cuda")


def c_prototypes():
    text = open(os.path.join(ROOT, "include", "chunkycu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\b(int|const char \*)\s*(ccu_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        kinds = []
        args = " ".join(args.split())
        if args not in ("", "void"):
            for a in args.split(","):
                a = a.strip()
                if "*" in a or "[" in a:
                    kinds.append("ADDRESS")
                elif re.match(r"(const )?(int64_t|uint64_t|size_t)\b", a):
                    kinds.append("JAVA_LONG")
                elif re.match(r"(const )?float\b", a):
                    kinds.append("JAVA_FLOAT")
                elif re.match(r"(const )?double\b", a):
                    kinds.append("JAVA_DOUBLE")
                elif re.match(r"(const )?(int|int32_t|uint32_t)\b", a):
                    kinds.append("JAVA_INT")
                else:
                    raise AssertionError(f"unclassified C parameter {a!r} of {name}")
        protos[name] = ("ADDRESS" if "char" in ret else "JAVA_INT", kinds)
    return protos


def java_descriptors():
    src = open(os.path.join(JAVA, "ChunkyCu.java")).read()
    named = dict(re.findall(r"FunctionDescriptor\s+(\w+)\s*=\s*FunctionDescriptor\.of\(([^)]*)\)", src))
    out = {}
    # h("name", FunctionDescriptor.of(...)) / h("name", WORDS) / LINKER.downcallHandle(LIB.find("name")..., FunctionDescriptor.of(...), ...)
    for name, fd in re.findall(r'h\("(ccu_\w+)",\s*(FunctionDescriptor\.of\([^)]*\)|\w+)\)', src):
        out[name] = fd
    for name, fd in re.findall(r'LIB\.find\("(ccu_\w+)"\)[^;]*?(FunctionDescriptor\.of\([^)]*\))', src, flags=re.S):
        out[name] = fd
    parsed = {}
    for name, fd in out.items():
        m = re.match(r"FunctionDescriptor\.of\(([^)]*)\)", fd)
        body = m.group(1) if m else named[fd]
        parts = [p.strip() for p in body.split(",") if p.strip()]
        parsed[name] = (parts[0], parts[1:])
    return parsed


def test_every_java_downcall_matches_the_c_header():
    c, j = c_prototypes(), java_descriptors()
    assert len(j) >= 40, sorted(j)
    for name, (ret, args) in j.items():
        assert name in c, f"ChunkyCu.java binds {name}, which include/chunkycu.h does not declare"
        assert (ret, args) == c[name], f"{name}: Java {ret}({', '.join(args)}) vs C {c[name][0]}({', '.join(c[name][1])})"


def test_java_binds_the_render_path_entry_points():
    j = java_descriptors()
    needed = ["ccu_ctx_create", "ccu_ctx_destroy", "ccu_scene_begin", "ccu_scene_commit", "ccu_scene_set_octree", "ccu_camera_set",
              "ccu_render_begin", "ccu_render_set_params", "ccu_render_passes", "ccu_render_window_close", "ccu_render_window_merge",
              "ccu_render_end", "ccu_preview", "ccu_tonemap", "ccu_group_create", "ccu_group_replicate_scene", "ccu_group_render_passes",
              "ccu_group_render_merge"]
    for n in needed:
        assert n in j, n


def test_render_params_struct_layout_matches():
    """ccu_render_params {int32 draw_depth, int32 max_depth, float emitter_scale, int32 kernel, int32 flags}: same field order
    in the C header, the ctypes binding and the Java StructLayout."""
    hdr = open(os.path.join(ROOT, "include", "chunkycu.h")).read()
    m = re.search(r"typedef struct ccu_render_params \{(.*?)\} ccu_render_params;", hdr, flags=re.S)
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    c_fields = re.findall(r"\b(int32_t|float)\s+(\w+)\s*;", body)
    from chunkyclplugin_b200 import native
    assert [n for _, n in c_fields] == [n for n, _ in native.RenderParams._fields_]
    java = open(os.path.join(JAVA, "ChunkyCu.java")).read()
    j_fields = re.findall(r'(JAVA_INT|JAVA_FLOAT)\.withName\("(\w+)"\)', java)
    assert [(("JAVA_FLOAT" if t == "float" else "JAVA_INT"), n) for t, n in c_fields] == j_fields


def test_renderer_ids_are_the_references():
    """The renderer selector ids must not change (OpenClPathTracingRenderer.java:33-45, OpenClPreviewRenderer.java:27-29)."""
    assert '"ChunkyClRenderer"' in open(os.path.join(JAVA, "CudaPathTracingRenderer.java")).read()
    assert '"ChunkyClPreviewRenderer"' in open(os.path.join(JAVA, "CudaPreviewRenderer.java")).read()
    from chunkyclplugin_b200.renderer import CudaPathTracingRenderer, CudaPreviewRenderer
    assert CudaPathTracingRenderer.getId(None) == "ChunkyClRenderer" and CudaPreviewRenderer.getId(None) == "ChunkyClPreviewRenderer"
