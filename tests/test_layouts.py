"""Commit-time traversal layouts (top table + 64-ary nodes, and the air layout with 128-bit leaf maps) against the
reference's root descent (octree.h:81-88) - host-only, no GPU needed."""
import numpy as np
import pytest

from chunkyclplugin_b200 import native


def root_descent(tree, depth, xyz):
    """octree.h:81-88 restated with numpy: leaf word and leaf level for every voxel."""
    x, y, z = xyz[:, 0].astype(np.int64), xyz[:, 1].astype(np.int64), xyz[:, 2].astype(np.int64)
    data = np.full(x.shape, tree[0], dtype=np.int64)
    level = np.full(x.shape, depth, dtype=np.int64)
    for _ in range(depth):
        go = data > 0
        if not go.any():
            break
        lvl = level - 1
        idx = data + ((((x >> lvl) & 1) << 2) | (((y >> lvl) & 1) << 1) | ((z >> lvl) & 1))
        data = np.where(go, tree[np.where(go, idx, 0)], data)
        level = np.where(go, lvl, level)
    assert (data <= 0).all()
    return -data, level


@pytest.mark.parametrize("name", ["tiny8", "terrain64", "terrain128", "decorated", "mixed", "indoor", "large512"])
def test_layouts_answer_like_the_root_descent(name, scenes):
    p = scenes(name)
    tree, depth = np.asarray(p.octree, dtype=np.int32), p.octree_depth
    rng = np.random.default_rng(5)
    edge = 1 << depth
    if edge <= 32:
        g = np.arange(edge)
        xyz = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    else:
        xyz = rng.integers(0, edge, size=(200000, 3))
        # plus voxels next to the surface, where the small leaves are
        xyz[:50000, 1] = rng.integers(max(0, edge // 4 - 24), min(edge, edge // 4 + 64), size=50000)
    want_value, want_level = root_descent(tree, depth, xyz)
    got = native.layout_lookup(tree, depth, xyz)
    assert np.array_equal(got["wide_value"], want_value)
    assert np.array_equal(got["wide_level"], want_level)
    air = want_value == 0
    assert np.array_equal(got["air_solid"] == 0, air)
    assert np.array_equal(got["air_level"][air], want_level[air])
    assert (got["air_level"][~air] == -1).all()
    assert air.any() and (~air).any()


def test_layout_lookup_rejects_bad_input():
    tree = np.zeros(1, dtype=np.int32)
    with pytest.raises(native.ChunkyCuError):
        native.layout_lookup(tree, 3, np.array([[8, 0, 0]]))     # outside the cube
    got = native.layout_lookup(tree, 3, np.array([[7, 7, 7]]))   # a single air leaf at the root
    assert got["air_solid"][0] == 0 and got["air_level"][0] == 3 and got["wide_level"][0] == 3


# ------------------------------------------------------------------------------------------------------
# BVH stage layout (csrc/ccu_layout.h: bvh_ref / TriRepack) against the reference's own arrays (bvh.h:47-110)
# ------------------------------------------------------------------------------------------------------
def _walk_reference(bvh, trigs):
    """Pre-order list of the reference arrays: ('inner', left box, right box) / ('leaf', [20-word triangles])."""
    out, stack = [], [0]
    while stack:
        node = stack.pop()
        head = int(bvh[node])
        if head <= 0:
            prim = -head
            count = int(trigs[prim])
            out.append(("leaf", [tuple(trigs[prim + 1 + 20 * i: prim + 21 + 20 * i]) for i in range(count)]))
        else:
            left, right = node + 7, head
            out.append(("inner", tuple(bvh[left + 1: left + 7]), tuple(bvh[right + 1: right + 7])))
            stack.append(right)
            stack.append(left)          # first child first
    return out


def _walk_layout(rec, tris, root):
    out, stack = [], [root]
    while stack:
        ref = stack.pop()
        if ref < 0:
            off = (-(ref + 1)) * 8
            count = int(tris[off])
            assert not tris[off + 1: off + 8].any()                                   # header padding
            tr = []
            for i in range(count):
                t = tris[off + 8 + 24 * i: off + 8 + 24 * (i + 1)]
                assert not t[20:].any()                                               # triangle padding
                tr.append(tuple(t[:20]))
            out.append(("leaf", tr))
        else:
            r = rec[ref]
            assert r[7] == 0 and r[15] == 0
            out.append(("inner", tuple(r[0:6]), tuple(r[8:14])))
            stack.append(int(r[14]))
            stack.append(int(r[6]))
    return out


def test_bvh_layout_holds_the_reference_tree(scenes):
    from chunkyclplugin_b200 import scenes as S
    p = scenes("entities")
    for bvh in (np.asarray(p.world_bvh, np.int32), np.asarray(p.actor_bvh, np.int32)):
        trigs = np.asarray(p.bvh_trigs, np.int32)
        rec, tris, root, ok = native.debug_bvh_layout(bvh, trigs)
        assert ok and rec.shape[0] > 0 and root == 0
        assert not tris[-8:].any() and tris.size % 8 == 0        # the read-ahead padding; 32-byte granularity
        want, got = _walk_reference(bvh, trigs), _walk_layout(rec, tris, root)
        assert len(want) == len(got)
        assert want == got
        n_inner = sum(1 for k in want if k[0] == "inner")
        assert rec.shape[0] == n_inner
    # a smaller mesh with other leaf sizes
    v, f = S._icosphere(1)
    tri = (v * 3.0 + 10.0)[f]
    palette = []
    packed = S.pack_triangles(tri, np.full(f.shape[0], 6, np.int64), np.zeros(f.shape[0], bool))
    for leaf_size in (1, 2, 3, 5):
        palette.clear()
        nodes, nxt = S.build_bvh(tri, packed, palette, 0, leaf_size=leaf_size)
        trigs = np.concatenate(palette).astype(np.int32)
        rec, tris, root, ok = native.debug_bvh_layout(np.asarray(nodes, np.int32), trigs)
        assert ok
        assert _walk_reference(np.asarray(nodes, np.int32), trigs) == _walk_layout(rec, tris, root)


def test_bvh_layout_empty_and_malformed():
    empty = np.array([0] + [0x7FC00000] * 6, dtype=np.int32)              # EMPTY_NODE: leaf 0 with NaN bounds (bvh.h:23-32)
    rec, tris, root, ok = native.debug_bvh_layout(empty, np.zeros(1, np.int32))
    assert ok and rec.shape[0] == 0 and tris.size == 8
    # an inner node whose second child lies outside the array
    bad = np.array([700, 0, 0, 0, 0, 0, 0] + [0] * 7, dtype=np.int32)
    rec, tris, root, ok = native.debug_bvh_layout(bad, np.zeros(1, np.int32))
    assert not ok and rec.shape[0] == 0
    # a leaf that points outside the triangle palette
    bad_leaf = np.array([14, 0, 1, 0, 1, 0, 1, -5, 0, 1, 0, 1, 0, 1, -999999, 0, 1, 0, 1, 0, 1], dtype=np.int32)
    rec, tris, root, ok = native.debug_bvh_layout(bad_leaf, np.zeros(8, np.int32))
    assert not ok
