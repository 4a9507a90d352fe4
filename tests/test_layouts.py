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
