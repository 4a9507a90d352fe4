"""Known-answer tests that pin the oracle and the host logic to the reference source (CPU only)."""
import numpy as np

import oracle
from chunkyclplugin_b200 import scenes as S
from chunkyclplugin_b200.javarandom import JavaRandom, pass_seeds


def test_java_random_seed_sequence():
    # new Random(0).nextInt() x8 - the per-pass seeds of OpenClPathTracingRenderer.java:95,106-107
    assert pass_seeds(8) == [-1155484576, -723955400, 1033096058, -1690734402, -1557280266, 1327362106, -1930858313, 502539523]
    r = JavaRandom(0)
    assert [r.next_int() for _ in range(3)] == pass_seeds(3)
    assert pass_seeds(2, skip=2) == pass_seeds(4)[2:]


def test_pcg_hash_chain_from_zero():
    # randomness.h:6-11 iterated from state 0; floats = (state >> 8) / 2^24 (randomness.h:15-17)
    st, fl = oracle.rng_chain(0, 5)
    assert [int(x) for x in st] == [0x07BB2FE2, 0x270D659D, 0x5B322158, 0x9D86E4F0, 0xB31EDFB3]
    assert np.allclose(fl[:3], [0.030199945, 0.152548134, 0.356233656], rtol=0, atol=1e-9)
    assert np.array_equal(fl, (st >> 8).astype(np.float32) / np.float32(1 << 24))


def test_pass0_pixel_jitter_vectors():
    # rayTracer.cl:55-57,68-69 for pass 0: gid 0 and gid 1
    seed = pass_seeds(1)[0]
    st, fl = oracle.rng_chain((seed + 0) & 0xFFFFFFFF, 3)
    assert int(st[0]) == 0xDE2F18E4 and int(st[1]) == 0x679A44ED and int(st[2]) == 0x90E290D8
    assert abs(fl[1] - 0.40469766) < 1e-8 and abs(fl[2] - 0.56595707) < 1e-8
    st, fl = oracle.rng_chain((seed + 1) & 0xFFFFFFFF, 3)
    assert int(st[0]) == 0x55A5ACEA
    assert abs(fl[1] - 0.36708313) < 1e-8 and abs(fl[2] - 0.79857838) < 1e-8


def _ulp(a, b):
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


def test_detmath_close_to_libm():
    """The deterministic sin/cos/atan2/asin/acos stay within a few ulp / 3e-7 abs of libm on the ranges the path uses."""
    x = np.linspace(-7.0, 7.0, 400001).astype(np.float32)          # angles are 2*pi*u, sun angles, 0.03
    for fn in ("sin", "cos"):
        a, b = oracle.math_fn(fn, x), oracle.math_fn(fn, x, libm=True)
        assert np.abs(a - b).max() <= 1.2e-7
    u = np.linspace(-1.0, 1.0, 400001).astype(np.float32)
    assert _ulp(oracle.math_fn("asin", u), oracle.math_fn("asin", u, libm=True)).max() <= 3
    assert np.abs(oracle.math_fn("acos", u) - oracle.math_fn("acos", u, libm=True)).max() <= 2.4e-7
    rng = np.random.default_rng(3)
    y, xx = rng.normal(size=200000).astype(np.float32), rng.normal(size=200000).astype(np.float32)
    assert np.abs(oracle.math_fn("atan2", y, xx) - oracle.math_fn("atan2", y, xx, libm=True)).max() <= 4.8e-7
    # out-of-domain arguments give NaN (un-normalised primary directions reach acos/asin)
    assert np.isnan(oracle.math_fn("acos", np.array([1.5, -2.0], np.float32))).all()
    # exact special values
    assert oracle.math_fn("atan2", np.array([0.0], np.float32), np.array([0.0], np.float32))[0] == 0.0
    assert oracle.math_fn("sin", np.array([0.0], np.float32))[0] == 0.0
    assert oracle.math_fn("cos", np.array([0.0], np.float32))[0] == 1.0


def test_octree_builder_matches_dense_voxels():
    pal = S.Palettes()
    vox = np.zeros((64, 64, 64), np.int32)
    full = S.terrain_voxels(0, 0, 64, 256, 64, seed=5)
    vox[:, :64, :] = full[:, 40:104, :]
    tree, depth = S.build_octree(vox, pal.block_mapping)
    assert depth == 6 and tree[0] == 1 and (tree.size - 1) % 8 == 0
    rng = np.random.default_rng(0)
    for x, y, z in rng.integers(0, 64, size=(3000, 3)):
        v, level, node = S.octree_get(tree, depth, int(x), int(y), int(z))
        assert v == 2 * vox[x, y, z]
        assert 0 < node < tree.size and tree[node] == -v
    # tiled construction gives the same leaves
    tiled, _ = S.build_octree_tiled(lambda tx, ty, tz: vox[tx * 16:(tx + 1) * 16, ty * 16:(ty + 1) * 16, tz * 16:(tz + 1) * 16].copy(),
                                    6, 4, pal.block_mapping)
    for x, y, z in rng.integers(0, 64, size=(2000, 3)):
        assert S.octree_get(tiled, 6, int(x), int(y), int(z))[:2] == S.octree_get(tree, 6, int(x), int(y), int(z))[:2]
    # a uniform world is a single leaf word; an empty one too
    one, d = S.build_octree(np.full((8, 8, 8), S.STONE, np.int32), pal.block_mapping)
    assert one.tolist() == [-2 * S.STONE]


def test_bvh_layout_invariants():
    p = S.entity_scene(64, 32, 18, n_world=12, n_actor=3, subdiv=1)
    for bvh in (p.world_bvh, p.actor_bvh):
        assert bvh.size % 7 == 0
        n = bvh.size // 7
        seen_leaf = 0
        for i in range(n):
            head = int(bvh[7 * i])
            if head > 0:
                assert head % 7 == 0 and head // 7 > i + 1 and head // 7 < n      # second child after the first child's subtree
            else:
                cnt = int(p.bvh_trigs[-head])
                assert 1 <= cnt <= 4
                seen_leaf += cnt
            lo_hi = bvh[7 * i + 1:7 * i + 7].view(np.float32)
            assert lo_hi[0] <= lo_hi[1] and lo_hi[2] <= lo_hi[3] and lo_hi[4] <= lo_hi[5]
        assert seen_leaf > 0
    assert S.EMPTY_BVH[0] == 0 and np.isnan(S.EMPTY_BVH[1:].view(np.float32)).all()      # PackedBvhNode.java:16-18


def test_oracle_accumulation_is_running_mean():
    """rayTracer.cl:109-112: buf = (buf*spp + c)/(spp+1); two calls of n passes == one call of 2n passes."""
    p = S.terrain_scene(64, 48, 27, seed=7)
    o = oracle.Oracle(p)
    seeds = pass_seeds(4)
    a = o.render(seeds)
    b = o.render(seeds[:2])
    b = o.render(seeds[2:], start_spp=2, res=b)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # thread count does not change results
    c = oracle.Oracle(p).render(seeds, threads=1)
    assert np.array_equal(a.view(np.uint32), c.view(np.uint32))
    # sub-set rendering (gids) equals the full frame on those pixels
    g = np.array([0, 5, 47, 48 * 13 + 7, 48 * 27 - 1], np.int32)
    d = oracle.Oracle(p).render(seeds, gids=g).reshape(-1, 3)
    assert np.array_equal(d[g].view(np.uint32), a.reshape(-1, 3)[g].view(np.uint32))
    assert not d[1].any()


def test_oracle_sun_flag_changes_draw_order():
    """With the sun flag off no sun-sample draws happen (sky.h:69-71), so paths differ from the second segment on."""
    a = S.terrain_scene(64, 32, 18, seed=7, sun=True)
    b = S.terrain_scene(64, 32, 18, seed=7, sun=False)
    ia = oracle.Oracle(a).render(pass_seeds(2))
    ib = oracle.Oracle(b).render(pass_seeds(2))
    assert not np.array_equal(ia, ib)
    fa, fb = oracle.Oracle(a).first_hit(1), oracle.Oracle(b).first_hit(1)
    assert np.array_equal(fa["node"], fb["node"])          # first hit does not depend on the sun


def test_libm_mode_statistically_equal():
    """Swapping detmath for libm changes low bits only: 32-pass means agree to a small relative RMSE."""
    p = S.terrain_scene(64, 48, 27, seed=7)
    seeds = pass_seeds(32)
    a = oracle.Oracle(p).render(seeds).astype(np.float64)
    b = oracle.Oracle(p, math_mode=1).render(seeds).astype(np.float64)
    rel = np.sqrt(((a - b) ** 2).mean()) / np.sqrt((a ** 2).mean())
    assert rel < 0.02
    assert abs(a.mean() - b.mean()) / a.mean() < 1e-3


def test_emittance_byte_quotient_equals_unorm_table():
    """material.h:79 computes (float)((double)b / 255.0); the device reads the UNORM table entry (float)b / 255.0f instead.
    Both are the same float for every byte value (the double quotient never lands on a rounding tie of the float grid)."""
    b = np.arange(256)
    via_double = (b.astype(np.float64) / 255.0).astype(np.float32)
    via_float = b.astype(np.float32) / np.float32(255.0)
    assert np.array_equal(via_double.view(np.uint32), via_float.view(np.uint32))
