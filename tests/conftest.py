import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def scenes():
    """Small scenes shared by the parity tests (built once)."""
    from chunkyclplugin_b200 import scenes as S
    cache = {}

    def get(name):
        if name not in cache:
            if name == "terrain64":
                cache[name] = S.terrain_scene(64, 160, 90, seed=7)
            elif name == "tiny8":
                cache[name] = S.terrain_scene(8, 64, 36, seed=3)            # depth 3: top table only two cells wide
            elif name == "terrain128":
                cache[name] = S.terrain_scene(128, 160, 90, seed=5)         # odd depth (7)
            elif name == "terrain256":
                cache[name] = S.terrain_scene(256, 320, 180)
            elif name == "terrain256_nosun":
                cache[name] = S.terrain_scene(256, 160, 90, sun=False)
            elif name == "decorated":
                cache[name] = S.terrain_scene(64, 160, 90, seed=11, decorate=True)
            elif name == "mixed":
                cache[name] = S.mixed_test_scene()
            elif name == "indoor":
                cache[name] = S.indoor_scene(64, 128, 72)
            elif name == "large512":
                cache[name] = S.large_world_scene(512, 64, 512, 160, 90)     # depth 9: top table too large for shared memory
            elif name == "large2048":
                cache[name] = S.large_world_scene(2048, 32, 2048, 128, 72)   # depth 11 (BASELINE config 5 shape)
            elif name == "entities":
                cache[name] = S.entity_scene(128, 160, 90, n_world=96, n_actor=8, subdiv=1)
            else:
                raise KeyError(name)
        return cache[name]
    return get


@pytest.fixture(scope="session")
def cuda_ctx():
    from chunkyclplugin_b200 import native
    ctx = native.Context(0)
    yield ctx
    ctx.close()


def load_scene(ctx, p):
    """Upload a PackedScene through the C ABI in the order the reference's loader uses."""
    ctx.scene_begin()
    ctx.set_atlas(p.atlas)
    ctx.set_block_palette(p.block_palette)
    ctx.set_material_palette(p.mat_palette)
    ctx.set_aabb_models(p.aabb_models)
    ctx.set_quad_models(p.quad_models)
    ctx.set_triangles(p.bvh_trigs)
    ctx.set_world_bvh(p.world_bvh)
    ctx.set_actor_bvh(p.actor_bvh)
    ctx.set_sun(p.sun)
    ctx.set_sky(p.sky, p.sky_intensity)
    ctx.set_octree(p.octree, p.octree_depth)
    ctx.scene_commit()
    ctx.camera_set(p.projector_type, p.camera)
    ctx.render_begin(p.width, p.height)
