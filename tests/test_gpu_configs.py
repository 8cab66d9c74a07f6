"""Driver-run parity on the BASELINE configurations the reference itself cannot run: config 3 (100 000
spheres), config 4 (the full 2237 x 2237 heightfield, 9 999 392 triangles: BVH depth ~24, millions of nodes) and
config 5 (config-4 mesh, samples split over 8 ranks and summed).  CUDA path through the C ABI vs the oracle: >= 20 000
primary rays, >= 20 000 secondary rays and shadow segments, and a small film each — all bit for bit."""
import numpy as np
import pytest

from oracle_lib import OracleScene, camera_rays
from test_gpu_parity import _assert_film_equal, _cmp_closest, _secondary_rays

pytestmark = pytest.mark.gpu


def _ray_parity(gp, g, o, integ, what, seed):
    rng = np.random.default_rng(seed)
    xs, ys = rng.integers(0, 1920, 24000), rng.integers(0, 1080, 24000)
    ro, rd = camera_rays(integ, xs, ys)
    G = g.Intersect(ro, rd)
    _cmp_closest(G, o.intersect(ro, rd, threads=16), f"{what} primary")
    assert np.count_nonzero(G[0] >= 0) >= 5000, "the camera should see the scene"
    p, dirs = _secondary_rays(o, ro, rd, seed + 1)
    k = 24000 // max(1, len(p)) + 1  # several directions per hit point when few primary rays hit
    p, dirs = np.repeat(p, k, axis=0), np.random.default_rng(seed + 2).normal(size=(len(p) * k, 3))
    assert len(p) >= 20000
    _cmp_closest(g.Intersect(p, dirs), o.intersect(p, dirs, threads=16), f"{what} secondary")
    tm = np.where(np.arange(len(p)) % 3 == 0, 0.9999, np.inf)  # shadow segments (tMax = 1 - ShadowEpsilon) and open rays
    seg = dirs * np.where(np.arange(len(p)) % 3 == 0, 40.0, 1.0)[:, None]
    assert np.array_equal(g.IntersectP(p, seg, tm), o.intersect_p(p, seg, tm, threads=16)), f"{what}: any-hit differs"


@pytest.fixture(scope="module")
def cfg4(gp, dev):
    scene, integ = gp.scenes.config4()
    g = gp.pbrt.GpuScene(dev, scene)
    o = OracleScene(scene, 1)
    yield scene, integ, g, o
    g.close(); o.close()


def test_config4_full_mesh_rays_bit_exact(gp, cfg4):
    scene, integ, g, o = cfg4
    st = gp.pbrt.Render(g, gp.scenes.config4_integrator(48, 27), 1)  # touches the render path once: BVH statistics
    assert st["bvh_depth"] >= 20 and st["bvh_nodes"] > 4_000_000 and st["stack_overflows"] == 0
    _ray_parity(gp, g, o, integ, "config4", 41)


def test_config4_small_film_bit_exact(gp, cfg4):
    scene, _, g, o = cfg4
    integ = gp.scenes.config4_integrator(96, 54)
    st = gp.pbrt.Render(g, integ, 1)
    film = integ.GetCamera().GetFilm().pixels.copy()
    ofilm, ost = o.render(integ, 1, threads=16)
    _assert_film_equal(film, ofilm, st, ost, "config4 96x54")
    assert st["closest_rays"] > 96 * 54 * 15 * 0.5


def test_config5_rank_shares_sum_to_the_frame(gp, cfg4):
    # config 5 = the config-4 mesh with the samples split over the GPUs (FAST mode: rank r renders samples s % world == r)
    # and one film reduce.  Here at 160x90 and 8x8 samples: the eight ranks' shares, rendered one after the other on this
    # GPU, add up to the single-GPU film (summation order differs: 1e-12 relative; filterWeightSum exact), and the
    # single-GPU FAST film is the oracle's FAST film bit for bit.
    scene, _, g, o = cfg4
    integ = gp.scenes.config4_integrator(160, 90, spp=(8, 8))
    F = gp.abi.MODE_FAST
    st = gp.pbrt.Render(g, integ, 1, mode=F, groups=1)
    single = integ.GetCamera().GetFilm().pixels.copy()
    ofilm, ost = o.render(integ, 1, mode=F, threads=16)
    _assert_film_equal(single, ofilm, st, ost, "config5 single rank, FAST")
    acc, rays = np.zeros_like(single), 0
    for r in range(8):
        sr = gp.pbrt.Render(g, integ, 1, mode=F, rank=r, world=8)
        acc += integ.GetCamera().GetFilm().pixels
        rays += sr["closest_rays"] + sr["shadow_rays"]
    assert rays == st["closest_rays"] + st["shadow_rays"]
    assert np.array_equal(acc[..., 3], single[..., 3])
    assert np.allclose(acc[..., :3], single[..., :3], rtol=1e-12, atol=0)


def test_config3_sphere_field_rays_and_film_bit_exact(gp, dev):
    scene, integ = gp.scenes.config3()
    g = gp.pbrt.GpuScene(dev, scene)
    o = OracleScene(scene, 1)
    _ray_parity(gp, g, o, integ, "config3", 43)
    scene2, integ2 = gp.scenes.config3(W=96, H=54, spp=(3, 3))  # same 100 000 spheres (deterministic), small film
    st = gp.pbrt.Render(g, integ2, 1)
    film = integ2.GetCamera().GetFilm().pixels.copy()
    ofilm, ost = o.render(integ2, 1, threads=16)
    _assert_film_equal(film, ofilm, st, ost, "config3 96x54")
    assert st["efloat_panics"] == 0 and st["bvh_nodes"] > 100_000
    g.close(); o.close()


# ---- on-device BVH build (csrc/gp_build.cuh): every scene of more than 64 primitives is built on the GPU by default
def _render_small(gp, dev, scene, integ, tile=8):
    g = gp.pbrt.GpuScene(dev, scene)
    st = gp.pbrt.Render(g, integ, tile)
    film = integ.GetCamera().GetFilm().pixels.copy()
    g.close()
    return film, st


@pytest.mark.parametrize("what", ["mesh", "spheres", "mixed"])
def test_device_built_tree_is_valid_and_equals_the_host_built_one(gp, dev, monkeypatch, what):
    """GOPBRT_CHECK_BVH walks the device-built tree on the host (every primitive in exactly one leaf, leaf sizes, every float32 box
    around the float64 bounds below it).  Then the same frame through the HOST builder (GOPBRT_HOST_BVH): hits do not depend on
    the tree, so the two films and ray counts must be identical bit for bit, and so must the world bound."""
    if what == "mesh":
        scene, _ = gp.scenes.config4(W=96, H=54, grid=320)      # 203 522 triangles
        integ = gp.scenes.config4_integrator(96, 54, (2, 2))
    elif what == "spheres":
        scene, integ = gp.scenes.config3(W=96, H=54, spp=(2, 2), n_spheres=30000)
    else:
        scene, integ = gp.scenes.mixed_test_scene(200, seed=11), gp.scenes.test_integrator(96, 54, spp=(2, 2), maxDepth=5)
    monkeypatch.setenv("GOPBRT_CHECK_BVH", "1")
    g = gp.pbrt.GpuScene(dev, scene)   # raises if the check fails
    wb_dev = g.WorldBound()
    g.close()
    monkeypatch.delenv("GOPBRT_CHECK_BVH")
    film_d, st_d = _render_small(gp, dev, scene, integ)
    monkeypatch.setenv("GOPBRT_HOST_BVH", "1")
    g = gp.pbrt.GpuScene(dev, scene)
    wb_host = g.WorldBound()
    g.close()
    film_h, st_h = _render_small(gp, dev, scene, integ)
    assert wb_dev == wb_host
    assert np.array_equal(film_d, film_h)
    for k in ("camera_rays", "closest_rays", "shadow_rays"):
        assert st_d[k] == st_h[k], k
    assert st_d["stack_overflows"] == 0 and st_h["stack_overflows"] == 0


def test_device_build_rejects_bad_indices(gp, dev):
    import ctypes as C
    scene, _ = gp.scenes.config4(W=32, H=18, grid=40)
    d = scene.desc()
    tri = d.triangles[5]
    keep = tri.v[1]
    tri.v[1] = d.n_vertices + 3
    h = C.c_void_p()
    assert dev.lib.gopbrt_scene_create(dev.h, C.byref(d), C.byref(h)) == gp.abi.ERR_INVALID and "vertex index" in dev.error()
    tri.v[1] = keep
    assert dev.lib.gopbrt_scene_create(dev.h, C.byref(d), C.byref(h)) == gp.abi.OK
    dev.lib.gopbrt_scene_destroy(h)
