"""Committed golden vectors (tests/golden/hotpath_golden.npz, made by tests/golden/make_golden.py from the oracle):
 - CPU: the oracle must still reproduce them bit for bit (pins the checker itself against silent drift);
 - GPU (-m gpu): the CUDA path, through the C ABI, must reproduce them bit for bit — films in STRICT and FAST mode,
   primary-ray primitive ids and hit distances.  Nothing here reads /root/reference."""
import importlib.util
import os

import numpy as np
import pytest

from oracle_lib import OracleScene

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "hotpath_golden.npz"))


CASES = ["config1_tile16", "config1_tile1", "config2_tile1", "mixed_tile8", "mixed_fast", "config2_direct_one"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden_films(gp, golden, name):
    scene, integ, tile, mode = make_golden.cases(gp)[name]
    o = OracleScene(scene, 1)
    film, st = o.render(integ, tile, mode=mode, threads=4)
    o.close()
    assert np.array_equal(film, golden[name + "_film"])
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"]] == list(golden[name + "_rays"])


def test_oracle_reproduces_golden_primary_rays(gp, golden):
    scene, ro, rd = make_golden.primary_ray_case(gp)
    o = OracleScene(scene, 1)
    prim, t, p, n = o.intersect(ro, rd)
    o.close()
    assert np.array_equal(prim, golden["config1_primary_prim"]) and np.array_equal(t, golden["config1_primary_t"])
    # the reference-faithful tree (RecursiveBuild SplitSAH + [64]-stack traversal) gives the same answers on config 1
    o = OracleScene(scene, 0)
    prim0, t0, _, _ = o.intersect(ro, rd)
    o.close()
    assert np.array_equal(prim0, golden["config1_primary_prim"]) and np.array_equal(t0, golden["config1_primary_t"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_reproduces_golden_films(gp, dev, golden, name):
    scene, integ, tile, mode = make_golden.cases(gp)[name]
    g = gp.pbrt.GpuScene(dev, scene)
    st = gp.pbrt.Render(g, integ, tile, mode=mode, groups=1)
    film = integ.GetCamera().GetFilm().pixels
    g.close()
    assert np.array_equal(film, golden[name + "_film"]), f"{name}: {np.count_nonzero(np.any(film != golden[name + '_film'], axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"]] == list(golden[name + "_rays"])
    assert st["efloat_panics"] == 0 and st["stack_overflows"] == 0


@pytest.mark.gpu
def test_gpu_reproduces_golden_primary_rays(gp, dev, golden):
    scene, ro, rd = make_golden.primary_ray_case(gp)
    g = gp.pbrt.GpuScene(dev, scene)
    prim, t, p, n = g.Intersect(ro, rd)
    g.close()
    assert np.array_equal(prim, golden["config1_primary_prim"]) and np.array_equal(t, golden["config1_primary_t"])


@pytest.mark.parametrize("mode", [0, 1])
def test_oracle_obeys_the_composition_invariants(gp, mode):
    # The exact properties of the composed Path.Li loop that tests/test_gpu_parity.py checks on the CUDA path at BASELINE size
    # (linear in the emitted radiance; maxDepth 1 and zero albedo give a black film without a shadow ray) hold for the checker too.
    P = gp.pbrt

    def render(scene, integ):
        o = OracleScene(scene, 1)
        film, st = o.render(integ, 1, mode=mode)
        o.close()
        return film, st

    dims = dict(W=64, H=36, spp=(3, 3))
    scene, integ = gp.scenes.config2(**dims)
    f1, s1 = render(scene, integ)
    scene2, integ2 = gp.scenes.config2(**dims)
    for l in scene2.lights:
        l.LEmit = [2.0 * v for v in l.LEmit]
    f2, s2 = render(scene2, integ2)
    assert np.array_equal(f2[..., 3], f1[..., 3]) and np.array_equal(f2[..., :3], 2.0 * f1[..., :3]) and f1[..., :3].max() > 0
    for k in ("camera_rays", "closest_rays", "shadow_rays"):
        assert s1[k] == s2[k], k
    scene3, integ3 = gp.scenes.config2(**dims)
    f3, s3 = render(scene3, P.NewPath(1, integ3.GetCamera(), integ3.GetSampler(), None, 1, P.Uniform))
    assert not f3[..., :3].any() and s3["closest_rays"] == s3["camera_rays"] == s1["camera_rays"] and s3["shadow_rays"] == 0
    scene4, integ4 = gp.scenes.config2(**dims)
    for prim in scene4.aggregate.primitives:
        if isinstance(prim.material, P.MatteMaterial):
            prim.material.Kd.value = [0.0, 0.0, 0.0]
    f4, s4 = render(scene4, integ4)
    assert not f4[..., :3].any() and s4["shadow_rays"] == 0 and s4["closest_rays"] > s4["camera_rays"]
