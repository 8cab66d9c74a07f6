"""BASELINE config 2 — the scene bench.py's headline is quoted on — against the plain-Python restatement, in STRICT and in FAST mode:
tests/golden/config2_golden.json holds the two 20x12 films of tests/golden/make_config2_golden.py (see its docstring for what a scene
with triangles can and cannot pin: the triangle is the library's own definition; the renderer around it is the independent reading of
the Go source).
 - CPU: the oracle must reproduce them bit for bit, ray counts included (own tree and brute force — the reference-faithful builder
   has no triangles); the generator is deterministic.
 - GPU (-m gpu): the CUDA path, through the C ABI, must reproduce them bit for bit — flat table and BVH kernels both.
Nothing here reads /root/reference."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle_lib import OracleScene

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_config2_golden", os.path.join(HERE, "golden", "make_config2_golden.py"))
X = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(X)

with open(os.path.join(HERE, "golden", "config2_golden.json")) as _f:
    RAW = json.load(_f)
GOLDEN = {k: (c["tile"], np.array([[[float.fromhex(v) for v in p] for p in row] for row in c["film"]]), c["rays"] + [c["nondelta_estimates"]])
          for k, c in RAW["cases"].items()}
CASES = sorted(GOLDEN)


def _mode(gp, name):
    return gp.abi.MODE_FAST if name == "fast" else gp.abi.MODE_STRICT


def test_golden_file_covers_what_it_claims():
    assert CASES == sorted(X.CASES) and (RAW["width"], RAW["height"], RAW["spp"]) == (X.W, X.H, list(X.SPP))
    for name, (tile, film, rays) in GOLDEN.items():
        assert tile == X.CASES[name][1] and film.shape == (X.H, X.W, 4) and np.isfinite(film).all()
        assert np.count_nonzero(film[..., 1] > 0) > X.W * X.H // 2 and rays[0] == X.W * X.H * 8 and rays[1] > 2 * rays[0]
        b = RAW["cases"][name]["bounces"]
        assert b["lambert:0"] > 2000 and b["fresnel:17"] > 100 and b["fresnel:18"] > 50
        assert RAW["cases"][name]["radiance_gt10"] == 0
        # on this film the reference's running-tMax rule and the library's order-independent closest hit agree
        assert RAW["cases"][name]["same_film_under_the_running_tmax_rule"] is True
    assert np.array_equal(GOLDEN["fast"][1][..., 3], GOLDEN["strict"][1][..., 3])   # same weights, other sample sequences


@pytest.mark.parametrize("name", CASES)
def test_generator_is_deterministic_and_matches_the_committed_file(gp, name):
    film, st = X.render(gp, name)
    tile, gf, rays = GOLDEN[name]
    assert np.array_equal(np.array(film), gf) and [st["camera"], st["closest"], st["shadow"], st["nondelta"]] == rays


@pytest.mark.parametrize("accel", [1, 2])
@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_config2_films(gp, name, accel):
    tile, gf, rays = GOLDEN[name]
    scene, integ = X.scene_and_integrator(gp)
    o = OracleScene(scene, accel)
    film, st = o.render(integ, tile, mode=_mode(gp, name), threads=2)
    o.close()
    assert np.array_equal(film, gf), f"{np.count_nonzero(np.any(film != gf, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["dead_mis_rays"]] == rays
    assert st["radiance_gt10"] == 0 and st["nan_samples"] == 0 and st["unsupported_material"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("no_flat", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_gpu_reproduces_the_config2_films(gp, dev, monkeypatch, name, no_flat):
    tile, gf, rays = GOLDEN[name]
    if no_flat:
        monkeypatch.setenv("GOPBRT_NO_FLAT", "1")   # the BVH kernels instead of the flat table
    scene, integ = X.scene_and_integrator(gp)
    g = gp.pbrt.GpuScene(dev, scene)
    st = gp.pbrt.Render(g, integ, tile, mode=_mode(gp, name), groups=1)
    film = integ.GetCamera().GetFilm().pixels
    g.close()
    assert np.array_equal(film, gf), f"{np.count_nonzero(np.any(film != gf, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["dead_mis_rays"]] == rays
    assert st["efloat_panics"] == 0 and st["stack_overflows"] == 0 and st["radiance_gt10"] == 0
