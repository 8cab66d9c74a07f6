"""The composed DirectLighting hot path against an independent restatement: tests/golden/direct_golden.json holds two small films
(UniformSampleAll and UniformSampleOne, maxDepth 5) rendered by tests/golden/make_direct_golden.py — a plain-Python reading of
directlighting.go:62-119, integrator.go:21-76 and :352-434, glass.go:38-73 (two lobes when allowMultipleLobes is false) and
reflection.go:128-277 with several lobes, on top of tests/golden/make_path_golden.py.
 - CPU: the oracle must reproduce them bit for bit, ray counts included; the generator must still produce the committed file.
 - GPU (-m gpu): the CUDA path, through the C ABI, must reproduce them bit for bit — flat table and BVH kernels both.
Nothing here reads /root/reference."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle_lib import OracleScene

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_direct_golden", os.path.join(HERE, "golden", "make_direct_golden.py"))
D = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(D)

with open(os.path.join(HERE, "golden", "direct_golden.json")) as _f:
    RAW = json.load(_f)["cases"]
GOLDEN = {k: (c["strategy"], c["tile"], np.array([[[float.fromhex(v) for v in p] for p in row] for row in c["film"]]),
              c["rays"] + [c["nondelta_estimates"]]) for k, c in RAW.items()}
CASES = sorted(GOLDEN)


def test_golden_file_covers_what_it_claims():
    assert CASES == sorted(D.CASES)
    for name, c in RAW.items():
        assert c["max_depth"] == D.MAX_DEPTH and c["tile"] == D.TILE
        assert c["coverage"]["transmitted_rays"] > 100 and c["coverage"]["deepest_level"] == 2   # Li at depth 0, 2 and 4
        film = GOLDEN[name][2]
        assert film.shape == (12, 16, 4) and np.isfinite(film).all() and np.count_nonzero(film[..., 1] > 0) > 100
        # one closest-hit query per camera ray and per transmitted ray (SpecularReflect never finds a lobe)
        assert c["rays"][1] == c["rays"][0] + c["coverage"]["transmitted_rays"]
    # every light at every hit against one light per hit
    assert RAW["all"]["nondelta_estimates"] > 3 * RAW["one"]["nondelta_estimates"]


@pytest.mark.parametrize("name", CASES)
def test_generator_is_deterministic_and_matches_the_committed_file(gp, name):
    strategy, tile, gf, rays = GOLDEN[name]
    sc = D.M.plain_scene(*D.M.scene_and_integrator(gp))
    film, st = D.render(sc, tile, strategy)
    assert np.array_equal(np.array(film), gf) and [st["camera"], st["closest"], st["shadow"], st["nondelta"]] == rays


@pytest.mark.parametrize("accel", [0, 1, 2])
@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_independent_direct_lighting_films(gp, name, accel):
    strategy, tile, gf, rays = GOLDEN[name]
    scene, integ = D.scene_and_integrator(gp, strategy)
    o = OracleScene(scene, accel)
    film, st = o.render(integ, tile, mode=gp.abi.MODE_STRICT, threads=2)
    o.close()
    assert np.array_equal(film, gf), f"{np.count_nonzero(np.any(film != gf, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["dead_mis_rays"]] == rays
    assert st["radiance_gt10"] == 0 and st["nan_samples"] == 0 and st["unsupported_material"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("no_flat", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_gpu_reproduces_the_independent_direct_lighting_films(gp, dev, monkeypatch, name, no_flat):
    strategy, tile, gf, rays = GOLDEN[name]
    if no_flat:
        monkeypatch.setenv("GOPBRT_NO_FLAT", "1")   # the BVH kernels instead of the flat table
    scene, integ = D.scene_and_integrator(gp, strategy)
    g = gp.pbrt.GpuScene(dev, scene)
    st = gp.pbrt.Render(g, integ, tile, mode=gp.abi.MODE_STRICT)
    film = integ.GetCamera().GetFilm().pixels
    g.close()
    assert np.array_equal(film, gf), f"{np.count_nonzero(np.any(film != gf, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["dead_mis_rays"]] == rays
    assert st["efloat_panics"] == 0 and st["stack_overflows"] == 0 and st["radiance_gt10"] == 0
