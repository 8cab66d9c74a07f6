"""Oren-Nayar surfaces, rotating TransformedPrimitives, reverseOrientation spheres, a thin-lens camera, a checkerboard ground and all
four light kinds — under Path with the Stratified sampler, Path with the Random sampler and DirectLighting(UniformSampleAll) —
against the independent plain-Python restatement: tests/golden/mixed_golden.json holds the three 18x12 films of
tests/golden/make_mixed_golden.py.
 - CPU: the oracle must reproduce them bit for bit, ray counts included (reference-faithful BVH, own tree, brute force);
   the generator is deterministic.
 - GPU (-m gpu): the CUDA path, through the C ABI, must reproduce them bit for bit — flat table and BVH kernels both.
Nothing here reads /root/reference."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle_lib import OracleScene

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_mixed_golden", os.path.join(HERE, "golden", "make_mixed_golden.py"))
X = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(X)

with open(os.path.join(HERE, "golden", "mixed_golden.json")) as _f:
    RAW = json.load(_f)
GOLDEN = {k: (np.array([[[float.fromhex(v) for v in p] for p in row] for row in c["film"]]), c["rays"] + [c["nondelta_estimates"]])
          for k, c in RAW["cases"].items()}
CASES = sorted(GOLDEN)


def test_golden_file_covers_what_it_claims():
    assert CASES == sorted(X.CASES) and (RAW["width"], RAW["height"], RAW["tile"]) == (X.W, X.H, X.TILE)
    for name, (film, rays) in GOLDEN.items():
        assert np.isfinite(film).all()
        if name == "path_crop":    # ceil(18 * 0.2) .. ceil(18 * 0.9) by ceil(12 * 0.1) .. ceil(12 * 0.8); 2x3 strata minus the skipped first
            assert film.shape == (10 - 2, 17 - 4, 4) and rays[0] == 8 * 13 * 5
            # a box filter of radius (1.5, 0.75) on pixel-corner samples reaches x-2 .. x+1 and y-1 .. y (film.go:218-222)
            assert film[4, 6, 3] == 4 * 2 * 5
            continue
        assert film.shape == (X.H, X.W, 4)
        if name == "path_power":   # no light is ever sampled: black, no visibility test, but the paths still bounce
            assert not film[..., :3].any() and rays[2] == 0 and rays[3] == 0 and rays[1] > rays[0]
        else:
            assert np.count_nonzero(film[..., 1] > 0) > X.W * X.H // 2
    for name in ("path_stratified", "path_random"):
        b = RAW["cases"][name]["bounces"]
        assert b["orennayar:0"] > 1000 and b["specrefl:0"] > 50 and b["fresnel:17"] > 50 and b["fresnel:18"] > 50 and b["lambert:0"] > 20
    assert GOLDEN["path_random"][1][0] == X.W * X.H * 6 and GOLDEN["path_stratified"][1][0] == X.W * X.H * 8   # sample 0 is skipped
    assert RAW["cases"]["direct_all"]["transmitted_rays"] > 50 and RAW["cases"]["direct_one"]["transmitted_rays"] > 50
    assert GOLDEN["direct_all"][1][3] > 3 * GOLDEN["direct_one"][1][3]


@pytest.mark.parametrize("name", CASES)
def test_generator_is_deterministic_and_matches_the_committed_file(gp, name):
    film, st = X.render(X.plain(gp, name), name)
    gf, rays = GOLDEN[name]
    assert np.array_equal(np.array(film), gf) and [st["camera"], st["closest"], st["shadow"], st["nondelta"]] == rays


@pytest.mark.parametrize("accel", [0, 1, 2])
@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_independent_mixed_films(gp, name, accel):
    gf, rays = GOLDEN[name]
    scene, integ = X.scene_and_integrator(gp, name)
    o = OracleScene(scene, accel)
    film, st = o.render(integ, X.TILES.get(name, X.TILE), mode=gp.abi.MODE_STRICT, threads=2)
    o.close()
    assert np.array_equal(film, gf), f"{np.count_nonzero(np.any(film != gf, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["dead_mis_rays"]] == rays
    assert st["radiance_gt10"] == 0 and st["nan_samples"] == 0 and st["unsupported_material"] == 0


def test_oracle_counts_the_reference_panic_and_carries_on(gp):
    # UniformSampleOneLight > 10: the reference panics (integrator.go:72-74); the library counts and goes on with the value
    c = RAW["bright"]
    gf = np.array([[[float.fromhex(v) for v in p] for p in row] for row in c["film"]])
    scene, integ = X.bright_scene_and_integrator(gp)
    o = OracleScene(scene, 1)
    film, st = o.render(integ, X.TILE, mode=gp.abi.MODE_STRICT, threads=2)
    o.close()
    assert np.array_equal(film, gf), f"{np.count_nonzero(np.any(film != gf, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["dead_mis_rays"]] == c["rays"] + [c["nondelta_estimates"]]
    assert st["radiance_gt10"] == c["radiance_gt10"] > 20


@pytest.mark.gpu
@pytest.mark.parametrize("no_flat", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_gpu_reproduces_the_independent_mixed_films(gp, dev, monkeypatch, name, no_flat):
    gf, rays = GOLDEN[name]
    if no_flat:
        monkeypatch.setenv("GOPBRT_NO_FLAT", "1")   # the BVH kernels instead of the flat table
    scene, integ = X.scene_and_integrator(gp, name)
    g = gp.pbrt.GpuScene(dev, scene)
    st = gp.pbrt.Render(g, integ, X.TILES.get(name, X.TILE), mode=gp.abi.MODE_STRICT)
    film = integ.GetCamera().GetFilm().pixels
    g.close()
    assert np.array_equal(film, gf), f"{np.count_nonzero(np.any(film != gf, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["dead_mis_rays"]] == rays
    assert st["efloat_panics"] == 0 and st["stack_overflows"] == 0 and st["radiance_gt10"] == 0
