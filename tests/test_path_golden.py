"""The COMPOSED hot path against an independent restatement: tests/golden/path_golden.json holds two small films rendered
by tests/golden/make_path_golden.py — a plain-Python reading of integrator.go:228-350 (Render / renderWorker), path.go:32-157
(Path.Li), integrator.go:46-195 (UniformSampleOneLight / EstimateDirect), disk.go:64-158, transform.go:302-333, bounds.go:149-192,
reflection.go:128-298, pixel.go / sampler.go / stratified.go and film.go:106-140,211-248, written from the Go source and from
nothing under oracle/ or csrc/.
 - CPU: the oracle must reproduce them bit for bit, ray counts included, with the reference-faithful BVH (accel 0), its own
   tree (1) and brute force (2);  the generator must still produce the committed file (it is deterministic).
 - GPU (-m gpu): the CUDA path, through the C ABI, must reproduce them bit for bit — flat table and BVH kernels both.
Nothing here reads /root/reference."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle_lib import OracleScene

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_path_golden", os.path.join(HERE, "golden", "make_path_golden.py"))
M = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(M)


def _load():
    with open(os.path.join(HERE, "golden", "path_golden.json")) as f:
        g = json.load(f)
    return {k: (c["tile"], np.array([[[float.fromhex(v) for v in p] for p in row] for row in c["film"]]), c["rays"] + [c["nondelta_estimates"], c["dead_mis_rays"]])
            for k, c in g["cases"].items()}


GOLDEN = _load()
CASES = sorted(GOLDEN)


def test_golden_file_covers_what_it_claims():
    assert CASES == sorted(f"tile{t}" for t in M.TILE_SIZES)
    with open(os.path.join(HERE, "golden", "path_golden.json")) as f:
        for c in json.load(f)["cases"].values():
            cov = c["coverage"]
            assert cov["russian_roulette_tests"] > 50
            b = cov["bounces"]   # lobe : sampled type (16 = Specular, 1 = Reflection, 2 = Transmission; the mirror answers 0)
            assert b["lambert:0"] > 1000 and b["specrefl:0"] > 30 and b["fresnel:17"] > 20 and b["fresnel:18"] > 50
    for tile, film, rays in GOLDEN.values():
        assert film.shape == (12, 16, 4) and np.isfinite(film).all()
        assert np.count_nonzero(film[..., 1] > 0) > 100        # most of the frame is lit
        assert rays[0] == 16 * 12 * 8                          # 3x3 strata, the first one is skipped (sampler.go:29-35)
        assert rays[1] > 2 * rays[0] and rays[2] > rays[0] // 2   # paths do bounce and do test visibility
        # EstimateDirect's BSDF-sampling leg (integrator.go:139-193) adds nothing in the reference (the hit primitive's area light is
        # always nil).  The library's `dead_mis_rays` counts one per area-light estimate (rays[3]); the reference would trace that ray
        # only when SampleF succeeds and PdfLi != 0 (rays[4]) — the counter is an upper bound, and no film value depends on it
        assert 0 < rays[4] < rays[3]
        # box filter of radius 1, samples on the pixels' upper-left corners (the 2-D tables are zeros): a pixel collects its own
        # samples and those of its right / lower neighbours — the last pixel only its own
        assert film[5, 5, 3] == 4 * 8 and film[11, 15, 3] == 8 and film[11, 3, 3] == 2 * 8


@pytest.mark.parametrize("tile", M.TILE_SIZES)
def test_generator_is_deterministic_and_matches_the_committed_file(gp, tile):
    scene, integ = M.scene_and_integrator(gp)
    film, st = M.render(M.plain_scene(scene, integ), tile)
    t, gf, rays = GOLDEN[f"tile{tile}"]
    assert t == tile and np.array_equal(np.array(film), gf) and [st["camera"], st["closest"], st["shadow"], st["nondelta"], st["dead_mis"]] == rays


@pytest.mark.parametrize("accel", [0, 1, 2])
@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_independent_path_films(gp, name, accel):
    tile, gf, rays = GOLDEN[name]
    scene, integ = M.scene_and_integrator(gp)
    o = OracleScene(scene, accel)
    film, st = o.render(integ, tile, mode=gp.abi.MODE_STRICT, threads=2)
    o.close()
    assert np.array_equal(film, gf), f"{np.count_nonzero(np.any(film != gf, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["dead_mis_rays"]] == rays[:4]
    assert st["radiance_gt10"] == 0 and st["nan_samples"] == 0 and st["unsupported_material"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("no_flat", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_gpu_reproduces_the_independent_path_films(gp, dev, monkeypatch, name, no_flat):
    tile, gf, rays = GOLDEN[name]
    if no_flat:
        monkeypatch.setenv("GOPBRT_NO_FLAT", "1")   # the BVH kernels instead of the flat table
    scene, integ = M.scene_and_integrator(gp)
    g = gp.pbrt.GpuScene(dev, scene)
    st = gp.pbrt.Render(g, integ, tile, mode=gp.abi.MODE_STRICT)
    film = integ.GetCamera().GetFilm().pixels
    g.close()
    assert np.array_equal(film, gf), f"{np.count_nonzero(np.any(film != gf, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["dead_mis_rays"]] == rays[:4]
    assert st["efloat_panics"] == 0 and st["stack_overflows"] == 0 and st["radiance_gt10"] == 0
