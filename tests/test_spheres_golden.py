"""A cluster of matte / mirror / glass spheres under sphere area lights and a distant light (config 3's ingredients) against the
independent plain-Python restatement: tests/golden/spheres_golden.json is the 20x14 film of tests/golden/make_spheres_golden.py.
 - CPU: the oracle must reproduce it bit for bit, ray counts included, with all three aggregates; the generator is deterministic.
 - GPU (-m gpu): the CUDA path, through the C ABI, must reproduce it bit for bit — flat table and BVH kernels both.
Nothing here reads /root/reference."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle_lib import OracleScene

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_spheres_golden", os.path.join(HERE, "golden", "make_spheres_golden.py"))
S = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(S)

with open(os.path.join(HERE, "golden", "spheres_golden.json")) as _f:
    RAW = json.load(_f)
FILM = np.array([[[float.fromhex(v) for v in p] for p in row] for row in RAW["film"]])
RAYS = RAW["rays"] + [RAW["nondelta_estimates"]]


def test_golden_file_covers_what_it_claims():
    assert (RAW["width"], RAW["height"], RAW["spp"], RAW["tile"]) == (S.W, S.H, list(S.SPP), S.TILE)
    assert FILM.shape == (S.H, S.W, 4) and np.isfinite(FILM).all() and np.count_nonzero(FILM[..., 1] > 0) > S.W * S.H // 2
    b = RAW["coverage"]["bounces"]   # lobe : sampled type (17 = specular reflection, 18 = specular transmission; the mirror answers 0)
    assert b["lambert:0"] > 1000 and b["specrefl:0"] > 100 and b["fresnel:17"] > 100 and b["fresnel:18"] > 100
    assert RAW["coverage"]["russian_roulette_tests"] > 100 and RAYS[3] > 1000


def test_generator_is_deterministic_and_matches_the_committed_file(gp):
    film, st = S.C.render(S.C.plain_scene(*S.scene_and_integrator(gp)), S.TILE)
    assert np.array_equal(np.array(film), FILM) and [st["camera"], st["closest"], st["shadow"], st["nondelta"]] == RAYS


@pytest.mark.parametrize("accel", [0, 1, 2])
def test_oracle_reproduces_the_independent_sphere_cluster_film(gp, accel):
    scene, integ = S.scene_and_integrator(gp)
    o = OracleScene(scene, accel)
    film, st = o.render(integ, S.TILE, mode=gp.abi.MODE_STRICT, threads=2)
    o.close()
    assert np.array_equal(film, FILM), f"{np.count_nonzero(np.any(film != FILM, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["dead_mis_rays"]] == RAYS
    assert st["radiance_gt10"] == 0 and st["nan_samples"] == 0 and st["unsupported_material"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("no_flat", [False, True])
def test_gpu_reproduces_the_independent_sphere_cluster_film(gp, dev, monkeypatch, no_flat):
    if no_flat:
        monkeypatch.setenv("GOPBRT_NO_FLAT", "1")   # the BVH kernels instead of the flat table
    scene, integ = S.scene_and_integrator(gp)
    g = gp.pbrt.GpuScene(dev, scene)
    st = gp.pbrt.Render(g, integ, S.TILE, mode=gp.abi.MODE_STRICT)
    film = integ.GetCamera().GetFilm().pixels
    g.close()
    assert np.array_equal(film, FILM), f"{np.count_nonzero(np.any(film != FILM, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["dead_mis_rays"]] == RAYS
    assert st["efloat_panics"] == 0 and st["stack_overflows"] == 0 and st["radiance_gt10"] == 0
