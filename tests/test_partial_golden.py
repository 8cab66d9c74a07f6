"""Partial spheres and disks (z / phi clipping, the second-root retry, inner radii) and UV-mapped checkerboards against the independent
plain-Python restatement: tests/golden/partial_golden.json is the 20x14 film of tests/golden/make_partial_golden.py.
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
_spec = importlib.util.spec_from_file_location("make_partial_golden", os.path.join(HERE, "golden", "make_partial_golden.py"))
X = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(X)

with open(os.path.join(HERE, "golden", "partial_golden.json")) as _f:
    RAW = json.load(_f)
FILM = np.array([[[float.fromhex(v) for v in p] for p in row] for row in RAW["film"]])
RAYS = RAW["rays"]


def test_golden_file_covers_what_it_claims():
    assert (RAW["width"], RAW["height"], RAW["spp"], RAW["tile"]) == (X.W, X.H, list(X.SPP), X.TILE)
    assert FILM.shape == (X.H, X.W, 4) and np.isfinite(FILM).all() and np.count_nonzero(FILM[..., 1] > 0) > X.W * X.H * 3 // 4
    cov = RAW["coverage"]
    assert cov["second_root_retries"]["tried"] > 200 and 100 < cov["second_root_retries"]["hit"] < cov["second_root_retries"]["tried"]
    b = cov["bounces"]
    assert b["lambert:0"] > 2000 and b["specrefl:0"] > 30 and b["fresnel:17"] > 30 and b["fresnel:18"] > 10


def test_generator_is_deterministic_and_matches_the_committed_file(gp):
    film, st = X.C.render(X.C.plain_scene(*X.scene_and_integrator(gp)), X.TILE)
    assert np.array_equal(np.array(film), FILM) and [st["camera"], st["closest"], st["shadow"]] == RAYS


@pytest.mark.parametrize("accel", [0, 1, 2])
def test_oracle_reproduces_the_independent_partial_shapes_film(gp, accel):
    scene, integ = X.scene_and_integrator(gp)
    o = OracleScene(scene, accel)
    film, st = o.render(integ, X.TILE, mode=gp.abi.MODE_STRICT, threads=2)
    o.close()
    assert np.array_equal(film, FILM), f"{np.count_nonzero(np.any(film != FILM, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"]] == RAYS
    assert st["radiance_gt10"] == 0 and st["nan_samples"] == 0 and st["unsupported_material"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("no_flat", [False, True])
def test_gpu_reproduces_the_independent_partial_shapes_film(gp, dev, monkeypatch, no_flat):
    if no_flat:
        monkeypatch.setenv("GOPBRT_NO_FLAT", "1")   # the BVH kernels instead of the flat table
    scene, integ = X.scene_and_integrator(gp)
    g = gp.pbrt.GpuScene(dev, scene)
    st = gp.pbrt.Render(g, integ, X.TILE, mode=gp.abi.MODE_STRICT)
    film = integ.GetCamera().GetFilm().pixels
    g.close()
    assert np.array_equal(film, FILM), f"{np.count_nonzero(np.any(film != FILM, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"]] == RAYS
    assert st["efloat_panics"] == 0 and st["stack_overflows"] == 0 and st["radiance_gt10"] == 0
