"""BASELINE config 1 — the reference's own scene (internal/render/server.go:29-164) — against an independent restatement:
tests/golden/config1_golden.json is the 24x14 film rendered by tests/golden/make_config1_golden.py, a plain-Python reading of the
hot path (EFloat sphere intersection, TransformedPrimitive, checkerboard, distant / point / sphere-area lights, Path.Li, the
stratified sampler, the film) written from the Go source and from nothing under oracle/ or csrc/.
 - CPU: the oracle must reproduce it bit for bit, ray counts included, with the reference-faithful BVH (accel 0) and its own tree (1);
   the generator must still produce the committed file.
 - GPU (-m gpu): the CUDA path, through the C ABI, must reproduce it bit for bit — flat table and BVH kernels both.
Nothing here reads /root/reference."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle_lib import OracleScene

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_config1_golden", os.path.join(HERE, "golden", "make_config1_golden.py"))
C = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(C)

with open(os.path.join(HERE, "golden", "config1_golden.json")) as _f:
    RAW = json.load(_f)
FILM = np.array([[[float.fromhex(v) for v in p] for p in row] for row in RAW["film"]])
RAYS = RAW["rays"] + [RAW["nondelta_estimates"]]


def test_golden_file_covers_what_it_claims():
    assert (RAW["width"], RAW["height"], RAW["spp"], RAW["tile"]) == (C.W, C.H, list(C.SPP), C.TILE)
    assert FILM.shape == (C.H, C.W, 4) and np.isfinite(FILM).all()
    assert np.count_nonzero(FILM[..., 1] > 0) == C.W * C.H     # the distant light reaches every pixel of the ground disks
    assert RAYS[0] == C.W * C.H * 8 and RAYS[1] > 3 * RAYS[0] and RAYS[2] > RAYS[0] and RAYS[3] > 1000
    assert float.fromhex(RAW["world_radius"]) > 17000         # two 10 km disks: Distant.Preprocess's bounding sphere


def test_generator_is_deterministic_and_matches_the_committed_file(gp):
    film, st = C.render(C.plain_scene(*C.scene_and_integrator(gp)), C.TILE)
    assert np.array_equal(np.array(film), FILM) and [st["camera"], st["closest"], st["shadow"], st["nondelta"]] == RAYS


@pytest.mark.parametrize("accel", [0, 1])
def test_oracle_reproduces_the_independent_config1_film(gp, accel):
    scene, integ = C.scene_and_integrator(gp)
    o = OracleScene(scene, accel)
    film, st = o.render(integ, C.TILE, mode=gp.abi.MODE_STRICT, threads=2)
    o.close()
    assert np.array_equal(film, FILM), f"{np.count_nonzero(np.any(film != FILM, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["dead_mis_rays"]] == RAYS
    assert st["radiance_gt10"] == 0 and st["nan_samples"] == 0 and st["unsupported_material"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("no_flat", [False, True])
def test_gpu_reproduces_the_independent_config1_film(gp, dev, monkeypatch, no_flat):
    if no_flat:
        monkeypatch.setenv("GOPBRT_NO_FLAT", "1")   # the BVH kernels instead of the flat table
    scene, integ = C.scene_and_integrator(gp)
    g = gp.pbrt.GpuScene(dev, scene)
    st = gp.pbrt.Render(g, integ, C.TILE, mode=gp.abi.MODE_STRICT)
    film = integ.GetCamera().GetFilm().pixels
    g.close()
    assert np.array_equal(film, FILM), f"{np.count_nonzero(np.any(film != FILM, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["dead_mis_rays"]] == RAYS
    assert st["efloat_panics"] == 0 and st["stack_overflows"] == 0 and st["radiance_gt10"] == 0
