"""BASELINE configs 1 and 2 with their BASELINE samplers (Stratified 4x4 / 8x8, tileSize 16; config 2 also in FAST mode) at 96x54 /
48x27: tests/golden/baseline_spp_golden.npz holds the films of tests/golden/make_baseline_spp_golden.py — the plain-Python
restatement, about 80 000 camera rays and 180 000 - 250 000 closest-hit queries per film.  The generator takes minutes, so it is not
re-run here (tests/test_config1_golden.py / test_config2_golden.py do that at small size).
 - CPU: the oracle must reproduce the films bit for bit, ray counts included.
 - GPU (-m gpu): the CUDA path, through the C ABI, must reproduce them bit for bit.
Nothing here reads /root/reference."""
import importlib.util
import os

import numpy as np
import pytest

from oracle_lib import OracleScene

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_baseline_spp_golden", os.path.join(HERE, "golden", "make_baseline_spp_golden.py"))
B = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(B)
GOLDEN = np.load(os.path.join(HERE, "golden", "baseline_spp_golden.npz"))
CASES = sorted(B.CASES)


def _mode(gp, name):
    return gp.abi.MODE_FAST if B.CASES[name][1] == "fast" else gp.abi.MODE_STRICT


def test_golden_file_covers_what_it_claims():
    assert sorted(k[:-5] for k in GOLDEN.files if k.endswith("_film")) == CASES
    for name in CASES:
        (cfg, w, h, spp), _, _, _ = B.CASES[name]
        film, rays = GOLDEN[name + "_film"], GOLDEN[name + "_rays"]
        assert film.shape == (h, w, 4) and np.isfinite(film).all()
        assert rays[0] == w * h * (spp[0] * spp[1] - 1) > 75000 and rays[1] > 2 * rays[0] and rays[4] == 0


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_baseline_sampler_films(gp, name):
    scene, integ = B.scene_and_integrator(gp, name)
    o = OracleScene(scene, 1)
    film, st = o.render(integ, B.CASES[name][2], mode=_mode(gp, name), threads=4)
    o.close()
    gf = GOLDEN[name + "_film"]
    assert np.array_equal(film, gf), f"{np.count_nonzero(np.any(film != gf, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["dead_mis_rays"], st["radiance_gt10"]] == list(GOLDEN[name + "_rays"])


def test_reference_faithful_bvh_reproduces_config1_too(gp):
    scene, integ = B.scene_and_integrator(gp, "config1_strict")
    o = OracleScene(scene, 0)   # RecursiveBuild(SplitSAH) + the [64]-stack traversal of bvh.go:659-712
    film, st = o.render(integ, 16, mode=gp.abi.MODE_STRICT, threads=4)
    o.close()
    assert np.array_equal(film, GOLDEN["config1_strict_film"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_reproduces_the_baseline_sampler_films(gp, dev, name):
    scene, integ = B.scene_and_integrator(gp, name)
    g = gp.pbrt.GpuScene(dev, scene)
    st = gp.pbrt.Render(g, integ, B.CASES[name][2], mode=_mode(gp, name), groups=1)
    film = integ.GetCamera().GetFilm().pixels
    g.close()
    gf = GOLDEN[name + "_film"]
    assert np.array_equal(film, gf), f"{np.count_nonzero(np.any(film != gf, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["dead_mis_rays"], st["radiance_gt10"]] == list(GOLDEN[name + "_rays"])
    assert st["efloat_panics"] == 0 and st["stack_overflows"] == 0
