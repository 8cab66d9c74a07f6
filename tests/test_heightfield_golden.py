"""BASELINE configs 4 / 5 in miniature (a 3042-triangle heightfield mesh, two point lights, a distant light, Path maxDepth 10) against the
plain-Python restatement, in STRICT and in FAST mode: tests/golden/heightfield_golden.json holds the two 24x14 films of
tests/golden/make_heightfield_golden.py (as for config 2, the triangle is the library's own definition; the renderer around it is the
independent reading of the Go source).  The generator needs half a minute, so it is not re-run here.
 - CPU: the oracle must reproduce the films bit for bit, ray counts included (own tree and brute force).
 - GPU (-m gpu): the CUDA path — device-built BVH, triangle-only traversal kernels — must reproduce them bit for bit.
Nothing here reads /root/reference."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle_lib import OracleScene

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_heightfield_golden", os.path.join(HERE, "golden", "make_heightfield_golden.py"))
X = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(X)

with open(os.path.join(HERE, "golden", "heightfield_golden.json")) as _f:
    RAW = json.load(_f)
GOLDEN = {k: (c["tile"], np.array([[[float.fromhex(v) for v in p] for p in row] for row in c["film"]]), c["rays"])
          for k, c in RAW["cases"].items()}
CASES = sorted(GOLDEN)


def _mode(gp, name):
    return gp.abi.MODE_FAST if name == "fast" else gp.abi.MODE_STRICT


def test_golden_file_covers_what_it_claims(gp):
    assert CASES == sorted(X.CASES) and (RAW["width"], RAW["height"], RAW["spp"], RAW["grid"]) == (X.W, X.H, list(X.SPP), X.GRID)
    scene, _ = X.scene_and_integrator(gp)
    assert sum(len(m.indices) for m in scene.aggregate.primitives) == 2 * (X.GRID - 1) ** 2 == 3042
    for name, (tile, film, rays) in GOLDEN.items():
        assert tile == X.CASES[name][1] and film.shape == (X.H, X.W, 4) and np.isfinite(film).all()
        assert np.count_nonzero(film[..., 1] > 0) > X.W * X.H // 3 and rays[0] == X.W * X.H * 8 and rays[1] > rays[0] and rays[2] > 500
        assert RAW["cases"][name]["radiance_gt10"] == 0
    assert np.array_equal(GOLDEN["fast"][1][..., 3], GOLDEN["strict"][1][..., 3])


@pytest.mark.parametrize("accel", [1, 2])
@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_heightfield_films(gp, name, accel):
    tile, gf, rays = GOLDEN[name]
    scene, integ = X.scene_and_integrator(gp)
    o = OracleScene(scene, accel)
    film, st = o.render(integ, tile, mode=_mode(gp, name), threads=4)
    o.close()
    assert np.array_equal(film, gf), f"{np.count_nonzero(np.any(film != gf, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"]] == rays
    assert st["radiance_gt10"] == 0 and st["nan_samples"] == 0 and st["unsupported_material"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_reproduces_the_heightfield_films(gp, dev, name):
    tile, gf, rays = GOLDEN[name]
    scene, integ = X.scene_and_integrator(gp)
    g = gp.pbrt.GpuScene(dev, scene)
    st = gp.pbrt.Render(g, integ, tile, mode=_mode(gp, name), groups=1)
    film = integ.GetCamera().GetFilm().pixels
    g.close()
    assert np.array_equal(film, gf), f"{np.count_nonzero(np.any(film != gf, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"]] == rays
    assert st["efloat_panics"] == 0 and st["stack_overflows"] == 0 and st["radiance_gt10"] == 0
