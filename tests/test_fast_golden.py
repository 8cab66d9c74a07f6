"""FAST sampler mode — bench.py's headline mode — against its specification: tests/golden/fast_golden.json holds the films of
tests/golden/make_fast_golden.py, i.e. of the independent plain-Python restatement of the reference renderer run with a Python
reading of go-pbrt_b200/go/gopbrt/fast_sampler.go (the sampler a Go host compiles to make the CPU renderer draw FAST mode's numbers),
on BASELINE config 1 and on the mixed scene with jittered strata.
 - CPU: the oracle's FAST mode must reproduce them bit for bit, ray counts included; the generator is deterministic.
 - GPU (-m gpu): the CUDA path's FAST mode, through the C ABI, must reproduce them bit for bit with one lane group (the additions
   of the reference's tile loop at tileSize 1), and to 1e-12 with identical weights and ray counts under its default lane grouping.
Nothing here reads /root/reference."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle_lib import OracleScene

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_fast_golden", os.path.join(HERE, "golden", "make_fast_golden.py"))
F = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(F)

with open(os.path.join(HERE, "golden", "fast_golden.json")) as _f:
    RAW = json.load(_f)
GOLDEN = {k: (np.array([[[float.fromhex(v) for v in p] for p in row] for row in c["film"]]), c["rays"] + [c["nondelta_estimates"]])
          for k, c in RAW["cases"].items()}
CASES = sorted(GOLDEN)


def test_golden_file_covers_what_it_claims(gp):
    assert CASES == sorted(F.CASES) and RAW["tile"] == 1
    for name, (film, rays) in GOLDEN.items():
        assert np.isfinite(film).all() and np.count_nonzero(film[..., 1] > 0) > film.shape[0] * film.shape[1] // 2
        assert rays[1] > rays[0] and rays[2] > rays[0] // 2
    # FAST is a different sample sequence, not a different estimator: same weights as the STRICT film of the same scene
    with open(os.path.join(HERE, "golden", "config1_golden.json")) as f:
        strict = np.array([[[float.fromhex(v) for v in p] for p in row] for row in json.load(f)["film"]])
    fast = GOLDEN["config1"][0]
    assert np.array_equal(fast[..., 3], strict[..., 3]) and not np.array_equal(fast[..., :3], strict[..., :3])
    # (medians: the reference's estimator is heavy-tailed — the local wi of SURVEY §0.8 — and single samples dominate a mean)
    assert abs(np.median(fast[..., 1]) / np.median(strict[..., 1]) - 1) < 0.2


@pytest.mark.parametrize("name", CASES)
def test_generator_is_deterministic_and_matches_the_committed_file(gp, name):
    film, st = F.render(gp, name)
    gf, rays = GOLDEN[name]
    assert np.array_equal(np.array(film), gf) and [st["camera"], st["closest"], st["shadow"], st["nondelta"]] == rays


@pytest.mark.parametrize("accel", [0, 1])
@pytest.mark.parametrize("name", CASES)
def test_oracle_fast_mode_reproduces_the_specified_films(gp, name, accel):
    gf, rays = GOLDEN[name]
    scene, integ = F.scene_and_integrator(gp, name)
    o = OracleScene(scene, accel)
    film, st = o.render(integ, 1, mode=gp.abi.MODE_FAST, threads=2)
    o.close()
    assert np.array_equal(film, gf), f"{np.count_nonzero(np.any(film != gf, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["dead_mis_rays"]] == rays


@pytest.mark.parametrize("name", sorted(F.PARTITIONS))
def test_oracle_reproduces_one_rank_of_a_split_frame(gp, name):
    # the multi-GPU partition is the library's, not the reference's: samples s % world == rank in FAST mode, tiles
    # t % world == rank in STRICT mode, each rank into a full-resolution film of its own
    case, sampler, tile, rank, world = F.PARTITIONS[name]
    c = RAW["partitions"][name]
    gf = np.array([[[float.fromhex(v) for v in p] for p in row] for row in c["film"]])
    pf, pst = F.render_partition(gp, name)
    assert np.array_equal(np.array(pf), gf) and [pst["camera"], pst["closest"], pst["shadow"]] == c["rays"]
    scene, integ = F.scene_and_integrator(gp, case)
    o = OracleScene(scene, 1)
    mode = gp.abi.MODE_FAST if sampler == "fast" else gp.abi.MODE_STRICT
    film, st = o.render(integ, tile, mode=mode, rank=rank, world=world, threads=2)
    o.close()
    assert np.array_equal(film, gf), f"{np.count_nonzero(np.any(film != gf, axis=2))} pixels differ"
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"]] == c["rays"]
    assert 0 < c["rays"][0] < GOLDEN["config1"][1][0]


def test_the_ranks_shares_add_up_to_the_frame(gp):
    # FAST mode's point: a pixel's samples are independent streams, so the shares of ranks 0..world-1 are disjoint and complete — their
    # films sum to the single-rank film (to rounding: the additions associate differently), weights and ray counts exactly
    total, rays = None, [0, 0, 0]
    for rank in range(3):
        F.PARTITIONS["_tmp"] = ("config1", "fast", 1, rank, 3)
        try:
            film, st = F.render_partition(gp, "_tmp")
        finally:
            del F.PARTITIONS["_tmp"]
        film = np.array(film)
        total = film if total is None else total + film
        rays = [rays[0] + st["camera"], rays[1] + st["closest"], rays[2] + st["shadow"]]
    gf, want = GOLDEN["config1"]
    assert np.array_equal(total[..., 3], gf[..., 3]) and np.allclose(total, gf, rtol=1e-12, atol=0.0) and rays == want[:3]


@pytest.mark.gpu
@pytest.mark.parametrize("groups", [1, 0])
@pytest.mark.parametrize("name", CASES)
def test_gpu_fast_mode_reproduces_the_specified_films(gp, dev, name, groups):
    gf, rays = GOLDEN[name]
    scene, integ = F.scene_and_integrator(gp, name)
    g = gp.pbrt.GpuScene(dev, scene)
    st = gp.pbrt.Render(g, integ, 1, mode=gp.abi.MODE_FAST, groups=groups)
    film = integ.GetCamera().GetFilm().pixels
    g.close()
    if groups == 1:
        assert np.array_equal(film, gf), f"{np.count_nonzero(np.any(film != gf, axis=2))} pixels differ"
    else:   # the default lane grouping adds a pixel's samples in another order
        assert np.array_equal(film[..., 3], gf[..., 3]) and np.allclose(film, gf, rtol=1e-12, atol=0.0)
    assert [st["camera_rays"], st["closest_rays"], st["shadow_rays"], st["dead_mis_rays"]] == rays
    assert st["efloat_panics"] == 0 and st["stack_overflows"] == 0
