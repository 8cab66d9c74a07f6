"""Generates tests/golden/hotpath_golden.npz from the ORACLE (oracle/liboracle.so, the C++ restatement of the Go renderer):
small films and primary-ray results of the BASELINE scenes.  The reference itself cannot be run in this image (no Go
toolchain), so these vectors pin the oracle against regressions and travel to the GPU box, where the CUDA
path must reproduce them bit for bit (tests/test_golden_fixtures.py).

    python tests/golden/make_golden.py            (re-run only when the oracle changes on purpose)
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cases(gp):
    """name -> (scene, integrator, tileSize, mode); shared with the tests so both sides build identical inputs"""
    S = gp.scenes
    out = {}
    sc, ig = S.config1(W=64, H=36)
    out["config1_tile16"] = (sc, ig, 16, gp.abi.MODE_STRICT)
    sc, ig = S.config1(W=64, H=36)
    out["config1_tile1"] = (sc, ig, 1, gp.abi.MODE_STRICT)
    sc, ig = S.config2(W=48, H=27, spp=(3, 3))
    out["config2_tile1"] = (sc, ig, 1, gp.abi.MODE_STRICT)
    sc = S.mixed_test_scene(60, seed=21)
    out["mixed_tile8"] = (sc, S.test_integrator(48, 32, spp=(3, 3), maxDepth=5), 8, gp.abi.MODE_STRICT)
    sc = S.mixed_test_scene(60, seed=21)
    out["mixed_fast"] = (sc, S.test_integrator(48, 32, spp=(3, 3), maxDepth=5), 1, gp.abi.MODE_FAST)
    sc, ig = S.config2(W=48, H=27, spp=(3, 3))
    dl = gp.pbrt.NewDirectLighting(gp.pbrt.UniformSampleOne, 5, ig.GetCamera(), ig.GetSampler(), None)
    out["config2_direct_one"] = (sc, dl, 1, gp.abi.MODE_STRICT)
    return out


def primary_ray_case(gp):
    from oracle_lib import camera_rays
    scene, integ = gp.scenes.config1()
    ys, xs = np.meshgrid(np.arange(0, 1080, 30), np.arange(0, 1920, 30), indexing="ij")
    o, d = camera_rays(integ, xs.ravel(), ys.ravel())
    return scene, o, d


def main():
    gp = importlib.import_module("go-pbrt_b200")
    from oracle_lib import OracleScene
    data = {}
    for name, (scene, integ, tile, mode) in cases(gp).items():
        o = OracleScene(scene, 1)
        film, st = o.render(integ, tile, mode=mode)
        o.close()
        data[name + "_film"] = film
        data[name + "_rays"] = np.array([st["camera_rays"], st["closest_rays"], st["shadow_rays"]], dtype=np.int64)
    scene, ro, rd = primary_ray_case(gp)
    o = OracleScene(scene, 1)
    prim, t, p, n = o.intersect(ro, rd)
    o.close()
    data["config1_primary_prim"] = prim.astype(np.int32)
    data["config1_primary_t"] = t
    np.savez_compressed(os.path.join(HERE, "hotpath_golden.npz"), **data)
    print({k: v.shape for k, v in data.items()})


if __name__ == "__main__":
    main()
