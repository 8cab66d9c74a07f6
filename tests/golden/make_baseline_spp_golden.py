"""BASELINE configs 1 and 2 with their BASELINE samplers (Stratified 4x4 / 8x8, the reference's tileSize 16) at reduced resolution
(96x54 and 48x27), through the plain-Python restatement — about 80 000 camera rays per film instead of the ~2000 of the small
goldens, so that rarer branches (total internal reflection, grazing shadow rays, long roulette chains) are met.  Stored as float64
arrays in tests/golden/baseline_spp_golden.npz (the generator needs minutes, so the tests compare the oracle and the CUDA path with
the committed file and do not re-run it; the small goldens prove the generator deterministic).

    python tests/golden/make_baseline_spp_golden.py        # rewrites tests/golden/baseline_spp_golden.npz (~4 min)
"""
import importlib
import importlib.util
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
_spec = importlib.util.spec_from_file_location("make_config2_golden", os.path.join(HERE, "make_config2_golden.py"))
X2 = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(X2)
C = X2.C

# name -> (scene builder args, sampler, tile, order-independent closest hit)
CASES = {
    "config1_strict": (("config1", 96, 54, (4, 4)), "stratified", 16, False),
    "config2_strict": (("config2", 48, 27, (8, 8)), "stratified", 16, True),
    "config2_fast": (("config2", 48, 27, (8, 8)), "fast", 1, True),
}


def scene_and_integrator(gp, name):
    (cfg, w, h, spp), _, _, _ = CASES[name]
    return getattr(gp.scenes, cfg)(W=w, H=h, spp=spp)


def render(gp, name):
    _, sampler, tile, order_independent = CASES[name]
    sc = C.plain_scene(*scene_and_integrator(gp, name))
    sc["sampler"] = sampler
    C.Scene.ORDER_INDEPENDENT = order_independent
    try:
        return C.render(sc, tile)
    finally:
        C.Scene.ORDER_INDEPENDENT = False


def main():
    gp = importlib.import_module("go-pbrt_b200")
    data = {}
    for name in CASES:
        t0 = time.time()
        film, st = render(gp, name)
        print(f"{name}: camera {st['camera']}, closest {st['closest']}, shadow {st['shadow']}, area-light estimates {st['nondelta']}, "
              f"> 10 events {st['gt10']}, bounces {st['bounce_kinds']}, roulette tests {st['rr_tests']}, {time.time() - t0:.0f} s", flush=True)
        data[name + "_film"] = np.array(film, dtype=np.float64)
        data[name + "_rays"] = np.array([st["camera"], st["closest"], st["shadow"], st["nondelta"], st["gt10"]], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "baseline_spp_golden.npz"), **data)
    print("wrote baseline_spp_golden.npz")


if __name__ == "__main__":
    main()
