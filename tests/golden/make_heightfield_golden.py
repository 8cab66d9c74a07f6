"""BASELINE configs 4 / 5 in miniature: the tessellated heightfield (one TriangleMesh, here 40x40 vertices = 3042 triangles instead
of 10 M), two point lights and a distant light, Path maxDepth 10 — at 20x12 pixels through the plain-Python restatement, in STRICT
mode (tile 4) and in FAST mode (tile 1, the mode configs 4 and 5 are measured in).  As in make_config2_golden.py the triangle is the
library's own definition and the closest hit follows its order-independent rule (adjacent triangles share edges and vertices: rays
through them meet equal distances); the renderer around it is the independent reading of the Go source.  A scene of more than 64
primitives: the CUDA path answers it with its BVH kernels (triangle-only variants), built on the device.  The generator needs a
minute or two (brute force over 3042 triangles per ray in Python), so the tests do not re-run it.

    python tests/golden/make_heightfield_golden.py        # rewrites tests/golden/heightfield_golden.json
"""
import importlib
import importlib.util
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
_spec = importlib.util.spec_from_file_location("make_config1_golden", os.path.join(HERE, "make_config1_golden.py"))
C = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(C)
W, H, SPP, GRID = 24, 14, (3, 3), 40
CASES = {"strict": ("stratified", 4), "fast": ("fast", 1)}   # sampler, tile size


def scene_and_integrator(gp):
    return gp.scenes.config4(W=W, H=H, spp=SPP, grid=GRID)


def render(gp, case):
    sampler, tile = CASES[case]
    sc = C.plain_scene(*scene_and_integrator(gp))
    sc["sampler"] = sampler
    C.Scene.ORDER_INDEPENDENT = True
    try:
        return C.render(sc, tile)
    finally:
        C.Scene.ORDER_INDEPENDENT = False


def main():
    gp = importlib.import_module("go-pbrt_b200")
    out = dict(note="made by tests/golden/make_heightfield_golden.py (plain-Python restatement of the hot path on a 3042-triangle heightfield); "
                    "film = [y][x][X, Y, Z, filterWeightSum] as float.hex()", width=W, height=H, spp=list(SPP), grid=GRID, cases={})
    for case, (sampler, tile) in CASES.items():
        t0 = time.time()
        film, st = render(gp, case)
        lit = sum(1 for row in film for p in row if p[1] > 0)
        print(f"heightfield {case} at {W}x{H}, tile {tile}: camera {st['camera']}, closest {st['closest']}, shadow {st['shadow']}, lit pixels "
              f"{lit}/{W * H}, max direct {st['max_direct']:.3f}, > 10 events {st['gt10']}, bounces {st['bounce_kinds']}, roulette tests {st['rr_tests']}, "
              f"{time.time() - t0:.0f} s", flush=True)
        out["cases"][case] = dict(tile=tile, rays=[st["camera"], st["closest"], st["shadow"]], radiance_gt10=st["gt10"],
                                  film=[[[v.hex() for v in p] for p in row] for row in film])
    with open(os.path.join(HERE, "heightfield_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote heightfield_golden.json")


if __name__ == "__main__":
    main()
