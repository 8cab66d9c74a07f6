"""BASELINE config 2 — the workload bench.py's headline is quoted on (a Cornell-box-style room of 36 triangles, a matte and a glass
sphere, a one-sided disk area light, Path maxDepth 10) — at 20x12 pixels through the plain-Python restatement, in STRICT mode (tile
4) and in FAST mode (the FastStratified sampler, tile 1).

What this can and cannot pin.  Triangles do not exist in the reference: their definition is the library's own (pbrt-v3's watertight
test in float64; make_config1_golden.tri_intersect restates that DEFINITION), and the closest hit is taken by the library's
order-independent rule (least t over the primitives whose own bound and shape test pass with the ray's original tMax, equal t to the
lower index — the room has coplanar faces, where the reference's running-tMax answer depends on the BVH's visit order).  Everything
else on the path — the renderer's loop, Path.Li, EstimateDirect with the area light's MIS bookkeeping, the spheres, the samplers, the
film — is the independent reading of the Go source.  So this is a golden of the composed hot path ON the headline scene, as
independent as a scene with triangles can be.

    python tests/golden/make_config2_golden.py        # rewrites tests/golden/config2_golden.json
"""
import importlib
import importlib.util
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
_spec = importlib.util.spec_from_file_location("make_config1_golden", os.path.join(HERE, "make_config1_golden.py"))
C = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(C)
M = C.M
W, H, SPP = 20, 12, (3, 3)
CASES = {"strict": ("stratified", 4), "fast": ("fast", 1)}   # sampler, tile size


def scene_and_integrator(gp):
    return gp.scenes.config2(W=W, H=H, spp=SPP)


def render(gp, case, order_independent=True):
    sampler, tile = CASES[case]
    sc = C.plain_scene(*scene_and_integrator(gp))
    sc["sampler"] = sampler
    C.Scene.ORDER_INDEPENDENT = order_independent
    try:
        return C.render(sc, tile)
    finally:
        C.Scene.ORDER_INDEPENDENT = False


def main():
    gp = importlib.import_module("go-pbrt_b200")
    out = dict(note="made by tests/golden/make_config2_golden.py (plain-Python restatement of the hot path on BASELINE config 2); "
                    "film = [y][x][X, Y, Z, filterWeightSum] as float.hex()", width=W, height=H, spp=list(SPP), cases={})
    for case, (sampler, tile) in CASES.items():
        film, st = render(gp, case)
        film_running, _ = render(gp, case, order_independent=False)
        same = film == film_running
        lit = sum(1 for row in film for p in row if p[1] > 0)
        print(f"config 2 {case} at {W}x{H}, tile {tile}: camera {st['camera']}, closest {st['closest']}, shadow {st['shadow']}, area-light estimates "
              f"{st['nondelta']}, lit pixels {lit}/{W * H}, max direct {st['max_direct']:.3f}, bounces {st['bounce_kinds']}, roulette tests {st['rr_tests']}; "
              f"running-tMax rule gives the same film: {same}")
        out["cases"][case] = dict(tile=tile, rays=[st["camera"], st["closest"], st["shadow"]], nondelta_estimates=st["nondelta"],
                                  radiance_gt10=st["gt10"], same_film_under_the_running_tmax_rule=same,
                                  bounces={f"{k[0]}:{k[1]}": v for k, v in sorted(st["bounce_kinds"].items())},
                                  film=[[[v.hex() for v in p] for p in row] for row in film])
    with open(os.path.join(HERE, "config2_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote config2_golden.json")


if __name__ == "__main__":
    main()
