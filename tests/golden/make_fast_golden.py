"""The library's FAST sampler mode (the HEADLINE mode of bench.py) against its specification: FAST = the reference's renderer,
unchanged, run with the FastStratified sampler of go-pbrt_b200/go/gopbrt/fast_sampler.go (every (pixel, sample) an independent
stream: reseeded generator, Kensler-permuted strata) at tileSize 1.  This file renders exactly that with the independent plain-Python
restatement of the renderer (make_path_golden.py / make_config1_golden.py / make_mixed_golden.py) and a Python reading of the Go
sampler — BASELINE config 1 and the mixed scene (jittered strata) — so that the oracle's and the CUDA path's FAST mode are pinned to
the documented sampler, not only to each other.

    python tests/golden/make_fast_golden.py        # rewrites tests/golden/fast_golden.json
"""
import importlib
import importlib.util
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
_spec = importlib.util.spec_from_file_location("make_mixed_golden", os.path.join(HERE, "make_mixed_golden.py"))
X = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(X)
C, M = X.C, X.M
CASES = ("config1", "mixed")
PARTITIONS = {"config1_fast_rank1of3": ("config1", "fast", 1, 1, 3), "config1_strict_rank2of3": ("config1", "stratified", 4, 2, 3)}   # case, sampler, tile, rank, world
TILE = 1   # a FAST film's additions happen in the order of the reference's tile loop at tileSize 1


def scene_and_integrator(gp, case):
    return C.scene_and_integrator(gp) if case == "config1" else X.scene_and_integrator(gp, "path_stratified")


def render(gp, case):
    sc = C.plain_scene(*scene_and_integrator(gp, case))
    sc["sampler"] = "fast"
    with C.patched(sc):
        return M.render(sc, TILE)


def render_partition(gp, name):
    """one rank's share of a frame split the library's way (samples s % world == rank in FAST mode, tiles t % world == rank in STRICT)"""
    case, sampler, tile, rank, world = PARTITIONS[name]
    sc = C.plain_scene(*scene_and_integrator(gp, case))
    sc.update(sampler=sampler, rank=rank, world=world)
    with C.patched(sc):
        return M.render(sc, tile)


def main():
    gp = importlib.import_module("go-pbrt_b200")
    out = dict(note="made by tests/golden/make_fast_golden.py (plain-Python restatement of the reference renderer + the FastStratified sampler); "
                    "film = [y][x][X, Y, Z, filterWeightSum] as float.hex()", tile=TILE, cases={})
    for case in CASES:
        film, st = render(gp, case)
        print(f"{case}: camera {st['camera']}, closest {st['closest']}, shadow {st['shadow']}, area-light estimates {st['nondelta']}, "
              f"max direct {st['max_direct']:.3f}, roulette tests {st['rr_tests']}")
        assert st["max_direct"] <= 10.0
        out["cases"][case] = dict(rays=[st["camera"], st["closest"], st["shadow"]], nondelta_estimates=st["nondelta"],
                                  film=[[[v.hex() for v in p] for p in row] for row in film])
    out["partitions"] = {}
    for name in PARTITIONS:
        film, st = render_partition(gp, name)
        print(f"{name}: camera {st['camera']}, closest {st['closest']}, shadow {st['shadow']}")
        out["partitions"][name] = dict(rays=[st["camera"], st["closest"], st["shadow"]], film=[[[v.hex() for v in p] for p in row] for row in film])
    with open(os.path.join(HERE, "fast_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote fast_golden.json")


if __name__ == "__main__":
    main()
