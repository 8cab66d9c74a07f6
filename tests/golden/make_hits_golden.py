"""Known answers for the batched Aggregate.Intersect / IntersectP drop-ins (gopbrt_trace_closest / gopbrt_trace_any): for camera rays and
for rays leaving the surfaces they hit, the primitive, tHit, hit point and geometric normal (and the any-hit bit for shadow segments
towards a point) as the plain-Python restatement of make_config1_golden.py computes them — BASELINE config 1 (spheres behind
TransformedPrimitives, disks), the partial-shape scene (clipped spheres, partial disks), the sphere cluster (BVH leaves of 4) and the
mixed scene (rotating TransformedPrimitives, reverseOrientation spheres).

    python tests/golden/make_hits_golden.py        # rewrites tests/golden/hits_golden.json
"""
import importlib
import importlib.util
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
_spec = importlib.util.spec_from_file_location("make_partial_golden", os.path.join(HERE, "make_partial_golden.py"))
PG = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(PG)
C = PG.C
M, K = C.M, C.K


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


SG, XG = _load("make_spheres_golden"), _load("make_mixed_golden")
INF = float("inf")
CASES = ("config1", "partial", "spheres", "mixed")
GRID = (16, 10)     # camera rays through a GRID of raster points
TARGET = {"config1": [50.0, 20.0, 50.0], "partial": [2.0, 9.0, 3.0], "spheres": [14.0, 12.0, 3.0], "mixed": [4.0, 8.0, 6.0]}   # the point the shadow segments aim at (a light of each scene)


def scene_and_integrator(gp, case):
    if case == "spheres":
        return SG.scene_and_integrator(gp)
    if case == "mixed":   # rotating TransformedPrimitives, reverseOrientation spheres
        return XG.scene_and_integrator(gp, "path_stratified")
    return C.scene_and_integrator(gp) if case == "config1" else PG.scene_and_integrator(gp)


def closest(prims, o, w, tmax):
    """make_config1_golden.Scene.intersect, keeping the primitive index and tHit"""
    inv, neg = C._inv_dir(w)
    best = None
    for i, pr in enumerate(prims):
        if M.bounds_intersect_p(pr.bound, o, tmax, inv, neg):
            r = pr.intersect(o, w, tmax)
            if r is not None:
                tmax = r[0]
                best = (i, r[0], r[1])
    return best


def any_hit(prims, o, w, tmax):
    inv, neg = C._inv_dir(w)
    return any(M.bounds_intersect_p(pr.bound, o, tmax, inv, neg) and pr.intersect_p(o, w, tmax) for pr in prims)


def rays_and_answers(gp, case):
    sc = C.plain_scene(*scene_and_integrator(gp, case))
    prims = sc["prims"]
    w_img, h_img = sc["res"]
    rays, shadow = [], []
    for j in range(GRID[1]):
        for i in range(GRID[0]):
            p_film = [(i + 0.37) * w_img / GRID[0], (j + 0.61) * h_img / GRID[1]]
            o, w = K.camera_ray(sc["r2c"], sc["c2w"], 0.0, 1.0, p_film, [0.0, 0.0], M.COS, M.SIN)
            rays.append((o, w, INF))
    # second generation: from every hit, the mirror direction about the geometric normal (un-normalised on purpose), tMax 1000;
    # and the shadow segment of SpawnRayToInteraction towards TARGET (origin = the un-offset hit point, tMax = 1 - ShadowEpsilon)
    for o, w, _ in list(rays):
        h = closest(prims, o, w, INF)
        if h is None:
            continue
        rec = h[2]
        d = K.v_sub(w, K.v_muls(rec["n"], 2 * K.v_dot(w, rec["n"])))
        rays.append((K.offset_ray_origin(rec["p"], rec["perr"], rec["n"], d), K.v_muls(d, 3.0), 1000.0))
        so, sw = K.spawn_ray_to(rec["p"], rec["perr"], rec["n"], TARGET[case], [0.0] * 3, [0.0] * 3)
        shadow.append((so, sw, 1 - M.SHADOW_EPSILON))
    hits = []
    for o, w, tmax in rays:
        h = closest(prims, o, w, tmax)
        hits.append(None if h is None else dict(prim=h[0], t=h[1].hex(), p=[v.hex() for v in h[2]["p"]], n=[v.hex() for v in h[2]["n"]]))
    occluded = [any_hit(prims, o, w, tmax) for o, w, tmax in shadow]
    enc = lambda rs: [dict(o=[v.hex() for v in o], d=[v.hex() for v in w], tmax=("inf" if tmax == INF else tmax.hex())) for o, w, tmax in rs]
    return dict(rays=enc(rays), hits=hits, shadow_rays=enc(shadow), occluded=occluded)


def main():
    gp = importlib.import_module("go-pbrt_b200")
    out = dict(note="made by tests/golden/make_hits_golden.py (plain-Python restatement of Aggregate.Intersect / IntersectP); floats as float.hex()", cases={})
    for case in CASES:
        c = rays_and_answers(gp, case)
        nh = sum(h is not None for h in c["hits"])
        print(f"{case}: {len(c['rays'])} rays, {nh} hits on {len({h['prim'] for h in c['hits'] if h})} distinct primitives; "
              f"{len(c['shadow_rays'])} shadow segments, {sum(c['occluded'])} occluded")
        out["cases"][case] = c
    with open(os.path.join(HERE, "hits_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote hits_golden.json")


if __name__ == "__main__":
    main()
