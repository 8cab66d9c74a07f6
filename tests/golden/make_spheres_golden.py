"""A cluster of matte, mirror and glass SPHERES (config 3's ingredients: GeometricPrimitives without a TransformedPrimitive, BVH
leaves of up to 4 primitives, sphere area lights + a distant light) through the independent plain-Python restatement of
make_config1_golden.py / make_path_golden.py.  New against config 1: rays that start INSIDE a sphere (the quadratic's second root,
sphere.go:88-94), FresnelSpecular on curved surfaces with wo = the ray's direction (both `etaScale` branches, total internal
reflection ending a path), mirror spheres, several area lights under one UniformSampleOneLight draw.

    python tests/golden/make_spheres_golden.py        # rewrites tests/golden/spheres_golden.json
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
W, H, SPP, TILE, MAX_DEPTH = 20, 14, (3, 3), 7, 8


def scene_and_integrator(gp):
    P, S = gp.pbrt, gp.scenes
    rng = S.RNG(0x5EED)
    U = rng.UniformFloat
    zero = P.NewConstantFloatTexture(0.0)
    mirror = P.NewMirror()
    glass = P.NewGlass(P.NewConstantSpectrumTexture(P.NewSpectrum(1.0)), P.NewConstantSpectrumTexture(P.NewRGBSpectrum(1.0, 0.9, 0.8)),
                       zero, zero, P.NewConstantFloatTexture(1.5))
    prims = []
    for i in range(40):
        c = (-6 + 12 * U(), -4 + 8 * U(), -6 + 12 * U())
        r = 0.6 + 1.0 * U()
        u = U()
        if u < 0.45:
            m = P.NewMatteMaterial(P.NewConstantSpectrumTexture(P.NewRGBSpectrum(0.1 + 0.8 * U(), 0.1 + 0.8 * U(), 0.1 + 0.8 * U())), zero)
        elif u < 0.65:
            m = mirror
        else:
            m = glass
        prims.append(P.NewGeometricPrimitive(P.NewSphereShape("s", P.Translate(c), False, r), m))
    # a matte ground disk well below the cluster (one elementary transform: see make_path_golden.scene_and_integrator)
    prims.append(P.NewGeometricPrimitive(P.NewDisk(P.RotateX(90), 7.5, 40.0, 0.0, 360),
                                         P.NewMatteMaterial(P.NewConstantSpectrumTexture(P.NewSpectrum(0.6)), zero)))
    agg = P.NewBVH(prims, 4, P.SplitSAH)
    ls = []
    for pos in ((14.0, 12.0, 3.0), (-9.0, 14.0, 10.0), (2.0, 13.0, -12.0)):
        xf = P.Translate(pos)
        ls.append(P.NewDiffuseAreaLight(xf, None, P.NewSpectrum(6.0), 1, P.NewSphereShape("l", xf, False, 2.0), False))
    ls.append(P.NewDistant(P.Translate((0.0, 0.0, 0.0)), P.NewSpectrum(0.4), (-1.0, 1.0, 1.0)))
    scene = P.NewScene(agg, ls)
    cam = S._camera((17.0, 9.0, 17.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 45.0, W, H)
    integ = P.NewPath(MAX_DEPTH, cam, P.NewStratified(SPP[0], SPP[1], True, 4), None, 1, P.Uniform)
    return scene, integ


def main():
    gp = importlib.import_module("go-pbrt_b200")
    scene, integ = scene_and_integrator(gp)
    sc = C.plain_scene(scene, integ)
    film, st = C.render(sc, TILE)
    lit = sum(1 for row in film for p in row if p[1] > 0)
    print(f"sphere cluster at {W}x{H}, tile {TILE}: camera {st['camera']}, closest {st['closest']}, shadow {st['shadow']}, area-light estimates "
          f"{st['nondelta']}, lit pixels {lit}/{W * H}, max direct {st['max_direct']:.3f}, bounces {st['bounce_kinds']}, roulette tests {st['rr_tests']}")
    assert st["max_direct"] <= 10.0
    out = dict(note="made by tests/golden/make_spheres_golden.py (plain-Python restatement of the hot path on a cluster of matte / mirror / glass "
                    "spheres); film = [y][x][X, Y, Z, filterWeightSum] as float.hex()",
               width=W, height=H, spp=list(SPP), tile=TILE, rays=[st["camera"], st["closest"], st["shadow"]], nondelta_estimates=st["nondelta"],
               coverage=dict(russian_roulette_tests=st["rr_tests"], bounces={f"{k[0]}:{k[1]}": v for k, v in sorted(st["bounce_kinds"].items())}),
               film=[[[v.hex() for v in p] for p in row] for row in film])
    with open(os.path.join(HERE, "spheres_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote spheres_golden.json")


if __name__ == "__main__":
    main()
