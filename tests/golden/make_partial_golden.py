"""Partial shapes and the UV mapping through the independent plain-Python restatement (make_config1_golden.py): spheres clipped in z
and phi (sphere.go:108-132 — the retry with the second root, `tShapeHit == t1`, and the inner `phi :=` that leaves u on the FIRST
root's phi), disks with an inner radius and a phiMax (disk.go:76-89), thetaMin / thetaMax in dpdv, Checkerboard2D over UVMapping2D
(texture.go:9-26: needs Go's Atan2 and Acos bit for bit — restated in make_path_golden.py) — under Path, Stratified 3x3.

    python tests/golden/make_partial_golden.py        # rewrites tests/golden/partial_golden.json
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
W, H, SPP, TILE, MAX_DEPTH = 20, 14, (3, 3), 6, 6


def scene_and_integrator(gp):
    P, S = gp.pbrt, gp.scenes
    zero = P.NewConstantFloatTexture(0.0)

    def uv_checker(su, sv, a, b):
        return P.NewMatteMaterial(P.NewCheckerboard2D(P.UVMapping2D(su, sv, 0.0, 0.0), P.NewConstantSpectrumTexture(P.NewRGBSpectrum(*a)),
                                                      P.NewConstantSpectrumTexture(P.NewRGBSpectrum(*b))), zero)

    def matte(r, g, b):
        return P.NewMatteMaterial(P.NewConstantSpectrumTexture(P.NewRGBSpectrum(r, g, b)), zero)

    glass = P.NewGlass(P.NewConstantSpectrumTexture(P.NewSpectrum(1.0)), P.NewConstantSpectrumTexture(P.NewSpectrum(0.95)), zero, zero,
                       P.NewConstantFloatTexture(1.5))
    prims = [
        # a bowl: the lower part of a sphere, seen from above — camera rays enter through the cut and meet the INSIDE (second root)
        P.NewGeometricPrimitive(P.Sphere("bowl", P.RotateX(-90), False, 3.0, -3.0, 0.5, 360.0), uv_checker(8.0, 6.0, (0.9, 0.9, 0.9), (0.8, 0.2, 0.2))),
        # a wedge: three quarters in phi, both caps cut
        P.NewGeometricPrimitive(P.Sphere("wedge", P.Translate((5.5, 1.0, -1.0)), False, 2.0, -1.2, 1.5, 270.0), uv_checker(6.0, 4.0, (0.2, 0.7, 0.9), (0.9, 0.8, 0.2))),
        # a glass dome with reversed orientation
        P.NewGeometricPrimitive(P.Sphere("dome", P.Translate((-5.0, -1.0, 2.0)), True, 2.2, -0.4, 2.2, 300.0), glass),
        # a mirror band
        P.NewGeometricPrimitive(P.Sphere("band", P.Translate((0.5, 0.5, -6.5)), False, 2.5, -1.0, 1.0, 360.0), P.NewMirror()),
        # an annulus sector above the bowl and a half disk
        P.NewGeometricPrimitive(P.NewDisk(P.RotateX(-90), 4.0, 2.5, 1.0, 250.0), uv_checker(5.0, 3.0, (0.9, 0.5, 0.1), (0.1, 0.3, 0.6))),
        P.NewGeometricPrimitive(P.NewDisk(P.Translate((4.0, 3.0, 4.0)), 0.5, 2.0, 0.0, 180.0), matte(0.4, 0.8, 0.4)),
        # the ground (plane y = -3.2), UV-mapped
        P.NewGeometricPrimitive(P.NewDisk(P.RotateX(90), 3.2, 40.0, 0.0, 360), uv_checker(24.0, 40.0, (0.7, 0.7, 0.7), (0.25, 0.25, 0.3))),
    ]
    lights = [P.NewPoint(P.Translate((2.0, 9.0, 3.0)), None, P.NewSpectrum(60.0)),
              P.NewDistant(P.Translate((0.0, 0.0, 0.0)), P.NewSpectrum(0.5), (0.3, 1.0, 0.6)),
              P.NewPoint(P.Translate((-6.0, 5.0, -5.0)), None, P.NewRGBSpectrum(30.0, 25.0, 20.0))]
    scene = P.NewScene(P.NewBVH(prims, 2, P.SplitSAH), lights)
    cam = S._camera((11.0, 10.0, 12.0), (0.0, -0.5, 0.0), (0.0, 1.0, 0.0), 50.0, W, H)
    return scene, P.NewPath(MAX_DEPTH, cam, P.NewStratified(SPP[0], SPP[1], True, 4), None, 1.0, P.Uniform)


def main():
    gp = importlib.import_module("go-pbrt_b200")
    sc = C.plain_scene(*scene_and_integrator(gp))
    film, st = C.render(sc, TILE)
    lit = sum(1 for row in film for p in row if p[1] > 0)
    print(f"partial shapes at {W}x{H}, tile {TILE}: camera {st['camera']}, closest {st['closest']}, shadow {st['shadow']}, lit pixels {lit}/{W * H}, "
          f"max direct {st['max_direct']:.3f}, bounces {st['bounce_kinds']}, roulette tests {st['rr_tests']}, second-root retries {C.RETRIES}")
    assert st["max_direct"] <= 10.0
    out = dict(note="made by tests/golden/make_partial_golden.py (plain-Python restatement of the hot path: partial spheres and disks, UV-mapped "
                    "checkerboards); film = [y][x][X, Y, Z, filterWeightSum] as float.hex()",
               width=W, height=H, spp=list(SPP), tile=TILE, rays=[st["camera"], st["closest"], st["shadow"]], nondelta_estimates=st["nondelta"],
               coverage=dict(second_root_retries=dict(C.RETRIES), bounces={f"{k[0]}:{k[1]}": v for k, v in sorted(st["bounce_kinds"].items())}),
               film=[[[v.hex() for v in p] for p in row] for row in film])
    with open(os.path.join(HERE, "partial_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote partial_golden.json")


if __name__ == "__main__":
    main()
