"""Everything else the hot path can meet in one scene, through the independent plain-Python restatement (make_path_golden.py,
make_direct_golden.py, make_config1_golden.py): Oren-Nayar surfaces (matte.go:31-36, reflection.go:616-652, OrenNayar's own `b`),
TransformedPrimitives that ROTATE (primitive.go:94-115: the second TransformSurfaceInteraction now changes normals and dpdu),
reverseOrientation spheres next to plain ones, a thin-lens camera (camera.go:199-211), a checkerboard ground, all four light kinds
in one light list — rendered three ways:
  path_stratified   Path maxDepth 8, Stratified 3x3 jittered, 3 sampled dimensions
  path_random       Path maxDepth 8, RandomSampler(7) (random.go: every draw from the tile's RNG, nothing consumed per pixel)
  path_power        Path with lightSampleStrategy Power: the reference's power distribution is all zeros (lightdistribution.go:58-68,
                    spectrum.go:227-229), no light is ever sampled — a black film whose RAY COUNTS pin the sampler draws that remain
  path_crop         Path, Stratified 2x3 unjittered, a crop window (film.go:42-46), a box filter of radius (1.5, 0.75), tileSize 3
  direct_all        DirectLighting(UniformSampleAll) maxDepth 5, Stratified 3x3
  direct_one        DirectLighting(UniformSampleOne) maxDepth 5, Stratified 3x3

    python tests/golden/make_mixed_golden.py        # rewrites tests/golden/mixed_golden.json
"""
import importlib
import importlib.util
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


D = _load("make_direct_golden")
M = D.M                                   # ONE make_path_golden instance: the patches below must land in the module D renders with
C = _load("make_config1_golden")
C.M = M                                   # ... and in the one C patches
C.K, C.Z3, C.INF = M.K, M.Z3, M.INF
W, H, TILE = 18, 12, 5
CASES = ("path_stratified", "path_random", "path_power", "path_crop", "direct_all", "direct_one")
TILES = {"path_crop": 3}   # every other case: TILE


def scene(gp, point_light_scale=1.0):
    P, S = gp.pbrt, gp.scenes
    rng = S.RNG(0xA11CE)
    U = rng.UniformFloat
    zero = P.NewConstantFloatTexture(0.0)
    glass = P.NewGlass(P.NewConstantSpectrumTexture(P.NewSpectrum(1.0)), P.NewConstantSpectrumTexture(P.NewRGBSpectrum(0.9, 1.0, 0.9)),
                       zero, zero, P.NewConstantFloatTexture(1.4))
    prims = []
    for i in range(24):
        c = (-5 + 10 * U(), -2.5 + 6 * U(), -5 + 10 * U())
        r = 0.5 + 0.9 * U()
        u = U()
        if u < 0.3:
            m = P.NewMatteMaterial(P.NewConstantSpectrumTexture(P.NewRGBSpectrum(0.2 + 0.7 * U(), 0.2 + 0.7 * U(), 0.2 + 0.7 * U())), zero)
        elif u < 0.6:
            m = P.NewMatteMaterial(P.NewConstantSpectrumTexture(P.NewRGBSpectrum(0.2 + 0.7 * U(), 0.2 + 0.7 * U(), 0.2 + 0.7 * U())),
                                   P.NewConstantFloatTexture(10.0 + 70.0 * U()))
        elif u < 0.75:
            m = P.NewMirror()
        else:
            m = glass
        geo = P.NewGeometricPrimitive(P.NewSphereShape("s", P.Translate(c), i % 3 == 0, r), m)
        if i % 2 == 0:   # one elementary transform per Transform (see make_path_golden.scene_and_integrator): the sphere's own
            xf = (P.RotateY(25.0 * (i % 5) - 40.0), P.RotateX(15.0 * (i % 4) - 20.0), P.Translate((0.3 * i - 3.0, 0.5, 0.0)))[i % 3]
            geo = P.NewTransformedPrimitive(geo, P.NewAnimatedTransform(xf, xf, 0, 1))   # translation, then the primitive's rotation
        prims.append(geo)
    checker = P.NewCheckerboard2D(P.NewPlanarMapping2D((0.5, 0, 0), (0, 0, 0.5), 0.25, 0.0),
                                  P.NewConstantSpectrumTexture(P.NewSpectrum(0.8)), P.NewConstantSpectrumTexture(P.NewRGBSpectrum(0.2, 0.3, 0.1)))
    prims.append(P.NewGeometricPrimitive(P.NewDisk(P.RotateX(90), 4.5, 30.0, 0.0, 360), P.NewMatteMaterial(checker, P.NewConstantFloatTexture(35.0))))
    sxf, dxf = P.Translate((9.0, 9.0, -4.0)), P.Translate((-3.0, 1.0, 9.5))
    lights = [P.NewDistant(P.Translate((0.0, 0.0, 0.0)), P.NewSpectrum(0.3), (-1.0, 1.0, 1.0)),
              P.NewDiffuseAreaLight(sxf, None, P.NewSpectrum(7.0), 1, P.NewSphereShape("l", sxf, False, 1.5), False),
              P.NewPoint(P.Translate((4.0, 8.0, 6.0)), None, P.NewRGBSpectrum(40.0 * point_light_scale, 35.0 * point_light_scale, 30.0 * point_light_scale)),
              P.NewDiffuseAreaLight(dxf, None, P.NewRGBSpectrum(5.0, 6.0, 7.0), 1, P.NewDisk(dxf, 0.0, 2.0, 0.0, 360), True)]
    return P.NewScene(P.NewBVH(prims, 3, P.SplitSAH), lights)


def camera(gp, crop=(0.0, 0.0, 1.0, 1.0), radius=(1.0, 1.0)):
    P, S = gp.pbrt, gp.scenes
    film = P.NewFilm("mixed.png", (W, H), crop, P.NewBoxFilter(radius), 100.0, 1.0, 1.0)
    c2w = P.LookAt((13.0, 7.0, 13.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0))
    return P.NewPerspectiveCamera(P.NewAnimatedTransform(c2w, c2w, 0, 1), S.centred_screen_window(W, H), 0.0, 1.0, 0.35, 18.0, 45.0, film, None)


def scene_and_integrator(gp, case):
    """what the oracle and the CUDA path render"""
    P = gp.pbrt
    if case == "path_stratified":
        return scene(gp), P.NewPath(8, camera(gp), P.NewStratified(3, 3, True, 3), None, 1.0, P.Uniform)
    if case == "path_random":
        return scene(gp), P.NewPath(8, camera(gp), P.NewRandomSampler(7), None, 1.0, P.Uniform)
    if case == "path_power":
        return scene(gp), P.NewPath(8, camera(gp), P.NewStratified(3, 3, True, 3), None, 1.0, P.Power)
    if case == "path_crop":
        return scene(gp), P.NewPath(8, camera(gp, (0.2, 0.1, 0.9, 0.8), (1.5, 0.75)), P.NewStratified(2, 3, False, 4), None, 1.0, P.Uniform)
    strategy = P.UniformSampleAll if case == "direct_all" else P.UniformSampleOne
    return scene(gp), P.NewDirectLighting(strategy, 5, camera(gp), P.NewStratified(3, 3, False, 3), None)


def bright_scene_and_integrator(gp):
    """path_stratified with the point light 150 times brighter: UniformSampleOneLight's result exceeds 10 at many vertices, where the
    reference panics (integrator.go:72-74).  The library counts the event (`radiance_gt10`) and carries on with the value; this case
    (CPU tests only) pins that the oracle does exactly that and nothing else."""
    P = gp.pbrt
    return scene(gp, 150.0), P.NewPath(8, camera(gp), P.NewStratified(3, 3, True, 3), None, 1.0, P.Uniform)


def render_bright(gp):
    sc = C.plain_scene(*bright_scene_and_integrator(gp))
    with C.patched(sc):
        return M.render(sc, TILE)


def plain(gp, case):
    """the same as plain numbers.  RandomSampler(ns) == a pixel sampler of ns samples with NO sampled dimension: StartPixel draws
    nothing (random.go:41-53 loops over empty arrays), every Get1D / Get2D comes from the RNG (random.go:21-27)"""
    P = gp.pbrt
    sc_scene, integ = scene_and_integrator(gp, case)
    twin = integ
    if case == "path_random":
        twin = P.NewPath(8, integ.GetCamera(), P.NewStratified(integ.GetSampler().ns, 1, False, 0), None, 1.0, P.Uniform)
    if case.startswith("direct_"):   # plain_scene reads the Path fields; DirectLighting has no roulette threshold
        twin = P.NewPath(integ.maxDepth, integ.GetCamera(), integ.GetSampler(), None, 0.0, P.Uniform)
    sc = C.plain_scene(sc_scene, twin)
    if case == "path_power":
        sc["light_strategy"] = "power"
    return sc


def render(sc, case):
    with C.patched(sc):
        if case.startswith("direct_"):
            assert sc["max_depth"] == D.MAX_DEPTH
            return D.render(sc, TILE, D.UNIFORM_SAMPLE_ALL if case == "direct_all" else D.UNIFORM_SAMPLE_ONE)
        return M.render(sc, TILES.get(case, TILE))


def main():
    gp = importlib.import_module("go-pbrt_b200")
    out = dict(note="made by tests/golden/make_mixed_golden.py (plain-Python restatement of the hot path: Oren-Nayar, rotating TransformedPrimitives, "
                    "thin lens, Random sampler, all light kinds); film = [y][x][X, Y, Z, filterWeightSum] as float.hex()", width=W, height=H, tile=TILE, cases={})
    for case in CASES:
        sc = plain(gp, case)
        film, st = render(sc, case)
        lit = sum(1 for row in film for p in row if p[1] != 0)
        print(f"{case}: camera {st['camera']}, closest {st['closest']}, shadow {st['shadow']}, area-light estimates {st['nondelta']}, lit pixels "
              f"{lit}/{W * H}, max direct {st['max_direct']:.3f}, bounces {st.get('bounce_kinds')}, transmitted {st.get('spec_transmit_rays')}")
        assert st["max_direct"] <= 10.0 and st["gt10"] == 0
        out["cases"][case] = dict(rays=[st["camera"], st["closest"], st["shadow"]], nondelta_estimates=st["nondelta"],
                                  bounces={f"{k[0]}:{k[1]}": v for k, v in sorted(st.get("bounce_kinds", {}).items())},
                                  transmitted_rays=st.get("spec_transmit_rays", 0),
                                  film=[[[v.hex() for v in p] for p in row] for row in film])
    film, st = render_bright(gp)
    print(f"bright: camera {st['camera']}, closest {st['closest']}, shadow {st['shadow']}, > 10 events {st['gt10']}, max direct {st['max_direct']:.1f}")
    assert st["gt10"] > 20
    out["bright"] = dict(rays=[st["camera"], st["closest"], st["shadow"]], nondelta_estimates=st["nondelta"], radiance_gt10=st["gt10"],
                         film=[[[v.hex() for v in p] for p in row] for row in film])
    with open(os.path.join(HERE, "mixed_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote mixed_golden.json")


if __name__ == "__main__":
    main()
