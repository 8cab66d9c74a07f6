"""feature-by-feature GPU-vs-oracle parity probe (debug aid): rays + small renders on micro-scenes"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
gp = importlib.import_module("go-pbrt_b200")
from oracle_lib import OracleScene
P = gp.pbrt; S = gp.scenes
dev = P.Device(0)
zero = P.NewConstantFloatTexture(0.0)
def matte(rgb, sigma=0.0): return P.NewMatteMaterial(P.NewConstantSpectrumTexture(P.NewRGBSpectrum(*rgb)), P.NewConstantFloatTexture(sigma))
glass = P.NewGlass(P.NewConstantSpectrumTexture(P.NewSpectrum(0.9)), P.NewConstantSpectrumTexture(P.NewSpectrum(0.95)), zero, zero, P.NewConstantFloatTexture(1.5))
floor = lambda m: P.NewGeometricPrimitive(P.NewDisk(P.Translate((0.0, 0.0, 0.0)).Mul(P.RotateX(90)), 0.0, 40.0, 0, 360), m)
def tri_box(cy=1.0):
    v, idx = [], []
    S._box(v, idx, (0.0, cy, 0.0), (2.0, 2.0, 2.0), 25.0)
    return P.TriangleMesh(v, idx, matte((.6, .6, .6)))
point = P.NewPoint(P.Translate((5.0, 9.0, 5.0)), None, P.NewSpectrum(60))
distant = P.NewDistant(P.Translate((0.0, 0.0, 0.0)), P.NewSpectrum(0.3), (-1.0, 1.0, 1.0))
lxf = P.Translate((0.0, 8.0, 0.0))
area_s = P.NewDiffuseAreaLight(lxf, None, P.NewSpectrum(3.0), 1, P.NewSphereShape("L", lxf, False, 2.0), False)
dxf = P.Translate((1.0, 7.0, -1.0)).Mul(P.RotateX(90))
area_d = P.NewDiffuseAreaLight(dxf, None, P.NewRGBSpectrum(8, 7, 6), 1, P.NewDisk(dxf, 0.0, 1.5, 0, 360), True)
sph = lambda m, c=(0.0, 1.0, 0.0), rev=False: P.NewGeometricPrimitive(P.NewSphereShape("s", P.Translate(c), rev, 1.0), m)
uvchecker = P.NewCheckerboard2D(P.NewUvMapping2D(8.0, 8.0, 0.0, 0.0), P.NewConstantSpectrumTexture(P.NewSpectrum(0.7)), P.NewConstantSpectrumTexture(P.NewRGBSpectrum(0.7, 0.2, 0.1)))
rot = P.NewTransform(P.Translate((2.5, 1.0, 0.5)).Mul(P.RotateY(20.0)).Matrix)
cases = {
  "lambert+point": ([sph(matte((.5,.6,.7))), floor(matte((.5,.5,.5)))], [point]),
  "lambert+distant": ([sph(matte((.5,.6,.7))), floor(matte((.5,.5,.5)))], [distant]),
  "lambert+area_sphere": ([sph(matte((.5,.6,.7))), floor(matte((.5,.5,.5)))], [area_s]),
  "lambert+area_disk": ([sph(matte((.5,.6,.7))), floor(matte((.5,.5,.5)))], [area_d]),
  "oren_nayar": ([sph(matte((.5,.6,.7), 20.0)), floor(matte((.5,.5,.5), 35.0))], [point]),
  "mirror": ([sph(P.NewMirror()), floor(matte((.5,.5,.5)))], [point]),
  "glass": ([sph(glass), floor(matte((.5,.5,.5)))], [point]),
  "uvchecker": ([sph(P.NewMatteMaterial(uvchecker, zero)), floor(matte((.5,.5,.5)))], [point]),
  "reversed_sphere": ([sph(matte((.5,.6,.7)), rev=True), floor(matte((.5,.5,.5)))], [point]),
  "triangles": ([tri_box(), floor(matte((.5,.5,.5)))], [point]),
  "triangles_only": ([tri_box()], [point]),
  "triangles_lifted": ([tri_box(1.01), floor(matte((.5,.5,.5)))], [point]),
  "rotated_tp": ([P.NewTransformedPrimitive(P.NewGeometricPrimitive(P.NewSphereShape("s", P.Translate((0.0,0.0,0.0)), False, 1.0), matte((.5,.6,.7))), P.NewAnimatedTransform(rot, rot, 0, 1)), floor(matte((.5,.5,.5)))], [point]),
}
only = sys.argv[1:] or list(cases)
for name in only:
    prims, lights = cases[name]
    scene = P.NewScene(P.NewBVH(prims, 4, P.SplitSAH), lights)
    integ = S.test_integrator(96, 64, spp=(3, 3), pos=(6.0, 4.0, 6.0), look=(0.0, 1.0, 0.0), maxDepth=6)
    g = P.GpuScene(dev, scene); o = OracleScene(scene, 1)
    st = P.Render(g, integ, 1); film = integ.GetCamera().GetFilm().pixels.copy()
    ofilm, ost = o.render(integ, 1)
    bad = np.any(film != ofilm, axis=2)
    print(f"{name:22s} pixels differing {bad.sum():5d}/{bad.size}  closest {st['closest_rays']} vs {ost['closest_rays']}  shadow {st['shadow_rays']} vs {ost['shadow_rays']}")
    if bad.any():
        ys, xs = np.nonzero(bad)
        k = 0
        print("   first bad pixel", (xs[k], ys[k]), film[ys[k], xs[k]], ofilm[ys[k], xs[k]])
        rel = np.abs(film - ofilm) / np.maximum(np.abs(ofilm), 1e-300)
        print("   max rel diff", rel[..., :3].max(), " median rel diff over bad", np.median(rel[..., :3][bad]))
    g.close(); o.close()
