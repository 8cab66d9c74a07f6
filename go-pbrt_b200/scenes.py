"""The BASELINE.json configurations as scene builders on the host mirror (pbrt.py).

config1: exactly internal/render/server.go:29-164.
config2-5: the synthetic scenes of SURVEY §8d; the reference has no triangle, so meshes use this backend's Triangle.
Every builder returns (scene, integrator) ready for pbrt.GpuScene / pbrt.Render.
"""
import math

import numpy as np

from . import pbrt as P


class RNG:
    """pbrt.RandomNumberGenerator (pkg/pbrt/rng.go:11-57), the reference's PCG32 *variant*  — host-side
    use only (procedural scene parameters)."""
    M64 = (1 << 64) - 1

    def __init__(self, seq=None):
        self.state, self.inc = 0x853c49e6748fea9b, 0xda3e39cb94b95bdb
        if seq is not None:
            self.SetSequence(seq)

    def SetSequence(self, seed):
        self.state = 0
        self.inc = ((seed << 1) | 1) & self.M64
        self.UniformUInt32()
        self.state = (self.state + 0x853c49e6748fea9b) & self.M64
        self.UniformUInt32()

    def UniformUInt32(self):
        old = self.state
        self.state = (old * 0x5851f42d4c957f2d + self.inc) & self.M64
        xs = (((old >> 18) ^ old) >> 27) & 0xffffffff
        rot = old >> 59
        return ((xs >> rot) | (xs << ((rot + 1) & 31))) & 0xffffffff

    def UniformFloat(self):
        return min(1.0 - 2.0 ** -53, self.UniformUInt32() * 2.3283064365386963e-10)


def centred_screen_window(W, H):
    """Screen window that makes the reference's (buggy) ProjectiveCamera pipeline produce a centred pinhole:
    Transform.Mul multiplies inverses in the wrong order (transform.go:179-184), so RasterToScreen maps
    x -> (x + min.x)*(max.x-min.x)/W and y -> (y + max.y)*(min.y-max.y)/H.  Solving for x_s in [-a,a], y_s in [1,-1]."""
    a = W / H
    return (-W / 2.0, -H / 2.0 - 2.0, 2.0 * a - W / 2.0, -H / 2.0)


def _camera(pos, look, up, fov, W, H, screen_window=None):
    film = P.NewFilm("render.png", (W, H), (0.0, 0.0, 1.0, 1.0), P.NewBoxFilter((1.0, 1.0)), 100.0, 1.0, 1.0)
    c2w = P.LookAt(pos, look, up)
    sw = screen_window or centred_screen_window(W, H)
    cam = P.NewPerspectiveCamera(P.NewAnimatedTransform(c2w, c2w, 0, 1), sw, 0.0, 1.0, 0.0, 20.0, fov, film, None)
    return cam


def config1(W=1920, H=1080, spp=(4, 4)):
    """README sphere scene (internal/render/server.go:29-164)."""
    prims = []
    n = 8
    for k in range(1, n):
        for i in range(3):
            x = y = z = 0.0
            if i == 0:
                x = float(k) / float(n) * 100
                color = P.NewRGBSpectrum(1, 0, 0)
            elif i == 1:
                y = float(k) / float(n) * 100
                color = P.NewRGBSpectrum(0, 1, 0)
            else:
                z = float(k) / float(n) * 100
                color = P.NewRGBSpectrum(0, 0, 1)
            radius = 2.0
            y = max(y, radius / 2)
            sphere = P.NewSphereShape(f"Sphere: {x}, {y}, {z} - MatteMaterial", P.Translate((0.0, 0.0, 0.0)), True, radius)
            xform = P.Translate((x, y, z))
            kd = P.NewConstantSpectrumTexture(color)
            sigma = P.NewConstantFloatTexture(0.0)
            geo = P.NewGeometricPrimitive(sphere, P.NewMatteMaterial(kd, sigma, None))
            prims.append(P.NewTransformedPrimitive(geo, P.NewAnimatedTransform(xform, xform, 0, 1)))
    sigma = P.NewConstantFloatTexture(0.0)
    checker = P.NewCheckerboard2D(P.NewPlanarMapping2D((.2, 0, 0), (0, 0, .2), 0, 0),
                                  P.NewConstantSpectrumTexture(P.NewSpectrum(1.0)),
                                  P.NewConstantSpectrumTexture(P.NewSpectrum(0.18)))
    m = P.NewMatteMaterial(checker, sigma, None)
    disk_xf = P.Translate((0.0, 0.0, 0.0)).Mul(P.RotateX(90))
    disk = P.NewDisk(disk_xf, 0.01, 10000, 0, 360)
    disk2 = P.NewDisk(P.Translate((-50.0, 0.0, -50.0)), 0.01, 10000, 0, 360)
    prims += [P.NewGeometricPrimitive(disk, m), P.NewGeometricPrimitive(disk2, m)]
    agg = P.NewBVH(prims, 2, P.SplitSAH)
    ls = [
        P.NewDistant(P.Translate((-100.0, 100.0, 100.0)), P.NewSpectrum(0.05), (-1.0, 1.0, 1.0)),
        P.NewPoint(P.Translate((50.0, 20.0, 50.0)), None, P.NewSpectrum(100)),
        P.NewPoint(P.Translate((-50.0, 30.0, -50.0)), None, P.NewSpectrum(50)),
        P.NewDiffuseAreaLight(P.Translate((-10.0, 5.0, 20.0)), None, P.NewSpectrum(0.2), 16,
                              P.NewSphereShape("Light Sphere", P.Translate((-10.0, 5.0, 20.0)), False, 5.0), False),
    ]
    scene = P.NewScene(agg, ls)
    film = P.NewFilm("build/render.png", (W, H), (0.0, 0.0, 1.0, 1.0), P.NewBoxFilter((1.0, 1.0)), 100.0, 1.0, 1.0)
    sampler = P.NewStratified(spp[0], spp[1], False, 4)
    cam_xf = P.LookAt((150.0, 150.0, 150.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0))
    cam_xf = cam_xf.Mul(P.RotateY(-30)).Mul(P.RotateX(-30))
    cam = P.NewPerspectiveCamera(P.NewAnimatedTransform(cam_xf, cam_xf, 0, 1), (0.0, 0.0, 1.0, 1.0), 0.0, 1.0, 0, 20, 100, film, None)
    integ = P.NewPath(10, cam, sampler, None, 1, P.Uniform)
    return scene, integ


def _quad(v, idx, a, b, c, d):
    base = len(v)
    v += [a, b, c, d]
    idx += [(base, base + 1, base + 2), (base, base + 2, base + 3)]


def _box(v, idx, centre, size, rot_deg):
    cx, cy, cz = centre
    hx, hy, hz = size[0] / 2, size[1] / 2, size[2] / 2
    c, s = math.cos(math.radians(rot_deg)), math.sin(math.radians(rot_deg))

    def pt(x, y, z):
        return (cx + c * x + s * z, cy + y, cz - s * x + c * z)

    p = [pt(sx * hx, sy * hy, sz * hz) for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]
    # index = 4*ix + 2*iy + iz
    faces = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    for f in faces:
        _quad(v, idx, p[f[0]], p[f[1]], p[f[2]], p[f[3]])


def config2(W=1920, H=1080, spp=(8, 8)):
    """Cornell-box-style room of triangles + a matte and a glass sphere, disk area light."""
    white = P.NewMatteMaterial(P.NewConstantSpectrumTexture(P.NewRGBSpectrum(.73, .73, .73)), P.NewConstantFloatTexture(0.0))
    red = P.NewMatteMaterial(P.NewConstantSpectrumTexture(P.NewRGBSpectrum(.63, .065, .05)), P.NewConstantFloatTexture(0.0))
    green = P.NewMatteMaterial(P.NewConstantSpectrumTexture(P.NewRGBSpectrum(.14, .45, .091)), P.NewConstantFloatTexture(0.0))
    prims = []
    v, idx = [], []
    _quad(v, idx, (-1, -1, -1), (1, -1, -1), (1, -1, 1), (-1, -1, 1))   # floor
    _quad(v, idx, (-1, 1, -1), (-1, 1, 1), (1, 1, 1), (1, 1, -1))       # ceiling
    _quad(v, idx, (-1, -1, -1), (-1, 1, -1), (1, 1, -1), (1, -1, -1))   # back
    prims.append(P.TriangleMesh(v, idx, white))
    v, idx = [], []
    _quad(v, idx, (-1, -1, -1), (-1, -1, 1), (-1, 1, 1), (-1, 1, -1))   # left (red)
    prims.append(P.TriangleMesh(v, idx, red))
    v, idx = [], []
    _quad(v, idx, (1, -1, -1), (1, 1, -1), (1, 1, 1), (1, -1, 1))       # right (green)
    prims.append(P.TriangleMesh(v, idx, green))
    v, idx = [], []
    # the boxes float 1 mm above the floor: coplanar faces would make closest-hit results depend on BVH visit order
    # (ties inside the efloat bound), which no two tree topologies can agree on
    _box(v, idx, (0.35, -0.699, 0.3), (0.6, 0.6, 0.6), -18.0)          # short box
    _box(v, idx, (-0.35, -0.399, -0.3), (0.6, 1.2, 0.6), 18.0)         # tall box
    prims.append(P.TriangleMesh(v, idx, white))
    matte = P.NewMatteMaterial(P.NewConstantSpectrumTexture(P.NewRGBSpectrum(.25, .35, .75)), P.NewConstantFloatTexture(0.0))
    glass = P.NewGlass(P.NewConstantSpectrumTexture(P.NewSpectrum(1.0)), P.NewConstantSpectrumTexture(P.NewSpectrum(1.0)),
                       P.NewConstantFloatTexture(0.0), P.NewConstantFloatTexture(0.0), P.NewConstantFloatTexture(1.5))
    prims.append(P.NewGeometricPrimitive(P.NewSphereShape("matte", P.Translate((0.35, -0.098, 0.3)), False, 0.3), matte))
    prims.append(P.NewGeometricPrimitive(P.NewSphereShape("glass", P.Translate((-0.45, -0.7, 0.6)), False, 0.3), glass))
    agg = P.NewBVH(prims, 4, P.SplitSAH)
    light_xf = P.Translate((0.0, 0.99, 0.0)).Mul(P.RotateX(90))
    ls = [P.NewDiffuseAreaLight(light_xf, None, P.NewRGBSpectrum(17, 12, 4), 1, P.NewDisk(light_xf, 0.0, 0.25, 0, 360), False)]
    scene = P.NewScene(agg, ls)
    cam = _camera((0.0, 0.0, 3.6), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 40.0, W, H)
    integ = P.NewPath(10, cam, P.NewStratified(spp[0], spp[1], False, 4), None, 1, P.Uniform)
    return scene, integ


def config3(W=1920, H=1080, spp=(16, 16), n_spheres=100000):
    """Random sphere field with mixed materials and sphere area lights."""
    rng = RNG(0xC0FFEE)
    U = rng.UniformFloat
    mirror = P.NewMirror()
    glass = P.NewGlass(P.NewConstantSpectrumTexture(P.NewSpectrum(1.0)), P.NewConstantSpectrumTexture(P.NewSpectrum(1.0)),
                       P.NewConstantFloatTexture(0.0), P.NewConstantFloatTexture(0.0), P.NewConstantFloatTexture(1.5))
    zero = P.NewConstantFloatTexture(0.0)
    prims = []
    for i in range(n_spheres):
        c = (-50 + 100 * U(), -50 + 100 * U(), -50 + 100 * U())
        r = 0.1 + 0.4 * U()
        u = U()
        if u < 0.70:
            kd = P.NewRGBSpectrum(0.1 + 0.8 * U(), 0.1 + 0.8 * U(), 0.1 + 0.8 * U())
            m = P.NewMatteMaterial(P.NewConstantSpectrumTexture(kd), zero)
        elif u < 0.85:
            m = mirror
        else:
            m = glass
        prims.append(P.NewGeometricPrimitive(P.NewSphereShape("s", P.Translate(c), False, r), m))
    agg = P.NewBVH(prims, 4, P.SplitSAH)
    ls = []
    for k in range(8):
        a = 2 * math.pi * k / 8
        pos = (70 * math.cos(a), 60.0, 70 * math.sin(a))
        xf = P.Translate(pos)
        ls.append(P.NewDiffuseAreaLight(xf, None, P.NewSpectrum(5.0), 1, P.NewSphereShape("l", xf, False, 2.0), False))
    ls.append(P.NewDistant(P.Translate((0.0, 0.0, 0.0)), P.NewSpectrum(0.5), (-1.0, 1.0, 1.0)))
    scene = P.NewScene(agg, ls)
    cam = _camera((120.0, 80.0, 120.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 45.0, W, H)
    integ = P.NewPath(10, cam, P.NewStratified(spp[0], spp[1], False, 4), None, 1, P.Uniform)
    return scene, integ


def heightfield_mesh(grid=2237, extent=50.0, seed=0xBEEF):
    """(grid-1)^2*2 triangles over [-extent, extent]^2, y = sum_k a_k sin(f_k x + p_k) sin(g_k z + q_k)."""
    rng = RNG(seed)
    U = rng.UniformFloat
    xs = np.linspace(-extent, extent, grid)
    X, Z = np.meshgrid(xs, xs, indexing="xy")
    Y = np.zeros_like(X)
    for k in range(1, 7):
        a = (0.5 + 2.5 * U()) / k
        f, g = 0.05 * k + 0.1 * U(), 0.05 * k + 0.1 * U()
        p, q = 2 * math.pi * U(), 2 * math.pi * U()
        Y += a * np.sin(f * X + p) * np.sin(g * Z + q)
    verts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    i = np.arange(grid - 1)
    I, J = np.meshgrid(i, i, indexing="xy")
    v00 = (J * grid + I).ravel().astype(np.int32)
    v10, v01, v11 = v00 + 1, v00 + grid, v00 + grid + 1
    tris = np.empty((2 * len(v00), 3), dtype=np.int32)
    tris[0::2] = np.stack([v00, v01, v10], axis=1)
    tris[1::2] = np.stack([v10, v01, v11], axis=1)
    return verts, tris


def config4(W=1920, H=1080, spp=(4, 4), grid=2237):
    """Tessellated heightfield: 2236^2*2 = 9 999 392 triangles at grid=2237."""
    verts, tris = heightfield_mesh(grid)
    matte = P.NewMatteMaterial(P.NewConstantSpectrumTexture(P.NewSpectrum(0.6)), P.NewConstantFloatTexture(0.0))
    agg = P.NewBVH([P.TriangleMesh(verts, tris, matte)], 4, P.SplitSAH)
    ls = [P.NewPoint(P.Translate((30.0, 40.0, 30.0)), None, P.NewSpectrum(3000)),
          P.NewPoint(P.Translate((-40.0, 30.0, -20.0)), None, P.NewSpectrum(2000)),
          P.NewDistant(P.Translate((0.0, 0.0, 0.0)), P.NewSpectrum(1.0), (0.3, 1.0, 0.2))]
    scene = P.NewScene(agg, ls)
    cam = _camera((0.0, 15.0, 80.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 50.0, W, H)
    integ = P.NewPath(10, cam, P.NewStratified(spp[0], spp[1], False, 4), None, 1, P.Uniform)
    return scene, integ


def config4_integrator(W, H, spp=(4, 4)):
    """config 4's camera and Path integrator at another resolution / sample count (for a scene that is already built)"""
    cam = _camera((0.0, 15.0, 80.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 50.0, W, H)
    return P.NewPath(10, cam, P.NewStratified(spp[0], spp[1], False, 4), None, 1, P.Uniform)


def config5(W=3840, H=2160, spp=(32, 32), grid=2237):
    """config-4 mesh at 4K, 1023 effective spp, meant to be split across 2/4/8 GPUs."""
    return config4(W, H, spp, grid)


def bvh_test_primitives():
    """pkg/accelerator/simple_test.go:10-38 — the three unit spheres of the reference's own accelerator tests."""
    return [
        P.NewGeometricPrimitive(P.NewSphereShape("prim1", P.Translate((0.0, 0.0, 5.0)), False, 1.0), None),
        P.NewGeometricPrimitive(P.NewSphereShape("prim2", P.Translate((0.0, 0.0, 10.0)), False, 1.0), None),
        P.NewGeometricPrimitive(P.NewSphereShape("prim3", P.Translate((10.0, 10.0, 10.0)), True, 1.0), None),
    ]


def mixed_test_scene(n=200, seed=7, with_triangles=True):
    """A small random scene touching every shape/material/light kind, for oracle-vs-GPU parity tests."""
    rng = RNG(seed)
    U = rng.UniformFloat
    zero = P.NewConstantFloatTexture(0.0)
    mirror = P.NewMirror()
    glass = P.NewGlass(P.NewConstantSpectrumTexture(P.NewSpectrum(0.9)), P.NewConstantSpectrumTexture(P.NewSpectrum(0.95)),
                       zero, zero, P.NewConstantFloatTexture(1.5))
    checker = P.NewCheckerboard2D(P.NewPlanarMapping2D((.5, 0, 0), (0, 0, .5), 0.1, 0.2),
                                  P.NewConstantSpectrumTexture(P.NewSpectrum(0.8)),
                                  P.NewConstantSpectrumTexture(P.NewRGBSpectrum(0.1, 0.2, 0.3)))
    uvchecker = P.NewCheckerboard2D(P.NewUvMapping2D(8.0, 8.0, 0.0, 0.0), P.NewConstantSpectrumTexture(P.NewSpectrum(0.7)),
                                    P.NewConstantSpectrumTexture(P.NewRGBSpectrum(0.7, 0.2, 0.1)))
    prims = []
    for i in range(n):
        c = (-10 + 20 * U(), 0.3 + 6 * U(), -10 + 20 * U())
        r = 0.2 + 0.8 * U()
        u = U()
        if u < 0.4:
            m = P.NewMatteMaterial(P.NewConstantSpectrumTexture(P.NewRGBSpectrum(0.1 + 0.8 * U(), 0.1 + 0.8 * U(), 0.1 + 0.8 * U())), zero)
        elif u < 0.55:
            m = P.NewMatteMaterial(uvchecker, P.NewConstantFloatTexture(20.0))  # Oren-Nayar branch
        elif u < 0.75:
            m = mirror
        else:
            m = glass
        sph = P.NewSphereShape("s", P.Translate(c) if i % 3 else P.Translate((0.0, 0.0, 0.0)), bool(i % 5 == 0), r)
        g = P.NewGeometricPrimitive(sph, m)
        if i % 3 == 0:
            # NewTransform re-derives a consistent inverse: Transform.Mul's own inverse is wrong for non-commuting
            # factors (transform.go:179-184), which would put the sphere outside its own world bound
            xf = P.NewTransform(P.Translate(c).Mul(P.RotateY(30.0 * U())).Matrix)
            g = P.NewTransformedPrimitive(g, P.NewAnimatedTransform(xf, xf, 0, 1))
        prims.append(g)
    floor = P.NewDisk(P.Translate((0.0, 0.0, 0.0)).Mul(P.RotateX(90)), 0.0, 40.0, 0, 360)
    prims.append(P.NewGeometricPrimitive(floor, P.NewMatteMaterial(checker, zero)))
    if with_triangles:
        v, idx = [], []
        _box(v, idx, (3.0, 1.01, 2.0), (2.0, 2.0, 2.0), 25.0)  # 1 cm above the floor disk: no coplanar ties
        _quad(v, idx, (-12, 0, -12), (-12, 9, -12), (12, 9, -12), (12, 0, -12))
        prims.append(P.TriangleMesh(v, idx, P.NewMatteMaterial(P.NewConstantSpectrumTexture(P.NewSpectrum(0.6)), zero)))
    agg = P.NewBVH(prims, 4, P.SplitSAH)
    lxf = P.Translate((0.0, 14.0, 0.0))
    dxf = P.Translate((4.0, 12.0, -3.0)).Mul(P.RotateX(90))
    ls = [P.NewDistant(P.Translate((0.0, 0.0, 0.0)), P.NewSpectrum(0.3), (-1.0, 1.0, 1.0)),
          P.NewPoint(P.Translate((5.0, 9.0, 5.0)), None, P.NewSpectrum(60)),
          P.NewDiffuseAreaLight(lxf, None, P.NewSpectrum(3.0), 1, P.NewSphereShape("L", lxf, False, 2.0), False),
          P.NewDiffuseAreaLight(dxf, None, P.NewRGBSpectrum(8, 7, 6), 1, P.NewDisk(dxf, 0.0, 1.5, 0, 360), True)]
    return P.NewScene(agg, ls)


def test_integrator(W, H, spp=(2, 2), pos=(18.0, 9.0, 18.0), look=(0.0, 2.0, 0.0), fov=50.0, maxDepth=6, jitter=False, ndims=4):
    cam = _camera(pos, look, (0.0, 1.0, 0.0), fov, W, H)
    return P.NewPath(maxDepth, cam, P.NewStratified(spp[0], spp[1], jitter, ndims), None, 1, P.Uniform)
