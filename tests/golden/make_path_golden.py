"""An INDEPENDENT restatement of the COMPOSED hot path — pbrt.Render -> renderWorker -> Path.Li -> UniformSampleOneLight ->
EstimateDirect -> FilmTile.AddSample -> MergeFilmTile — in plain Python floats, for scenes of disks, matte surfaces and
point lights.

Why: the reference cannot run in this image (no Go toolchain) and holds no golden film.  The per-function known answers
(make_shading_kats.py) pin the functions one by one; what they cannot pin is how Path.Li COMPOSES them (the order of the
sampler draws, `bounces++` first, wo = ray.Direction, the local wi handed to SpawnRay, the dropped DivScalar, the shadow
ray that starts ON the surface, the shared pointers inside TransformSurfaceInteraction, the tile loop's seeds).  The
oracle (oracle/*.h) and the CUDA path were written by the same builder; this file is the third reading of the same Go
source, written from the Go files cited below and from nothing under oracle/ or go-pbrt_b200/csrc/.  It renders two small
films; tests/test_path_golden.py then requires the oracle (CPU) and the CUDA path (GPU, through the C ABI) to reproduce
them BIT FOR BIT, ray counts included.

What it takes from elsewhere, and why that does not weaken it:
 * the per-function restatements of make_shading_kats.py (same independence rule; each is itself pinned as a KAT);
 * the 4x4 matrices of the camera and of the shapes, as DATA, from the host mirror go-pbrt_b200/pbrt.py (LookAt,
   Perspective, RotateX, Inverse — host-side constructors that run once per scene, outside the hot path);
 * sin/cos from go-pbrt_b200/gomath.py (Go's math.Sin/Cos are not libm's; pinned by transform_test.go:77-81).
 Go 1.11's math.Atan2 / Acos (atan.go, atan2.go, asin.go: Cephes' atan polynomial) are restated below.

    python tests/golden/make_path_golden.py        # rewrites tests/golden/path_golden.json
"""
import importlib
import importlib.util
import json
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
_spec = importlib.util.spec_from_file_location("make_shading_kats", os.path.join(HERE, "make_shading_kats.py"))
K = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(K)
gomath = K.gomath
COS, SIN = gomath.Cos, gomath.Sin
INF = float("inf")
SHADOW_EPSILON = 0.0001  # pkg/math/math.go
Z3 = [0.0, 0.0, 0.0]


# ---------------------------------------------------------------- Go 1.11 math: atan.go, asin.go
def go_xatan(x):
    P0, P1, P2, P3, P4 = (-8.750608600031904122785e-01, -1.615753718733365076637e+01, -7.500855792314704667340e+01,
                          -1.228866684490136173410e+02, -6.485021904942025371773e+01)
    Q0, Q1, Q2, Q3, Q4 = (+2.485846490142306297962e+01, +1.650270098316988542046e+02, +4.328810604912902668951e+02,
                          +4.853903996359136964868e+02, +1.945506571482613964425e+02)
    z = x * x
    z = z * ((((P0 * z + P1) * z + P2) * z + P3) * z + P4) / (((((z + Q0) * z + Q1) * z + Q2) * z + Q3) * z + Q4)
    return x * z + x


def go_satan(x):
    morebits, tan3pio8 = 6.123233995736765886130e-17, 2.41421356237309504880
    if x <= 0.66:
        return go_xatan(x)
    if x > tan3pio8:
        return math.pi / 2 - go_xatan(1 / x) + morebits
    return math.pi / 4 + go_xatan((x - 1) / (x + 1)) + 0.5 * morebits


def go_atan(x):  # atan.go: Atan
    if x == 0:
        return x
    if x > 0:
        return go_satan(x)
    return -go_satan(-x)


def go_atan2(y, x):  # atan2.go (finite arguments only)
    assert y == y and x == x and not math.isinf(x) and not math.isinf(y)
    if y == 0:
        if x >= 0 and math.copysign(1.0, x) > 0:
            return math.copysign(0.0, y)
        return math.copysign(math.pi, y)
    if x == 0:
        return math.copysign(math.pi / 2, y)
    q = go_atan(y / x)
    if x < 0:
        if q <= 0:
            return q + math.pi
        return q - math.pi
    return q


def go_asin(x):
    if x == 0:
        return x
    sign = False
    if x < 0:
        x, sign = -x, True
    if x > 1:
        return float("nan")
    temp = math.sqrt(1 - x * x)
    if x > 0.7:
        temp = math.pi / 2 - go_satan(temp / x)
    else:
        temp = go_satan(x / temp)
    return -temp if sign else temp


def go_acos(x):
    return math.pi / 2 - go_asin(x)


# ---------------------------------------------------------------- the scene, shared with the tests (host-mirror objects)
def scene_and_integrator(gp):
    """Four matte disks (floor, ceiling, an annulus at object height 0.25, one tilted by RotateX), a mirror disk (mirror.go: a
    SpecularReflection lobe TYPED Reflection|Diffuse), two glass disks (glass.go: one FresnelSpecular lobe), two point lights and two
    two-sided disk area lights (lights only: the reference never puts a light's shape into the aggregate by itself);
    16x12 pixels, Stratified 3x3 with jitter, 2 sampled dimensions (every later draw comes from the tile's RNG), Path
    maxDepth 6, rrThreshold 1 (Russian roulette is live from the fourth bounce on)."""
    P, S = gp.pbrt, gp.scenes
    zero = P.NewConstantFloatTexture(0.0)

    def matte(r, g, b):
        return P.NewMatteMaterial(P.NewConstantSpectrumTexture(P.NewRGBSpectrum(r, g, b)), zero)

    prims = [
        P.NewGeometricPrimitive(P.NewDisk(P.Translate((0.0, 0.0, 0.0)), 0.0, 6.0, 0.0, 360), matte(0.7, 0.6, 0.5)),
        P.NewGeometricPrimitive(P.NewDisk(P.Translate((0.5, -0.3, 5.0)), 0.0, 7.0, 0.0, 360), matte(0.8, 0.8, 0.8)),
        P.NewGeometricPrimitive(P.NewDisk(P.Translate((1.0, 0.5, 2.0)), 0.25, 1.5, 0.5, 360), matte(0.3, 0.6, 0.9)),
        P.NewGeometricPrimitive(P.NewDisk(P.RotateX(-40), 3.2, 1.4, 0.0, 360), matte(0.9, 0.4, 0.3)),
        P.NewGeometricPrimitive(P.NewDisk(P.RotateY(65), 2.2, 0.9, 0.0, 360), P.NewMirror()),
        # the two glass disks turn their geometric normal (object -z) towards the camera / upwards: DirectLighting's transmission
        # lobe reads `entering` off si.Wo, and from the back it would be past the critical angle
        P.NewGeometricPrimitive(P.NewDisk(P.RotateX(240), -3.0, 1.3, 0.0, 360),
                                P.NewGlass(P.NewConstantSpectrumTexture(P.NewSpectrum(1.0)), P.NewConstantSpectrumTexture(P.NewRGBSpectrum(0.9, 1.0, 0.8)),
                                           zero, zero, P.NewConstantFloatTexture(1.5))),
        P.NewGeometricPrimitive(P.NewDisk(P.RotateX(180), -0.3, 2.5, 0.0, 360),
                                P.NewGlass(P.NewConstantSpectrumTexture(P.NewSpectrum(1.0)), P.NewConstantSpectrumTexture(P.NewRGBSpectrum(0.8, 0.9, 1.0)),
                                           zero, zero, P.NewConstantFloatTexture(1.33))),
    ]
    # every shape transform is ONE elementary transform: Transform.Mul multiplies the inverses in the same order as the matrices
    # (transform.go:179-184), so a product of non-commuting transforms carries a wrong inverse and the shape's bound (from m) and
    # its intersection test (through m^-1) no longer meet — reproduced by every implementation here, but it would leave slivers
    axf, bxf = P.Translate((-0.5, 0.8, 4.2)), P.Translate((0.5, -0.5, -1.0))
    lights = [P.NewPoint(P.Translate((0.3, 0.2, 4.0)), None, P.NewSpectrum(25.0)),
              P.NewDiffuseAreaLight(axf, None, P.NewRGBSpectrum(6.0, 5.0, 4.0), 1, P.NewDisk(axf, 0.0, 2.0, 0.0, 360), True),
              P.NewPoint(P.Translate((-3.0, -2.0, 1.0)), None, P.NewRGBSpectrum(8.0, 6.0, 4.0)),
              # below the floor: it lights nothing (no transmission), but the floor's BSDF-sampling leg (a LOCAL wi, pointing down) finds it
              P.NewDiffuseAreaLight(bxf, None, P.NewRGBSpectrum(3.0, 3.0, 3.0), 1, P.NewDisk(bxf, 0.0, 3.0, 0.0, 360), True)]
    scene = P.NewScene(P.NewBVH(prims, 1, P.SplitSAH), lights)
    W, H = 16, 12
    cam = S._camera((7.0, -6.0, 3.5), (0.0, 0.0, 1.5), (0.0, 0.0, 1.0), 50.0, W, H)
    integ = P.NewPath(6, cam, P.NewStratified(3, 3, True, 2), None, 1.0, P.Uniform)
    return scene, integ


TILE_SIZES = (1, 5)   # 5 leaves ragged tiles on both axes (16 = 3*5 + 1, 12 = 2*5 + 2) and gives tiles more than one pixel


def plain_scene(scene, integ):
    """the same scene as plain numbers"""
    disks = []
    for gpr in scene.aggregate.primitives:
        sh, mt = gpr.Shape, gpr.material
        assert type(sh).__name__ == "Disk"
        kind = type(mt).__name__
        if kind == "MatteMaterial":
            mat = dict(kind="matte", kd=[K.clamp(c, 0.0, INF) for c in mt.Kd.value], sigma=K.clamp(mt.sigma.value, 0, 90))   # matte.go:29-30
        elif kind == "Mirror":
            mat = dict(kind="mirror", kr=[K.clamp(c, 0.0, INF) for c in mt.Kr.value])   # mirror.go:28
        else:
            assert kind == "Glass" and mt.uRoughness.value == 0.0 and mt.vRoughness.value == 0.0
            mat = dict(kind="glass", R=[K.clamp(c, 0.0, 1.0) for c in mt.Kr.value], T=[K.clamp(c, 0.0, 1.0) for c in mt.Kt.value],
                       eta=mt.index.value)   # glass.go:32-37
        m, minv = sh.objectToWorld.Matrix.m, sh.objectToWorld.MatrixInverse.m
        phi_max = K.radians(K.clamp(float(sh.phiMax), 0.0, 360.0))   # disk.go:34
        assert phi_max == 2 * math.pi
        d = dict(m=m, minv=minv, height=float(sh.height), radius=float(sh.radius), inner=float(sh.innerRadius), phi_max=phi_max, mat=mat)
        d["bound"] = disk_world_bound(d)
        disks.append(d)
    lights = []
    for l in scene.lights:
        if type(l).__name__ == "Point":
            lights.append(dict(kind="point", p=list(map(float, l.pLight)), I=list(l.I)))
        else:
            sh = l.shape
            assert type(l).__name__ == "DiffuseAreaLight" and type(sh).__name__ == "Disk" and float(sh.phiMax) == 360.0
            lights.append(dict(kind="area", shape="disk", L=list(l.LEmit), two_sided=bool(l.twoSided), m=sh.objectToWorld.Matrix.m,
                               minv=sh.objectToWorld.MatrixInverse.m, height=float(sh.height), radius=float(sh.radius),
                               inner=float(sh.innerRadius), phi_max=K.radians(K.clamp(float(sh.phiMax), 0.0, 360.0))))
    cam = integ.GetCamera()
    film = cam.GetFilm()
    smp = integ.GetSampler()
    return dict(disks=disks, lights=lights, r2c=cam.RasterToCamera.Matrix.m, c2w=cam.cameraToWorld.startTransform.Matrix.m,
                lens_radius=cam.lensRadius, focal=cam.focalDistance, res=film.FullResolution, crop=film.CroppedPixelBounds,
                fr=film.Filter.radius, nx=smp.xSamples, ny=smp.ySamples, jitter=smp.jitter, ndims=smp.nDims,
                max_depth=integ.maxDepth, rr=integ.rrThreshold)


# ---------------------------------------------------------------- bounds.go:114-120,149-192,209-219; transform.go:335-344
def disk_world_bound(d):
    """Disk.ObjectBound (disk.go:40-53) through Transform.TransformBounds"""
    lo, hi = [-d["radius"], -d["radius"], d["height"]], [d["radius"], d["radius"], d["height"]]
    pick = lambda i: [(lo, hi)[i & 1][0], (lo, hi)[(i & 2) // 2][1], (lo, hi)[(i & 4) // 4][2]]
    c, _ = K.transform_point(d["m"], lo, Z3)
    bmin, bmax = list(c), list(c)
    for i in range(1, 8):
        c, _ = K.transform_point(d["m"], pick(i), Z3)
        bmin = [K.go_min(bmin[k], c[k]) for k in range(3)]
        bmax = [K.go_max(bmax[k], c[k]) for k in range(3)]
    return bmin, bmax


def go_div(a, b):  # Go float division: x/0 = +-Inf, 0/0 = NaN (Python raises)
    if b == 0.0:
        if a == 0.0 or a != a:
            return float("nan")
        return math.copysign(INF, a) * math.copysign(1.0, b)
    return a / b


def bounds_intersect_p(b, o, tmax, inv, neg):  # bounds.go:149-192
    g = 1 + 2 * K.gamma(3.0)
    t0 = (b[neg[0]][0] - o[0]) * inv[0]
    t1 = (b[1 - neg[0]][0] - o[0]) * inv[0]
    ty0 = (b[neg[1]][1] - o[1]) * inv[1]
    ty1 = (b[1 - neg[1]][1] - o[1]) * inv[1]
    t1 *= g
    ty1 *= g
    if t0 > ty1 or ty0 > t1:
        return False
    if ty0 > t0:
        t0 = ty0
    if ty1 < t1:
        t1 = ty1
    tz0 = (b[neg[2]][2] - o[2]) * inv[2]
    tz1 = (b[1 - neg[2]][2] - o[2]) * inv[2]
    tz1 *= g
    if t0 > tz1 or tz0 > t1:
        return False
    if tz0 > t0:
        t0 = tz0
    if tz1 < t1:
        t1 = tz1
    return t0 < tmax and t1 > 0


# ---------------------------------------------------------------- disk.go:64-158, interaction.go:171-206, transform.go:302-333
def disk_plane_hit(d, o, w, tmax):
    """the part Intersect and IntersectP share (disk.go:64-91 / 132-157): object-space ray, t, pHit — or None"""
    ro, rd = K.transform_ray(d["minv"], o, w)
    if rd[2] == 0:
        return None
    t = (d["height"] - ro[2]) / rd[2]
    if t <= 0 or t >= tmax:
        return None
    ph = [ro[0] + rd[0] * t, ro[1] + rd[1] * t, ro[2] + rd[2] * t]   # ray.go PointAt: Origin.Add(Direction.MulScalar(t))
    dist2 = ph[0] * ph[0] + ph[1] * ph[1]
    if dist2 > d["radius"] * d["radius"] or dist2 < d["inner"] * d["inner"]:
        return None
    phi = go_atan2(ph[1], ph[0])   # disk.go:83-89
    if phi < 0:
        phi += 2 * math.pi
    if phi > d["phi_max"]:
        return None
    return t, ph, dist2, rd, phi


def disk_intersect(d, o, w, tmax):
    r = disk_plane_hit(d, o, w, tmax)
    if r is None:
        return None
    t, ph, dist2, rd, phi = r
    r_hit = math.sqrt(dist2)
    uv = [phi / d["phi_max"], 1 - (r_hit - d["inner"]) / (d["radius"] - d["inner"])]   # disk.go:92-95
    dpdu = [-d["phi_max"] * ph[1], d["phi_max"] * ph[0], 0.0]
    k = (d["radius"] - d["inner"]) / r_hit
    dpdv = [ph[0] * k, ph[1] * k, 0.0 * k]
    ph[2] = d["height"]
    # NewSurfaceInteractionWith (interaction.go:171-206): one normal object serves Normal and Shading.Normal;
    # reverseOrientation == transformSwapsHandedness (both false for a Disk, disk.go:27-28): no flip
    n_obj = K.v_normalized(K.v_cross(dpdu, dpdv))
    wo_obj = K.v_muls(rd, -1.0)
    # TransformSurfaceInteraction (transform.go:302-333): `ret := *si` shares si.interaction and si.Shading; Normal is REPLACED
    # by the normalised transformed normal, Shading.Normal still points at the object-space one and is transformed WITHOUT
    # normalising, then face-forwarded; Shading.dpdu = the transformed object-space dpdu
    p, perr = K.transform_point(d["m"], ph, Z3)
    n = K.v_normalized(K.transform_normal(d["minv"], n_obj))
    wo = K.v_normalized(K.transform_vector(d["m"], wo_obj))
    ns = K.face_forward(K.transform_normal(d["minv"], n_obj), n)
    sh_dpdu = K.transform_vector(d["m"], dpdu)
    return t, dict(p=p, perr=perr, n=n, wo=wo, ns=ns, sh_dpdu=sh_dpdu, disk=d, uv=uv)


class Scene:
    """scene.go:38-46 over a BVH whose leaves hold one primitive (maxPrimsInNode 1): a primitive is reached iff its own
    world bound passes Bounds3.IntersectP with the ray's CURRENT tMax (bvh.go:659-712; GeometricPrimitive.Intersect
    shortens it, primitive.go:50-55).  The scene has no two surfaces a ray could meet at distances closer than the slab
    arithmetic's rounding, so the visiting order (here: primitive index) cannot change an answer."""

    def __init__(self, disks):
        self.disks = disks
        self.closest = 0
        self.shadow = 0

    @staticmethod
    def _inv(w):
        inv = [go_div(1.0, w[0]), go_div(1.0, w[1]), go_div(1.0, w[2])]
        return inv, [1 if c < 0 else 0 for c in inv]   # bvh.go:665-666, xyz.go:545-556

    def intersect(self, o, w, tmax):
        self.closest += 1
        inv, neg = self._inv(w)
        best = None
        for d in self.disks:
            if bounds_intersect_p(d["bound"], o, tmax, inv, neg):
                r = disk_intersect(d, o, w, tmax)
                if r is not None:
                    tmax, best = r
        return best

    def intersect_p(self, o, w, tmax):
        self.shadow += 1
        inv, neg = self._inv(w)
        for d in self.disks:
            if bounds_intersect_p(d["bound"], o, tmax, inv, neg) and disk_plane_hit(d, o, w, tmax) is not None:
                return True
        return False


# ---------------------------------------------------------------- reflection.go:128-298; matte.go, mirror.go, glass.go
REFL, TRANS, DIFF, GLOSSY, SPEC = K.BSDF_REFLECTION, K.BSDF_TRANSMISSION, K.BSDF_DIFFUSE, K.BSDF_GLOSSY, K.BSDF_SPECULAR
ALL = REFL | TRANS | DIFF | GLOSSY | SPEC
NON_SPECULAR = ALL & ~SPEC


def black(c):
    return all(v == 0.0 for v in c)


class BSDF:
    """a BSDF of at most ONE lobe (what the three materials build for smooth surfaces with allowMultipleLobes = true)"""

    def __init__(self, hit):  # NewBSDF (reflection.go:128-140)
        self.ns, self.ng = hit["ns"], hit["n"]
        self.ss = K.v_normalized(hit["sh_dpdu"])
        self.ts = K.v_cross(self.ns, self.ss)
        m = hit["disk"]["mat"]
        self.eta, self.lobe, self.type = 1.0, None, 0
        if m["kind"] == "matte":      # matte.go:27-37 -> LambertianReflection (reflection.go:576-581) or OrenNayar (:616-626)
            if not black(m["kd"]):
                self.lobe, self.type = ("lambert", m["kd"]) if m.get("sigma", 0.0) == 0 else ("orennayar", m["kd"], m["sigma"]), REFL | DIFF
        elif m["kind"] == "mirror":   # mirror.go:27-31 -> SpecularReflection + FresnelNoOp, typed Reflection|Diffuse (reflection.go:540)
            if not black(m["kr"]):
                self.lobe, self.type = ("specrefl", m["kr"]), REFL | DIFF
        else:                         # glass.go:27-48 -> FresnelSpecular(R, T, 1, eta, Radiance)
            self.eta = m["eta"]
            if not (black(m["R"]) and black(m["T"])):
                self.lobe, self.type = ("fresnel", m["R"], m["T"], m["eta"]), REFL | TRANS | SPEC

    def matches(self, flags):  # MatchesFlags (reflection.go:300-302)
        return self.lobe is not None and (self.type & flags) == self.type

    def num_components(self, flags):  # :159-167
        return 1 if self.matches(flags) else 0

    def to_local(self, v):  # :147-149
        return [K.v_dot(v, self.ss), K.v_dot(v, self.ts), K.v_dot(v, self.ns)]

    def _lobe_f(self, wo, wi):   # LambertianReflection.F :589-591, OrenNayar.F :628-652; the specular lobes answer zero (:553-555, :478-480)
        if self.lobe[0] == "lambert":
            return [c * K.INV_PI for c in self.lobe[1]]
        if self.lobe[0] == "orennayar":
            return K.oren_nayar(self.lobe[1], self.lobe[2], wo, wi)
        return list(Z3)

    def _lobe_pdf(self, wo, wi):   # :343-348 for the two diffuse lobes; the specular lobes answer zero (:572-574)
        if self.lobe[0] in ("lambert", "orennayar"):
            return abs(wi[2]) * K.INV_PI if wo[2] * wi[2] > 0 else 0.0
        return 0.0

    def f(self, wo_w, wi_w, flags):  # BSDF.F :170-187
        wi, wo = self.to_local(wi_w), self.to_local(wo_w)
        if wo[2] == 0.0:
            return list(Z3)
        reflect = K.v_dot(wi_w, self.ng) * K.v_dot(wo_w, self.ng) > 0
        f = list(Z3)
        if self.matches(flags) and ((reflect and self.type & REFL > 0) or (not reflect and self.type & TRANS > 0)):
            lf = self._lobe_f(wo, wi)
            f = [f[i] + lf[i] for i in range(3)]
        return f

    def pdf(self, wo_w, wi_w, flags):  # :255-277
        if self.lobe is None:
            return 0.0
        wo, wi = self.to_local(wo_w), self.to_local(wi_w)
        if wo[2] == 0:
            return 0.0
        if not self.matches(flags):
            return 0.0
        return (0.0 + self._lobe_pdf(wo, wi)) / 1.0

    def sample_f(self, wo_w, u, flags):  # :189-253 — returns the LOCAL wi (wiWorld is computed and dropped); one lobe: no pdf/f sums
        none = (list(Z3), list(Z3), 0.0, 0)
        if not self.matches(flags):
            return none
        if self.lobe[0] == "lambert":   # the KAT restatement holds the whole of BSDF.SampleF for this lobe
            f, wi, pdf = K.lambert_sample_f(self.lobe[1], self.to_local(wo_w), u, COS, SIN)
            return f, wi, pdf, 0
        comp = K.go_min(math.floor(u[0] * 1.0), 1.0 - 1)
        ur = [K.go_min(u[0] * 1.0 - comp, K.ONE_MINUS_EPSILON), u[1]]
        wo = self.to_local(wo_w)
        if wo[2] == 0.0:
            return none
        if self.lobe[0] == "orennayar":   # sampleF :305-314 with OrenNayar's F and the cosine pdf
            wi = K.cosine_sample_hemisphere(ur, COS, SIN)
            if wo[2] < 0:
                wi[2] *= -1
            f, pdf, st = self._lobe_f(wo, wi), self._lobe_pdf(wo, wi), 0
        elif self.lobe[0] == "specrefl":   # SpecularReflection.SampleF :557-562, FresnelNoOp :383-385; sampledType 0
            wi = [-wo[0], -wo[1], wo[2]]
            f, pdf, st = [(1.0 * c) / abs(wi[2]) for c in self.lobe[1]], 1.0, 0
        else:                            # FresnelSpecular.SampleF :482-523
            f, wi, pdf, st = K.fresnel_specular_sample_f(self.lobe[1], self.lobe[2], 1.0, self.lobe[3], wo, ur)
        if pdf == 0.0:
            return none
        return f, wi, pdf, st


# ---------------------------------------------------------------- pixel.go:55-76, sampler.go:21-35,71-77, stratified.go:21-48
class TileSampler:
    def __init__(self, sc, seed):
        self.sc = sc
        self.rng = K.Rng()
        self.rng.set_sequence(seed)   # pixel.go:41 (clone)
        self.t1, self.idx, self.d1, self.d2 = None, 0, 0, 0

    def start_pixel(self, px, py):
        self.t1 = K.stratified_start_pixel(self.rng, self.sc["nx"], self.sc["ny"], self.sc["jitter"], self.sc["ndims"])
        self.idx = 0   # sampler.go:21-27

    def start_next_sample(self):  # pixel.go:49-53 + sampler.go:29-35: the index moves BEFORE the first sample
        self.d1 = self.d2 = 0
        self.idx += 1
        return self.idx < self.sc["nx"] * self.sc["ny"]

    def get1d(self):
        if self.d1 < self.sc["ndims"]:
            v = self.t1[self.d1][self.idx]
            self.d1 += 1
            return v
        return self.rng.uniform()

    def get2d(self):
        if self.d2 < self.sc["ndims"]:
            self.d2 += 1
            return [0.0, 0.0]   # the 2-D tables are never filled (sampling.go:122-124 shuffles a copy)
        x = self.rng.uniform()
        y = self.rng.uniform()
        return [x, y]


# ---------------------------------------------------------------- go-pbrt_b200/go/gopbrt/fast_sampler.go (GOPBRT_MODE_FAST)
M32, M64 = (1 << 32) - 1, (1 << 64) - 1


def kensler_permute(i, l, p):  # fast_sampler.go:103-134, uint32 arithmetic
    w = l - 1
    for sh in (1, 2, 4, 8, 16):
        w |= w >> sh
    while True:
        i ^= p
        i = (i * 0xe170893d) & M32
        i ^= p >> 16
        i ^= (i & w) >> 4
        i ^= p >> 8
        i = (i * 0x0929eb3f) & M32
        i ^= p >> 23
        i ^= (i & w) >> 1
        i = (i * (1 | p >> 27)) & M32
        i = (i * 0x6935fa69) & M32
        i ^= (i & w) >> 11
        i = (i * 0x74dcb303) & M32
        i ^= (i & w) >> 2
        i = (i * 0x9e501cc3) & M32
        i ^= (i & w) >> 2
        i = (i * 0xc860a3df) & M32
        i &= w
        i ^= i >> 5
        if i < l:
            break
    return ((i + p) & M32) % l


def hash_u32(a, b):  # fast_sampler.go:136-145, uint64 arithmetic
    x = (a * 0x9E3779B97F4A7C15 + b * 0xD1B54A32D192ED03 + 0x632BE59BD9B4E019) & M64
    x ^= x >> 32
    x = (x * 0xD6E8FEB86659FD93) & M64
    x ^= x >> 32
    x = (x * 0xD6E8FEB86659FD93) & M64
    x ^= x >> 32
    return x & M32


class FastSampler:
    """FastStratified: the sampler that DEFINES the library's FAST mode — the reference's renderer, unchanged, with every
    (pixel, sample) an independent stream.  Restated from the Go source of the sampler (the specification a Go host compiles),
    not from the CUDA or oracle code."""

    def __init__(self, sc, seed):   # Clone ignores the tile seed (fast_sampler.go:53-58)
        self.sc = sc
        self.rng = K.Rng()
        self.spp = sc["nx"] * sc["ny"]
        self.pixel, self.idx, self.d1, self.d2 = 0, 0, 0, 0

    def start_pixel(self, px, py):
        x0, y0, x1, _ = self.sc["crop"]
        self.pixel = (py - y0) * (x1 - x0) + (px - x0)
        self.idx = 0

    def start_next_sample(self):
        self.d1 = self.d2 = 0
        self.idx += 1
        if self.idx < self.spp:
            self.rng.set_sequence((self.pixel * self.spp + self.idx) & M64)
        return self.idx < self.spp

    def get1d(self):
        if self.d1 < self.sc["ndims"]:
            j = kensler_permute(self.idx, self.spp, hash_u32(self.pixel, self.d1))
            self.d1 += 1
            delta = self.rng.uniform() if self.sc["jitter"] else 0.5
            return K.go_min((float(j) + delta) * (1.0 / float(self.spp)), K.ONE_MINUS_EPSILON)
        return self.rng.uniform()

    def get2d(self):
        if self.d2 < self.sc["ndims"]:
            self.d2 += 1
            return [0.0, 0.0]
        x = self.rng.uniform()
        y = self.rng.uniform()
        return [x, y]


SAMPLERS = {"stratified": TileSampler, "fast": FastSampler}


# ---------------------------------------------------------------- shape.go:29-48, disk.go:177-185, sampling.go:208-212
def shape_pdf_wi(light, hit, wi):
    """PdfWi: a ray from the reference point along wi (interaction.go:65-74) against the light's shape alone"""
    o = K.offset_ray_origin(hit["p"], hit["perr"], hit["n"], wi)
    r = disk_intersect(light, o, wi, INF)
    if r is None:
        return 0.0
    _, li = r
    area = light["phi_max"] * 0.5 * (light["radius"] * light["radius"] - light["inner"] * light["inner"])
    pdf = go_div(K.v_dist2(hit["p"], li["p"]), abs(K.v_dot(li["n"], K.v_muls(K.v_muls(wi, -1.0), area))))
    return 0.0 if math.isinf(pdf) else pdf


def power_heuristic(f_pdf, g_pdf):
    f, g = 1.0 * f_pdf, 1.0 * g_pdf
    return (f * f) / (f * f + g * g)


# ---------------------------------------------------------------- integrator.go:46-195
def estimate_direct(hit, bsdf, light, u_light, u_scattering, scene, stats):
    Li, wi, light_pdf, lp, lperr, ln, delta = K.light_sample_li(light, hit["p"], hit["perr"], hit["n"], u_light, COS, SIN)
    Ld = list(Z3)
    if light_pdf > 0 and not all(c == 0.0 for c in Li):
        f = bsdf.f(hit["wo"], wi, NON_SPECULAR)
        k = abs(K.v_dot(wi, hit["ns"]))
        f = [c * k for c in f]
        scattering_pdf = bsdf.pdf(hit["wo"], wi, NON_SPECULAR)
        if not all(c == 0.0 for c in f):
            o, w = K.spawn_ray_to(hit["p"], hit["perr"], hit["n"], lp, lperr, ln)   # origin = the un-offset point
            if scene.intersect_p(o, w, 1 - SHADOW_EPSILON):
                Li = list(Z3)
            if not all(c == 0.0 for c in Li):
                if delta:
                    Ld = [Ld[i] + (f[i] * Li[i]) / light_pdf for i in range(3)]
                else:
                    weight = power_heuristic(light_pdf, scattering_pdf)
                    Ld = [Ld[i] + ((f[i] * Li[i]) * weight) / light_pdf for i in range(3)]
    if not delta:
        stats["nondelta"] += 1   # what the library's `dead_mis_rays` counter reports: an upper bound of the rays counted below
        # :139-193: the BSDF-sampling leg.  f.MulScalar(|wi.ns|) is dropped (:147); wi is the LOCAL direction again; the hit
        # primitive's area light is always nil (primitive.go:29-36 never sets it), so the ray's answer adds nothing: the leg can
        # only end the estimate early (PdfLi == 0) — or cost one closest-hit query, counted here
        f, wi, scattering_pdf, sampled = bsdf.sample_f(hit["wo"], u_scattering, NON_SPECULAR)
        if not all(c == 0.0 for c in f) and scattering_pdf > 0.0:
            if sampled & SPEC == 0:
                light_pdf = shape_pdf_wi(light, hit, wi)
                if light_pdf == 0:
                    return Ld
                power_heuristic(scattering_pdf, light_pdf)
            stats["dead_mis"] += 1
    return Ld


def uniform_sample_one_light(hit, bsdf, sc, scene, smp, stats):
    n = len(sc["lights"])
    if n == 0:
        return list(Z3)
    if sc.get("light_strategy", "uniform") == "power":
        # ComputeLightPowerDistribution (lightdistribution.go:58-68) APPENDS the powers to a slice that already holds n zeros, and
        # every power is Spectrum.Y() == 0 (spectrum.go:227-229): 2n zeros -> funcInt == 0 -> SampleDiscrete's pdf is 0 (sampling.go:50-53)
        weights = [0.0] * n + [0.0] * n
    else:
        weights = [1.0] * n   # lightdistribution.go:24-34
    num, pdf = K.sample_discrete(weights, smp.get1d())   # sampling.go:42-55
    if pdf == 0.0:
        return list(Z3)
    u_light = smp.get2d()
    u_scattering = smp.get2d()
    Ld = estimate_direct(hit, bsdf, sc["lights"][num], u_light, u_scattering, scene, stats)
    # integrator.go:71: spectrum.DivScalar(lightPdf) returns a NEW spectrum that is dropped; :72-74 panics above 10
    stats["max_direct"] = max(stats["max_direct"], max(Ld))
    if K.go_max(K.go_max(Ld[0], Ld[1]), Ld[2]) > 10:
        stats["gt10"] += 1   # the reference panics here; the library counts the event (`radiance_gt10`) and goes on
    return Ld


# ---------------------------------------------------------------- path.go:32-157
def path_li(o, w, sc, scene, smp, stats):
    L, beta = list(Z3), [1.0, 1.0, 1.0]
    tmax = INF
    bounces = 0
    specular_bounce = False
    eta_scale = 1.0
    while True:
        bounces += 1
        hit = scene.intersect(o, w, tmax)
        if bounces == 0 or specular_bounce:   # bounces is at least 1 here (path.go:41)
            if hit is not None:
                # isect.Le: no primitive carries an area light (primitive.go:29-36) -> a zero spectrum is added
                L = [L[i] + beta[i] * 0.0 for i in range(3)]
            # else: no light has LightFlagInfinite (scene.go:22-27): nothing to loop over
        if hit is None or bounces >= sc["max_depth"]:
            break
        bsdf = BSDF(hit)
        if bsdf.num_components(NON_SPECULAR) > 0:
            Ld = uniform_sample_one_light(hit, bsdf, sc, scene, smp, stats)
            L = [L[i] + beta[i] * Ld[i] for i in range(3)]
        f, wi, pdf, flags = bsdf.sample_f(w, smp.get2d(), ALL)   # wo := ray.Direction (path.go:92)
        if all(c == 0.0 for c in f) or pdf == 0.0:
            break
        k = abs(K.v_dot(wi, hit["ns"])) / pdf
        beta = [beta[i] * (f[i] * k) for i in range(3)]
        specular_bounce = flags & SPEC != 0
        stats["bounce_kinds"][(bsdf.lobe[0], flags)] = stats["bounce_kinds"].get((bsdf.lobe[0], flags), 0) + 1
        if flags & SPEC > 0 and flags & TRANS > 0:
            eta = bsdf.eta
            if K.v_dot(w, hit["n"]) > 0:   # `wo.Dot(isect.Normal)` with wo = the ray's direction
                eta_scale *= eta * eta
            else:
                eta_scale *= 1 / (eta * eta)
        o = K.offset_ray_origin(hit["p"], hit["perr"], hit["n"], wi)   # isect.SpawnRay(wi), wi still LOCAL
        w, tmax = wi, INF
        rr = [c * eta_scale for c in beta]
        m = K.go_max(K.go_max(rr[0], rr[1]), rr[2])
        if m < sc["rr"] and bounces > 3:
            q = K.go_max(0.05, 1 - m)
            stats["rr_tests"] += 1
            if smp.get1d() < q:
                break
            beta = [c / (1 - q) for c in beta]
    return L


# ---------------------------------------------------------------- integrator.go:228-350, film.go:106-140,211-248
def render(sc, tile_size):
    x0, y0, x1, y1 = sc["crop"]
    W, H = x1 - x0, y1 - y0
    film = [[[0.0, 0.0, 0.0, 0.0] for _ in range(W)] for _ in range(H)]
    scene = Scene(sc["disks"])
    stats = dict(max_direct=0.0, camera=0, dead_mis=0, nondelta=0, rr_tests=0, gt10=0, bounce_kinds={})
    ntx, nty = (W + tile_size - 1) // tile_size, (H + tile_size - 1) // tile_size
    # multi-GPU partition of the library (not of the reference): rank r of `world` renders samples s % world == r (FAST) or tiles
    # t % world == r (STRICT) into a full-resolution film of its own; the films are then summed
    fast, rank, world = sc.get("sampler") == "fast", sc.get("rank", 0), sc.get("world", 1)
    for ty in range(nty):
        for tx in range(ntx):
            if not fast and world > 1 and (ty * ntx + tx) % world != rank:
                continue   # the library's STRICT partition: whole tiles, so that a pixel's sample sequence stays the reference's
            smp = SAMPLERS[sc.get("sampler", "stratified")](sc, ty * ntx + tx)
            bx0, by0 = x0 + tx * tile_size, y0 + ty * tile_size
            bx1, by1 = int(K.go_min(float(bx0 + tile_size), float(x1))), int(K.go_min(float(by0 + tile_size), float(y1)))
            tile = {}
            pb = None
            for py in range(by0, by1):
                for px in range(bx0, bx1):
                    smp.start_pixel(px, py)
                    while smp.start_next_sample():
                        if fast and world > 1 and smp.idx % world != rank:
                            continue   # the library's FAST partition: a stream per (pixel, sample), nothing to skip over
                        off = smp.get2d()
                        p_film = [float(px) + off[0], float(py) + off[1]]   # sampler.go:71-77
                        p_lens = smp.get2d()
                        smp.get1d()   # time: shutterClose == shutterOpen == 0 and nothing moves
                        o, w = K.camera_ray(sc["r2c"], sc["c2w"], sc["lens_radius"], sc["focal"], p_film, p_lens, COS, SIN)
                        stats["camera"] += 1
                        L = path_li(o, w, sc, scene, smp, stats)
                        if any(c != c for c in L):
                            L = [0.1, 0.1, 0.1]   # integrator.go:251-252; Spectrum.Y() is 0 (spectrum.go:227-229): no other fix-up
                        pb, add = K.film_tile_add_sample((bx0, by0, bx1, by1), sc["fr"][0], sc["fr"][1], sc["crop"], p_film, L)
                        for key, (r, g, b, fw) in add.items():
                            acc = tile.setdefault(key, [0.0, 0.0, 0.0, 0.0])
                            acc[0] += r
                            acc[1] += g
                            acc[2] += b
                            acc[3] += fw
            for (x, y), acc in tile.items():   # MergeFilmTile (film.go:122-140)
                assert pb[0] <= x < pb[2] and pb[1] <= y < pb[3]
                xyz = K.rgb_to_xyz(acc[:3])
                px = film[y - y0][x - x0]
                for i in range(3):
                    px[i] += xyz[i]
                px[3] += acc[3]
    stats["closest"], stats["shadow"] = scene.closest, scene.shadow
    return film, stats


def main():
    gp = importlib.import_module("go-pbrt_b200")
    scene, integ = scene_and_integrator(gp)
    sc = plain_scene(scene, integ)
    out = dict(note="made by tests/golden/make_path_golden.py (plain-Python restatement of the composed Path.Li hot path); "
                    "film = [y][x][X, Y, Z, filterWeightSum] as float.hex()", cases={})
    for tile in TILE_SIZES:
        film, st = render(sc, tile)
        assert st["max_direct"] <= 10.0, "UniformSampleOneLight would panic in the reference (integrator.go:72-74)"
        lit = sum(1 for row in film for p in row if p[1] > 0)
        print(f"tile {tile}: camera {st['camera']}, closest {st['closest']}, shadow {st['shadow']}, dead MIS rays {st['dead_mis']} of {st['nondelta']} area-light estimates, lit pixels {lit}/{len(film) * len(film[0])}, "
              f"max direct {st['max_direct']:.3f}")
        out["cases"][f"tile{tile}"] = dict(tile=tile, rays=[st["camera"], st["closest"], st["shadow"]], dead_mis_rays=st["dead_mis"], nondelta_estimates=st["nondelta"],
                                           coverage=dict(russian_roulette_tests=st["rr_tests"],
                                                         bounces={f"{k[0]}:{k[1]}": v for k, v in sorted(st["bounce_kinds"].items())}),
                                           film=[[[v.hex() for v in p] for p in row] for row in film])
    with open(os.path.join(HERE, "path_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote path_golden.json")


if __name__ == "__main__":
    main()
