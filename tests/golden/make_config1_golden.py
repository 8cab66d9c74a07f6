"""BASELINE config 1 — the reference's OWN scene (internal/render/server.go:29-164: 21 matte spheres behind
TransformedPrimitives, two 10 km checkerboard disks, a distant light, two point lights, a sphere area light, the Path
integrator) — rendered at 24x14 pixels by an independent plain-Python restatement of the hot path.

Same rules as make_path_golden.py (whose render loop, Path.Li, EstimateDirect, samplers, disks, BSDFs and film this file
reuses): written from the Go source, nothing from oracle/ or go-pbrt_b200/csrc/.  What is new here:
 * Sphere.Intersect / IntersectP (sphere.go:64-262) over the EFloat interval arithmetic (pkg/efloat/efloat.go, math.go —
   with MachineEpsilon = the smallest denormal, pkg/math/math.go:17) and its Quadratic;
 * TransformedPrimitive (primitive.go:94-115): the ray goes through the inverse, the hit through a SECOND
   TransformSurfaceInteraction whose result is assigned to a local variable — only what lives behind the shared *interaction and
   *Shading pointers reaches the caller (transform.go:302-333);
 * Sphere.PdfWi (sphere.go:350-365), the sphere area light's SampleLi (through make_shading_kats.sphere_sample_at), the distant
   light with its pOutside = wLight * 2R (distant.go:36-44) and the world radius of scene.go:17-21 / bounds.go:105-112;
 * Checkerboard2D over PlanarMapping2D (checkerboard.go:30-40, texture.go:41-46);
 * Go's math.Acos / Atan2 (asin.go, atan.go, atan2.go of Go 1.11: Cephes' atan polynomial), restated in make_path_golden.py;
   math.Sin from the host mirror go-pbrt_b200/gomath.py.

    python tests/golden/make_config1_golden.py        # rewrites tests/golden/config1_golden.json
"""
import contextlib
import importlib
import importlib.util
import json
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
_spec = importlib.util.spec_from_file_location("make_path_golden", os.path.join(HERE, "make_path_golden.py"))
M = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(M)
K = M.K
Z3, INF = M.Z3, M.INF
W, H, SPP, TILE = 24, 14, (3, 3), 4
RETRIES = dict(tried=0, hit=0)   # coverage counter: the second-root retry of a clipped sphere (sphere.go:110-131)
_inv_dir = M.Scene._inv   # bvh.go:665-666 (M.Scene itself is swapped out while this file renders)


def scene_and_integrator(gp):
    return gp.scenes.config1(W=W, H=H, spp=SPP)


go_acos = M.go_acos   # Go 1.11's math.Acos, restated in make_path_golden.py


# ---------------------------------------------------------------- pkg/efloat
class EFloatPanic(Exception):
    pass


class EF:
    __slots__ = ("v", "lo", "hi")

    def __init__(self, v, err=0.0):  # New (efloat.go:9-22)
        self.v = self.lo = self.hi = v
        if err != 0:
            self.lo = K.next_float_down(v - err)
            self.hi = K.next_float_up(v + err)
        self.check()

    @staticmethod
    def raw(v, lo, hi):
        r = EF.__new__(EF)
        r.v, r.lo, r.hi = v, lo, hi
        r.check()
        return r

    def check(self):  # efloat.go:103-111: the reference panics
        for b in (self.lo, self.hi):
            if math.isinf(b) or b != b:
                raise EFloatPanic()
        if self.lo > self.hi:
            raise EFloatPanic()

    def add(self, o):
        return EF.raw(self.v + o.v, K.next_float_down(self.lo + o.lo), K.next_float_up(self.hi + o.hi))

    def sub(self, o):
        return EF.raw(self.v - o.v, K.next_float_down(self.lo - o.hi), K.next_float_up(self.hi - o.lo))

    def mul(self, o):
        p = [self.lo * o.lo, self.hi * o.lo, self.lo * o.hi, self.hi * o.hi]
        return EF.raw(self.v * o.v, K.next_float_down(K.go_min(K.go_min(p[0], p[1]), K.go_min(p[2], p[3]))),
                      K.next_float_up(K.go_max(K.go_max(p[0], p[1]), K.go_max(p[2], p[3]))))

    def mul_scalar(self, s):
        return self.mul(EF(s, 0.0))

    def div(self, o):
        v = M.go_div(self.v, o.v)
        if o.lo < 0 and o.hi > 0:
            return EF.raw(v, -INF, INF)   # panics in check(), as the reference does
        d = [M.go_div(self.lo, o.lo), M.go_div(self.hi, o.lo), M.go_div(self.lo, o.hi), M.go_div(self.hi, o.hi)]
        return EF.raw(v, K.next_float_down(K.go_min(K.go_min(d[0], d[1]), K.go_min(d[2], d[3]))),
                      K.next_float_up(K.go_max(K.go_max(d[0], d[1]), K.go_max(d[2], d[3]))))


def quadratic(a, b, c):  # pkg/efloat/math.go:34-57
    disc = b.v * b.v - 4. * a.v * c.v
    if disc < 0:
        return None
    root = math.sqrt(disc)
    froot = EF(root, K.MACHINE_EPSILON * root)
    q = (b.sub(froot) if b.v < 0 else b.add(froot)).mul_scalar(-0.5)
    t0 = q.div(a)
    t1 = c.div(q)
    if t0.v > t1.v:
        t0, t1 = t1, t0
    return t0, t1


# ---------------------------------------------------------------- transform.go:257-300,302-344
def transform_ray_err(m, o, w):
    o2, oerr = K.transform_point(m, o, Z3)
    d2 = K.transform_vector(m, w)
    g3 = K.gamma(3.0)
    derr = [g3 * (abs(m[r][0] * w[0]) + abs(m[r][1] * w[1]) + abs(m[r][2] * w[2])) for r in range(3)]
    l2 = K.v_len2(d2)
    if l2 > 0:
        dt = K.v_dot(K.v_abs(d2), oerr) / l2
        o2 = K.v_add(o2, K.v_muls(d2, dt))
    return o2, d2, oerr, derr


def tsi(m, minv, r):
    """TransformSurfaceInteraction as the caller sees it: Point/PointError/Normal/Wo (behind *interaction) and Shading.* (behind
    *Shading); Shading.Normal is transformed from the Shading's own object, NOT normalised, then face-forwarded"""
    p, perr = K.transform_point(m, r["p"], r["perr"])
    n = K.v_normalized(K.transform_normal(minv, r["n"]))
    wo = K.v_normalized(K.transform_vector(m, r["wo"]))
    ns = K.face_forward(K.transform_normal(minv, r["ns"]), n)
    return dict(r, p=p, perr=perr, n=n, wo=wo, ns=ns, sh_dpdu=K.transform_vector(m, r["sh_dpdu"]))


def transform_bounds(m, lo, hi):
    pick = lambda i: [(lo, hi)[i & 1][0], (lo, hi)[(i & 2) // 2][1], (lo, hi)[(i & 4) // 4][2]]
    c, _ = K.transform_point(m, lo, Z3)
    bmin, bmax = list(c), list(c)
    for i in range(1, 8):
        c, _ = K.transform_point(m, pick(i), Z3)
        bmin = [K.go_min(bmin[k], c[k]) for k in range(3)]
        bmax = [K.go_max(bmax[k], c[k]) for k in range(3)]
    return bmin, bmax


# ---------------------------------------------------------------- sphere.go:64-262
def sphere_roots(sp, o, w, tmax):
    """what Intersect and IntersectP share: object-space ray, the chosen root, pHit"""
    ro, rd, oerr, derr = transform_ray_err(sp["minv"], o, w)
    ox, oy, oz = EF(ro[0], oerr[0]), EF(ro[1], oerr[1]), EF(ro[2], oerr[2])
    dx, dy, dz = EF(rd[0], derr[0]), EF(rd[1], derr[1]), EF(rd[2], derr[2])
    a = dx.mul(dx).add(dy.mul(dy)).add(dz.mul(dz))
    b = dx.mul(ox).add(dy.mul(oy)).add(dz.mul(oz)).mul_scalar(2.0)
    c = ox.mul(ox).add(oy.mul(oy)).add(oz.mul(oz)).sub(EF(sp["radius"], 0).mul_scalar(sp["radius"]))
    q = quadratic(a, b, c)
    if q is None:
        return None
    t0, t1 = q
    if t0.hi > tmax or t1.lo <= 0:
        return None
    t = t0
    if t.lo <= 0:
        t = t1
        if t.hi > tmax:
            return None
    r = sp["radius"]

    def point_at(tv):   # :96-106
        ph = [ro[i] + rd[i] * tv for i in range(3)]
        ph = K.v_muls(ph, r / math.sqrt(K.v_dist2(ph, Z3)))
        if ph[0] == 0.0 and ph[1] == 0.0:
            ph[0] = 1e-5 * r
        phi = M.go_atan2(ph[1], ph[0])
        if phi < 0.0:
            phi += 2 * math.pi
        return ph, phi

    def clipped(ph, phi):   # :109
        return (sp["z_min"] > -r and ph[2] < sp["z_min"]) or (sp["z_max"] < r and ph[2] > sp["z_max"]) or phi > sp["phi_max"]

    ph, phi = point_at(t.v)
    if clipped(ph, phi):
        if t is t1:
            return None
        if t1.hi > tmax:
            return None
        RETRIES["tried"] += 1
        t = t1
        ph, phi2 = point_at(t.v)   # `phi := ...` (:125) declares a NEW variable: u below still comes from the first root's phi
        if clipped(ph, phi2):
            return None
        RETRIES["hit"] += 1
    return t.v, ph, rd, phi


def sphere_intersect(sp, o, w, tmax):
    q = sphere_roots(sp, o, w, tmax)
    if q is None:
        return None
    t, ph, rd, phi = q
    r, phi_max = sp["radius"], sp["phi_max"]
    theta = go_acos(K.clamp(ph[2] / r, -1, 1))
    uv = [phi / phi_max, (theta - sp["theta_min"]) / (sp["theta_max"] - sp["theta_min"])]   # :136-138
    z_radius = math.sqrt(ph[0] * ph[0] + ph[1] * ph[1])
    inv = 1.0 / z_radius
    cos_phi, sin_phi = ph[0] * inv, ph[1] * inv
    dpdu = [-phi_max * ph[1], phi_max * ph[0], 0.0]
    dt = sp["theta_max"] - sp["theta_min"]
    dpdv = K.v_muls([ph[2] * cos_phi, ph[2] * sin_phi, -r * M.SIN(theta)], dt)
    perr = K.v_muls(K.v_abs(ph), K.gamma(5.0))
    n = K.v_normalized(K.v_cross(dpdu, dpdv))   # NewSurfaceInteractionWith (interaction.go:171-177)
    if sp["reverse"]:                            # reverseOrientation != transformSwapsHandedness (never set: false)
        n = K.v_muls(n, -1.0)
    rec = dict(p=ph, perr=perr, n=n, ns=n, wo=K.v_muls(rd, -1.0), sh_dpdu=dpdu, uv=uv)
    return t, tsi(sp["m"], sp["minv"], rec)


def sphere_pdf_wi(light, hit, wi):  # sphere.go:350-365
    pc, _ = K.transform_point(light["m"], Z3, Z3)
    po = K.offset_ray_origin(hit["p"], hit["perr"], hit["n"], K.v_sub(hit["p"], pc))
    r2 = light["radius"] * light["radius"]
    assert K.v_dist2(po, pc) > r2, "a reference point inside the light sphere: the generic PdfWi path is not restated here"
    sin2 = r2 / K.v_dist2(hit["p"], pc)
    cos_max = math.sqrt(K.go_max(0, 1.0 - sin2))
    return 1.0 / (2.0 * math.pi * (1.0 - cos_max))   # UniformConePdf (sampling.go:169-171)


# ---------------------------------------------------------------- triangles (used by make_config2_golden.py only)
def tri_intersect(tr, o, w, tmax, want_hit=True):
    """NOT a reference function: the reference has no triangle.  This restates the library's own DEFINITION of the shape
    (include/gopbrt_cuda.h:66-71): world-space float64 vertices, pbrt-v3's watertight test (translate, permute,
    shear, edge functions) in float64, `t <= 0 || t >= tMax` rejected as disk.go:75 does, hit point = barycentric combination,
    pError = Gamma(7) * sum |b_i p_i|, uv (0,0),(1,0),(1,1), n = normalize(cross(p0 - p2, p1 - p2)) flipped by reverseOrientation,
    shading normal = n, wo = -d un-normalised (no TransformSurfaceInteraction: the vertices are in world space)."""
    p0, p1, p2 = tr["p"]
    p0t, p1t, p2t = K.v_sub(p0, o), K.v_sub(p1, o), K.v_sub(p2, o)
    a = K.v_abs(w)
    kz = (0 if a[0] > a[2] else 2) if a[0] > a[1] else (1 if a[1] > a[2] else 2)
    kx = (kz + 1) % 3
    ky = (kx + 1) % 3
    d = [w[kx], w[ky], w[kz]]
    p0t, p1t, p2t = ([q[kx], q[ky], q[kz]] for q in (p0t, p1t, p2t))
    sx, sy, sz = M.go_div(-d[0], d[2]), M.go_div(-d[1], d[2]), M.go_div(1.0, d[2])
    for q in (p0t, p1t, p2t):
        q[0] += sx * q[2]
        q[1] += sy * q[2]
    e0 = p1t[0] * p2t[1] - p1t[1] * p2t[0]
    e1 = p2t[0] * p0t[1] - p2t[1] * p0t[0]
    e2 = p0t[0] * p1t[1] - p0t[1] * p1t[0]
    if (e0 < 0 or e1 < 0 or e2 < 0) and (e0 > 0 or e1 > 0 or e2 > 0):
        return None
    det = e0 + e1 + e2
    if det == 0:
        return None
    for q in (p0t, p1t, p2t):
        q[2] *= sz
    t_scaled = e0 * p0t[2] + e1 * p1t[2] + e2 * p2t[2]
    if det < 0 and (t_scaled >= 0 or t_scaled < tmax * det):
        return None
    if det > 0 and (t_scaled <= 0 or t_scaled > tmax * det):
        return None
    inv_det = 1 / det
    b0, b1, b2 = e0 * inv_det, e1 * inv_det, e2 * inv_det
    t = t_scaled * inv_det
    if t <= 0 or t >= tmax:
        return None
    if not want_hit:
        return t, None
    dp02, dp12 = K.v_sub(p0, p2), K.v_sub(p1, p2)
    du02, dv02, du12, dv12 = -1.0, -1.0, 0.0, -1.0
    inv = 1 / (du02 * dv12 - dv02 * du12)
    dpdu = K.v_muls(K.v_sub(K.v_muls(dp02, dv12), K.v_muls(dp12, dv02)), inv)
    dpdv = K.v_muls(K.v_add(K.v_muls(dp02, -du12), K.v_muls(dp12, du02)), inv)
    n = K.v_normalized(K.v_cross(dp02, dp12))
    assert K.v_len2(K.v_cross(dpdu, dpdv)) != 0, "degenerate triangle: the CoordinateSystem fallback is not restated here"
    p_abs = K.v_add(K.v_add(K.v_abs(K.v_muls(p0, b0)), K.v_abs(K.v_muls(p1, b1))), K.v_abs(K.v_muls(p2, b2)))
    ph = K.v_add(K.v_add(K.v_muls(p0, b0), K.v_muls(p1, b1)), K.v_muls(p2, b2))
    if tr["reverse"]:
        n = K.v_muls(n, -1.0)
    return t, dict(p=ph, perr=K.v_muls(p_abs, K.gamma(7.0)), n=n, ns=n, wo=K.v_muls(w, -1.0), sh_dpdu=dpdu,
                   uv=[b0 * 0 + b1 * 1 + b2 * 1, b0 * 0 + b1 * 0 + b2 * 1])


# ---------------------------------------------------------------- primitives: GeometricPrimitive / TransformedPrimitive
class Prim:
    def __init__(self, kind, shape, mat, xf=None):
        self.kind, self.shape, self.mat, self.xf = kind, shape, mat, xf
        if kind == "tri":
            b = ([min(q[k] for q in shape["p"]) for k in range(3)], [max(q[k] for q in shape["p"]) for k in range(3)])
        elif kind == "disk":
            b = transform_bounds(shape["m"], [-shape["radius"], -shape["radius"], shape["height"]], [shape["radius"], shape["radius"], shape["height"]])
        else:
            r = shape["radius"]   # Sphere.ObjectBound (sphere.go:46-51)
            b = transform_bounds(shape["m"], [-r, -r, shape["z_min"]], [r, r, shape["z_max"]])
        if xf is not None:   # primitive.go:128-130 + transform.go:583-586
            b = transform_bounds(xf[0], b[0], b[1])
        self.bound = b

    def _local_ray(self, o, w):
        if self.xf is None:
            return o, w
        return K.transform_ray(self.xf[1], o, w)   # interpolatedPrimToWorld.Inverse().TransformRay (primitive.go:96)

    def intersect(self, o, w, tmax):
        lo, lw = self._local_ray(o, w)
        if self.kind == "tri":
            r = tri_intersect(self.shape, lo, lw, tmax)
        else:
            r = M.disk_intersect(self.shape, lo, lw, tmax) if self.kind == "disk" else sphere_intersect(self.shape, lo, lw, tmax)
        if r is None:
            return None
        t, rec = r
        if self.xf is not None and not self.xf[2]:   # IsIdentity (primitive.go:104-106)
            rec = tsi(self.xf[0], self.xf[1], rec)
        rec["disk"] = dict(mat=self.material_at(rec))
        return t, rec

    def intersect_p(self, o, w, tmax):
        lo, lw = self._local_ray(o, w)
        if self.kind == "tri":
            return tri_intersect(self.shape, lo, lw, tmax, want_hit=False) is not None
        if self.kind == "disk":
            return M.disk_plane_hit(self.shape, lo, lw, tmax) is not None
        return sphere_roots(self.shape, lo, lw, tmax) is not None

    def material_at(self, rec):
        m = self.mat
        if m["kind"] != "checker":
            return m
        if "uv" in m:   # UVMapping2D.Map (texture.go:22-26)
            su, sv, du, dv = m["uv"]
            s, t = su * rec["uv"][0] + du, sv * rec["uv"][1] + dv
        else:           # PlanarMapping2D.Map (texture.go:41-46); Checkerboard2D.Evaluate (checkerboard.go:30-40) below
            p = rec["p"]
            s = m["ds"] + K.v_dot(p, m["vs"])
            t = m["dt"] + K.v_dot(p, m["vt"])
        even = int(math.floor(s) + math.floor(t)) % 2 == 0
        return dict(kind="matte", kd=[K.clamp(c, 0.0, INF) for c in (m["tex1"] if even else m["tex2"])], sigma=m["sigma"])


class Scene:
    """as make_path_golden.Scene, over Prims: a primitive is tested iff its own world bound passes Bounds3.IntersectP with the
    ray's current tMax"""

    ORDER_INDEPENDENT = False   # make_config2_golden.py: the library's closest-hit definition instead of the running tMax

    def __init__(self, prims):
        self.prims = prims
        self.closest = self.shadow = 0

    def intersect(self, o, w, tmax):
        self.closest += 1
        inv, neg = _inv_dir(w)
        best = None
        if Scene.ORDER_INDEPENDENT:
            # every primitive whose own bound and shape test pass with the ray's ORIGINAL tMax; the least t wins, equal t goes to the
            # lower primitive index
            best_t = None
            for pr in self.prims:
                if M.bounds_intersect_p(pr.bound, o, tmax, inv, neg):
                    r = pr.intersect(o, w, tmax)
                    if r is not None and (best_t is None or r[0] < best_t):
                        best_t, best = r
            return best
        for pr in self.prims:
            if M.bounds_intersect_p(pr.bound, o, tmax, inv, neg):
                r = pr.intersect(o, w, tmax)
                if r is not None:
                    tmax, best = r
        return best

    def intersect_p(self, o, w, tmax):
        self.shadow += 1
        inv, neg = _inv_dir(w)
        for pr in self.prims:
            if M.bounds_intersect_p(pr.bound, o, tmax, inv, neg) and pr.intersect_p(o, w, tmax):
                return True
        return False


def plain_scene(scene, integ):
    def xf_of(t):
        return t.Matrix.m, t.MatrixInverse.m

    def shape_of(sh):
        m, minv = xf_of(sh.objectToWorld)
        if type(sh).__name__ == "Disk":
            phi_max = K.radians(K.clamp(float(sh.phiMax), 0.0, 360.0))   # disk.go:34
            return "disk", dict(m=m, minv=minv, height=float(sh.height), radius=float(sh.radius), inner=float(sh.innerRadius), phi_max=phi_max)
        assert type(sh).__name__ == "Sphere"
        r, z0, z1 = float(sh.radius), float(sh.zMin), float(sh.zMax)   # NewSphere (sphere.go:19-32)
        return "sphere", dict(m=m, minv=minv, radius=r, reverse=bool(sh.reverseOrientation), phi_max=K.radians(K.clamp(float(sh.phiMax), 0.0, 360.0)),
                              z_min=K.clamp(K.go_min(z0, z1), -r, r), z_max=K.clamp(K.go_max(z0, z1), -r, r),
                              theta_min=go_acos(K.clamp(K.go_min(z0, z1) / r, -1, 1)), theta_max=go_acos(K.clamp(K.go_max(z0, z1) / r, -1, 1)))

    def mat_of(mt):
        if type(mt).__name__ == "Mirror":
            return dict(kind="mirror", kr=[K.clamp(c, 0.0, INF) for c in mt.Kr.value])   # mirror.go:28
        if type(mt).__name__ == "Glass":
            assert mt.uRoughness.value == 0.0 and mt.vRoughness.value == 0.0
            return dict(kind="glass", R=[K.clamp(c, 0.0, 1.0) for c in mt.Kr.value], T=[K.clamp(c, 0.0, 1.0) for c in mt.Kt.value],
                        eta=mt.index.value)   # glass.go:32-37
        assert type(mt).__name__ == "MatteMaterial"
        sigma = K.clamp(mt.sigma.value, 0, 90)   # matte.go:30
        kd = mt.Kd
        if type(kd).__name__ == "Checkerboard2D":
            mp = kd.mapping
            if type(mp).__name__ == "UVMapping2D":
                return dict(kind="checker", uv=(float(mp.su), float(mp.sv), float(mp.du), float(mp.dv)),
                            tex1=list(kd.tex1.value), tex2=list(kd.tex2.value), sigma=sigma)
            assert type(mp).__name__ == "PlanarMapping2D"
            return dict(kind="checker", vs=list(map(float, mp.vs)), vt=list(map(float, mp.vt)), ds=float(mp.ds), dt=float(mp.dt),
                        tex1=list(kd.tex1.value), tex2=list(kd.tex2.value), sigma=sigma)
        return dict(kind="matte", kd=[K.clamp(c, 0.0, INF) for c in kd.value], sigma=sigma)

    prims = []
    for pr in scene.aggregate.primitives:
        xf = None
        if type(pr).__name__ == "TriangleMesh":   # M GeometricPrimitives in place, in index order
            for tri in pr.indices:
                prims.append(Prim("tri", dict(p=[[float(c) for c in pr.vertices[int(i)]] for i in tri], reverse=bool(pr.reverseOrientation)),
                                  mat_of(pr.material)))
            continue
        if type(pr).__name__ == "TransformedPrimitive":
            t = pr.primitiveToWorld.startTransform
            xf = xf_of(t) + (t.IsIdentity(),)
            pr = pr.primitive
        kind, shape = shape_of(pr.Shape)
        prims.append(Prim(kind, shape, mat_of(pr.material), xf))
    # scene.go:17-21: the aggregate's bound; BoundingSphere (bounds.go:105-112)
    lo = [min(p.bound[0][k] for p in prims) for k in range(3)]
    hi = [max(p.bound[1][k] for p in prims) for k in range(3)]
    centre = [(lo[k] + hi[k]) / 2.0 for k in range(3)]
    inside = all(lo[k] <= centre[k] <= hi[k] for k in range(3))
    world_radius = math.sqrt(K.v_dist2(centre, hi)) if inside else 0.0
    lights = []
    for l in scene.lights:
        kind = type(l).__name__
        if kind == "Point":
            lights.append(dict(kind="point", p=list(map(float, l.pLight)), I=list(l.I)))
        elif kind == "Distant":
            lights.append(dict(kind="distant", L=list(l.L), w=list(map(float, l.wLight)), world_radius=world_radius))
        else:
            k2, shape = shape_of(l.shape)
            assert kind == "DiffuseAreaLight" and not shape.get("reverse", False)
            lights.append(dict(shape, kind="area", shape=k2, L=list(l.LEmit), two_sided=bool(l.twoSided)))
    sc = {}   # the camera / film / sampler fields as make_path_golden.plain_scene fills them
    cam = integ.GetCamera()
    film = cam.GetFilm()
    smp = integ.GetSampler()
    sc.update(prims=prims, lights=lights, r2c=cam.RasterToCamera.Matrix.m, c2w=cam.cameraToWorld.startTransform.Matrix.m,
              lens_radius=cam.lensRadius, focal=cam.focalDistance, res=film.FullResolution, crop=film.CroppedPixelBounds,
              fr=film.Filter.radius, nx=smp.xSamples, ny=smp.ySamples, jitter=smp.jitter, ndims=smp.nDims,
              max_depth=integ.maxDepth, rr=integ.rrThreshold, disks=None, world_radius=world_radius)
    return sc


@contextlib.contextmanager
def patched(sc):
    """make_path_golden with this file's Scene and a PdfWi that knows spheres"""
    saved = M.Scene, M.shape_pdf_wi

    class _Scene(Scene):
        def __init__(self, _disks):
            Scene.__init__(self, sc["prims"])

    M.Scene = _Scene
    M.shape_pdf_wi = lambda light, hit, wi: sphere_pdf_wi(light, hit, wi) if light["shape"] == "sphere" else saved[1](light, hit, wi)
    try:
        yield
    finally:
        M.Scene, M.shape_pdf_wi = saved


def render(sc, tile):
    with patched(sc):
        return M.render(sc, tile)


def main():
    gp = importlib.import_module("go-pbrt_b200")
    scene, integ = scene_and_integrator(gp)
    sc = plain_scene(scene, integ)
    film, st = render(sc, TILE)
    lit = sum(1 for row in film for p in row if p[1] > 0)
    print(f"config 1 at {W}x{H}, tile {TILE}: camera {st['camera']}, closest {st['closest']}, shadow {st['shadow']}, area-light estimates {st['nondelta']}, "
          f"lit pixels {lit}/{W * H}, max direct {st['max_direct']:.3f}, bounces {st['bounce_kinds']}, roulette tests {st['rr_tests']}")
    out = dict(note="made by tests/golden/make_config1_golden.py (plain-Python restatement of the hot path on BASELINE config 1); "
                    "film = [y][x][X, Y, Z, filterWeightSum] as float.hex()",
               width=W, height=H, spp=list(SPP), tile=TILE, rays=[st["camera"], st["closest"], st["shadow"]], nondelta_estimates=st["nondelta"],
               radiance_gt10=st.get("gt10", 0), world_radius=sc["world_radius"].hex(),
               film=[[[v.hex() for v in p] for p in row] for row in film])
    with open(os.path.join(HERE, "config1_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote config1_golden.json")


if __name__ == "__main__":
    main()
