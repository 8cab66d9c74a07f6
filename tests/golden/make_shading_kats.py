"""Hand-derived known-answer vectors for the SHADING rows of the hot path.

The reference holds no golden value for these functions and cannot be run here (no Go toolchain), and the oracle
(oracle/*.h) and the CUDA code were written by the same builder from the same survey.  This script is the third,
INDEPENDENT restatement: plain Python floats (IEEE double, one rounding per operation, no FMA), one function per cited
Go line range, written from the Go source and from nothing in oracle/ or go-pbrt_b200/csrc/.  It emits
tests/golden/shading_kats.json; tests/test_shading_kats.py checks the oracle (CPU) and the CUDA device functions (GPU,
through gopbrt_kat_eval) against it bit for bit.

Trig: Go's math.Sin/Cos are its own Cephes-derived routines (not libm).  Cases marked trig="exact" only meet arguments
whose sine/cosine is exact in every implementation (0); cases marked trig="gomath" take sin/cos from the host mirror
go-pbrt_b200/gomath.py (pinned by the reference's transform_test.go:77-81) — every other operation is still independent.

    python tests/golden/make_shading_kats.py        # rewrites tests/golden/shading_kats.json
"""
import importlib
import json
import math
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
gomath = importlib.import_module("go-pbrt_b200.gomath")

INF = float("inf")


# ---------------------------------------------------------------- pkg/math/math.go
def nextafter(x, y):  # Go math.Nextafter
    if math.isnan(x) or math.isnan(y):
        return float("nan")
    if x == y:
        return x
    if x == 0:
        return math.copysign(5e-324, y)
    b = struct.unpack("<q", struct.pack("<d", x))[0]
    if (y > x) == (x > 0):
        b += 1
    else:
        b -= 1
    return struct.unpack("<d", struct.pack("<q", b))[0]


def next_float_up(v):  # math.go:122-124
    return nextafter(v, v + 1)


def next_float_down(v):  # math.go:126-128
    return nextafter(v, v - 1)


MACHINE_EPSILON = next_float_up(0.0)       # math.go:17 (sic: the smallest denormal)
ONE_MINUS_EPSILON = next_float_down(1.0)   # math.go:18
PI = math.pi
INV_PI = 1.0 / PI                          # math.go:10
PI_OVER_2 = PI / 2.0
PI_OVER_4 = PI / 4.0


def go_max(x, y):  # Go math.Max
    if x == INF or y == INF:
        return INF
    if math.isnan(x) or math.isnan(y):
        return float("nan")
    if x == 0 and x == y:
        return y if math.copysign(1, x) < 0 else x
    return x if x > y else y


def go_min(x, y):  # Go math.Min
    if x == -INF or y == -INF:
        return -INF
    if math.isnan(x) or math.isnan(y):
        return float("nan")
    if x == 0 and x == y:
        return x if math.copysign(1, x) < 0 else y
    return x if x < y else y


def clamp(v, lo, hi):  # math.go:42-50
    if v < lo:
        return lo
    if v > hi:
        return hi
    return v


def gamma(n):  # math.go:82-84
    return (n * MACHINE_EPSILON) / (1 - n * MACHINE_EPSILON)


def radians(deg):  # math.go:130-132
    return PI / 180.0 * deg


def find_interval(size, pred):  # math.go:64-80
    first, length = 0, size
    while length > 0:
        half = length >> 1
        middle = first + half
        if pred(middle):
            first = middle + 1
            length -= half + 1
        else:
            length = half
    return int(clamp(float(first - 1), 0, float(size - 2)))


# ---------------------------------------------------------------- pkg/geometry/xyz.go:424-614
def v_add(a, b): return [a[0] + b[0], a[1] + b[1], a[2] + b[2]]
def v_sub(a, b): return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]
def v_muls(a, s): return [a[0] * s, a[1] * s, a[2] * s]
def v_divs(a, s): return [a[0] / s, a[1] / s, a[2] / s]
def v_abs(a): return [abs(a[0]), abs(a[1]), abs(a[2])]
def v_dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
def v_len2(a): return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]
def v_dist2(a, b): return v_len2(v_sub(b, a))   # xyz.go:579-581: other.Sub(self).LengthSquared()
def v_cross(a, b): return [(a[1] * b[2]) - (a[2] * b[1]), (a[2] * b[0]) - (a[0] * b[2]), (a[0] * b[1]) - (a[1] * b[0])]


def v_normalized(a):  # xyz.go:587-606: multiplies by 1/sqrt(lengthSquared) when it is > 0
    n2 = v_len2(a)
    if n2 > 0:
        inv = 1.0 / math.sqrt(n2)
        return [a[0] * inv, a[1] * inv, a[2] * inv]
    return list(a)


def face_forward(n, v):  # geometry.go:115-120
    return v_muls(n, -1.0) if v_dot(n, v) < 0.0 else list(n)


def coordinate_system(v1):  # geometry.go:47-60 — divides by the squared length, not the length
    if abs(v1[0]) > abs(v1[1]):
        v = v1[0] * v1[0] + v1[2] * v1[2]
        v2 = [-v1[2] / v, 0 / v, v1[0] / v]
    else:
        v = v1[1] * v1[1] + v1[2] * v1[2]
        v2 = [0 / v, v1[2] / v, -v1[1] / v]
    return v2, v_cross(v1, v2)


def spherical_direction_xyz(sinTheta, cosTheta, phi, x, y, z, cos, sin):  # geometry.go:66-70
    return v_add(v_add(v_muls(x, sinTheta * cos(phi)), v_muls(y, sinTheta * sin(phi))), v_muls(z, cosTheta))


# ---------------------------------------------------------------- pkg/pbrt/transform.go
def transform_point(m, p, pe):  # transform.go:227-247 (asymmetric error expression as written)
    xp = m[0][0] * p[0] + m[0][1] * p[1] + m[0][2] * p[2] + m[0][3]
    yp = m[1][0] * p[0] + m[1][1] * p[1] + m[1][2] * p[2] + m[1][3]
    zp = m[2][0] * p[0] + m[2][1] * p[1] + m[2][2] * p[2] + m[2][3]
    wp = m[3][0] * p[0] + m[3][1] * p[1] + m[3][2] * p[2] + m[3][3]
    g3 = gamma(3.0)
    err = []
    for r in range(3):
        err.append((g3 + 1.0) * (abs(m[r][0]) * pe[0] + abs(m[r][1]) * pe[1] + abs(m[r][2]) * pe[2]) +
                   (g3 * (abs(m[r][0] * p[0]) + abs(m[r][1]) * p[1] + abs(m[r][2] * p[2] + abs(m[r][3])))))
    q = [xp, yp, zp]
    if wp == 1.0:
        return q, err
    return v_divs(q, wp), err


def transform_normal(minv, n):  # transform.go:271-277: transpose of the inverse
    return [minv[0][0] * n[0] + minv[1][0] * n[1] + minv[2][0] * n[2],
            minv[0][1] * n[0] + minv[1][1] * n[1] + minv[2][1] * n[2],
            minv[0][2] * n[0] + minv[1][2] * n[1] + minv[2][2] * n[2]]


def transform_vector(m, v):  # transform.go:249-255
    return [m[r][0] * v[0] + m[r][1] * v[1] + m[r][2] * v[2] for r in range(3)]


def transform_ray(m, o, d):  # transform.go:279-300 (value part)
    o2, oerr = transform_point(m, o, [0.0, 0.0, 0.0])
    d2 = transform_vector(m, d)
    l2 = v_len2(d2)
    if l2 > 0:
        dt = v_dot(v_abs(d2), oerr) / l2
        o2 = v_add(o2, v_muls(d2, dt))
    return o2, d2


def translate(d):  # transform.go:347-362
    m = [[1.0, 0.0, 0.0, d[0]], [0.0, 1.0, 0.0, d[1]], [0.0, 0.0, 1.0, d[2]], [0.0, 0.0, 0.0, 1.0]]
    mi = [[1.0, 0.0, 0.0, -d[0]], [0.0, 1.0, 0.0, -d[1]], [0.0, 0.0, 1.0, -d[2]], [0.0, 0.0, 0.0, 1.0]]
    return m, mi


# ---------------------------------------------------------------- pkg/pbrt/ray.go:57-74
def offset_ray_origin(p, perr, n, w):
    d = v_dot(v_abs(n), perr) * 1024.0
    off = v_muls(n, d)
    if v_dot(w, n) < 0:
        off = v_muls(off, -1.0)
    po = v_add(p, off)
    for i in range(3):
        if off[i] > 0:
            po[i] = next_float_up(po[i])
        elif off[i] < 0:
            po[i] = next_float_down(po[i])
    return po


# ---------------------------------------------------------------- pkg/pbrt/reflection.go
def fr_dielectric(cosThetaI, etaI, etaT):  # reflection.go:21-42
    cosThetaI = clamp(cosThetaI, -1, 1)
    entering = cosThetaI > 0
    if not entering:
        etaI, etaT = etaT, etaI
        cosThetaI = abs(cosThetaI)
    sinThetaI = math.sqrt(go_max(0, 1 - cosThetaI * cosThetaI))
    sinThetaT = etaI / etaT * sinThetaI
    if sinThetaT >= 1:
        return 1.0
    cosThetaT = math.sqrt(go_max(0, 1 - sinThetaT * sinThetaT))
    Rparl = ((etaT * cosThetaI) - (etaI * cosThetaT)) / ((etaT * cosThetaI) + (etaI * cosThetaT))
    Rperp = ((etaI * cosThetaI) - (etaT * cosThetaT)) / ((etaI * cosThetaI) + (etaT * cosThetaT))
    return (Rparl * Rparl + Rperp * Rperp) / 2


def sin2theta(w): return go_max(0, 1 - w[2] * w[2])   # reflection.go:60-62
def sintheta(w): return math.sqrt(sin2theta(w))       # :64-66


def cosphi(w):  # reflection.go:76-83
    s = sintheta(w)
    return 1.0 if s == 0 else clamp(w[0] / s, -1, 1)


def sinphi(w):  # reflection.go:85-92
    s = sintheta(w)
    return 0.0 if s == 0 else clamp(w[1] / s, -1, 1)


def oren_nayar(R, sigma_deg, wo, wi):  # NewOrenNayar reflection.go:616-626 + F :628-652
    sigma = radians(sigma_deg)
    sigma2 = sigma * sigma
    a = 1.0 - (sigma2 / (2.0 * (sigma2 + 0.33)))
    b = 0.45 * sigma2 / (sigma2 * 0.09)   # sic: parsed as (0.45*sigma2) / (sigma2*0.09)
    sinThetaI, sinThetaO = sintheta(wi), sintheta(wo)
    maxCos = 0.0
    if sinThetaI > 1e-4 and sinThetaO > 1e-4:
        dCos = cosphi(wi) * cosphi(wo) + sinphi(wi) * sinphi(wo)
        maxCos = go_max(0.0, dCos)
    if abs(wi[2]) > abs(wo[2]):
        sinAlpha = sinThetaO
        tanBeta = sinThetaO / abs(wo[2])
    else:
        sinAlpha = sinThetaI
        tanBeta = sinThetaO / abs(wo[2])   # sic: wo in both branches
    k = INV_PI * (a + b * maxCos * sinAlpha * tanBeta)
    return [R[0] * k, R[1] * k, R[2] * k]


def refract(wi, n, eta):  # reflection.go:106-118
    cosThetaI = v_dot(n, wi)
    sin2ThetaI = go_max(0, 1 - cosThetaI * cosThetaI)
    sin2ThetaT = eta * eta * sin2ThetaI
    if sin2ThetaT >= 1:
        return None
    cosThetaT = math.sqrt(1 - sin2ThetaT)
    return v_add(v_muls(wi, -eta), v_muls(n, eta * cosThetaI - cosThetaT))


BSDF_REFLECTION, BSDF_TRANSMISSION, BSDF_DIFFUSE, BSDF_GLOSSY, BSDF_SPECULAR = 1, 2, 4, 8, 16


def fresnel_specular_sample_f(R, T, etaA, etaB, wo, u):  # reflection.go:482-523, mode == Radiance
    F = fr_dielectric(wo[2], etaA, etaB)
    if u[0] < F:
        wi = [-wo[0], -wo[1], wo[2]]
        k = abs(wi[2])
        return [R[0] * F / k, R[1] * F / k, R[2] * F / k], wi, F, BSDF_SPECULAR | BSDF_REFLECTION
    entering = wo[2] > 0
    etaI, etaT = (etaA, etaB) if entering else (etaB, etaA)
    wi = refract(wo, face_forward([0.0, 0.0, 1.0], wo), etaI / etaT)
    if wi is None:
        return [0.0, 0.0, 0.0], [0.0, 0.0, 0.0], 0.0, 0
    ft = [T[0] * (1 - F), T[1] * (1 - F), T[2] * (1 - F)]
    s = (etaI * etaI) / (etaT / etaT)   # sic
    ft = [c * s for c in ft]
    k = abs(wi[2])
    return [c / k for c in ft], wi, 1 - F, BSDF_SPECULAR | BSDF_TRANSMISSION


def concentric_sample_disk(u, cos, sin):  # sampling.go:173-192
    ox, oy = u[0] * 2.0 - 1, u[1] * 2.0 - 1
    if ox == 0 and oy == 0:
        return [0.0, 0.0]
    if abs(ox) > abs(oy):
        r = ox
        theta = PI_OVER_4 * (oy / ox)
    else:
        r = oy
        theta = PI_OVER_2 - PI_OVER_4 * (ox / oy)
    return [cos(theta) * r, sin(theta) * r]


def cosine_sample_hemisphere(u, cos, sin):  # sampling.go:194-198
    d = concentric_sample_disk(u, cos, sin)
    z = math.sqrt(go_max(0.0, 1.0 - d[0] * d[0] - d[1] * d[1]))
    return [d[0], d[1], z]


def lambert_sample_f(R, wo, u, cos, sin):
    """BSDF.SampleF (reflection.go:183-253) over ONE Lambertian lobe in the identity shading frame (wo local == world):
    sampleF :305-314 + LambertianReflection.F :589-591 + pdf :343-348; returns the LOCAL wi (wiWorld is discarded)"""
    z3 = [0.0, 0.0, 0.0]
    comp = go_min(math.floor(u[0] * 1.0), 1.0 - 1)
    ur = [go_min(u[0] * 1.0 - comp, ONE_MINUS_EPSILON), u[1]]
    if wo[2] == 0.0:
        return z3, z3, 0.0
    wi = cosine_sample_hemisphere(ur, cos, sin)
    if wo[2] < 0:
        wi[2] *= -1
    pdf = abs(wi[2]) * INV_PI if wo[2] * wi[2] > 0 else 0.0
    if pdf == 0.0:
        return z3, z3, 0.0
    return [R[0] * INV_PI, R[1] * INV_PI, R[2] * INV_PI], wi, pdf


def uniform_sample_sphere(u, cos, sin):  # sampling.go:158-163
    z = 1.0 - 2.0 * u[0]
    r = math.sqrt(go_max(0, 1 - z * z))
    phi = 2 * PI * u[1]
    return [r * cos(phi), r * sin(phi), z]


# ---------------------------------------------------------------- sampling.go:11-55
def distribution1d(f):
    n = len(f)
    cdf = [0.0] * (n + 1)
    for i in range(1, n + 1):
        cdf[i] = cdf[i - 1] + f[i - 1] / float(n)
    func_int = cdf[n]
    if func_int == 0.0:
        for i in range(1, n + 1):
            cdf[i] = float(i) / float(n)
    else:
        for i in range(1, n + 1):
            cdf[i] /= func_int
    return cdf, func_int


def sample_discrete(f, u):
    cdf, func_int = distribution1d(f)
    offset = find_interval(len(cdf), lambda i: cdf[i] <= u)
    pdf = 0.0
    if func_int > 0:
        pdf = f[offset] / (func_int / float(len(f)))
    return offset, pdf


# ---------------------------------------------------------------- shapes as light sources
def sphere_sample(radius, m, minv, reverse, u, cos, sin):  # sphere.go:270-285 (full sphere: Area = phiMax*r*(zMax-zMin))
    pObj = v_muls(uniform_sample_sphere(u, cos, sin), radius)
    n = v_normalized(transform_normal(minv, pObj))
    if reverse:
        n = v_muls(n, -1)
    pObj = v_muls(pObj, radius / math.sqrt(v_dist2(pObj, [0.0, 0.0, 0.0])))
    pObjError = v_muls(v_abs(pObj), gamma(5))
    p, perr = transform_point(m, pObj, pObjError)
    phiMax = radians(clamp(360.0, 0, 360))   # sphere.go:28-30
    area = phiMax * radius * (radius - (-radius))
    return p, perr, n, 1.0 / area


def sphere_sample_at(radius, m, minv, reverse, ref_p, ref_perr, ref_n, u, cos, sin):  # sphere.go:287-344
    pCenter, _ = transform_point(m, [0.0, 0.0, 0.0], [0.0, 0.0, 0.0])
    pOrigin = offset_ray_origin(ref_p, ref_perr, ref_n, v_sub(pCenter, ref_p))
    if v_dist2(pOrigin, pCenter) <= radius * radius:
        p, perr, n, pdf = sphere_sample(radius, m, minv, reverse, u, cos, sin)
        wi = v_sub(p, ref_p)
        if v_len2(wi) == 0:
            pdf = 0.0
        else:
            wi = v_normalized(wi)
            pdf *= v_dist2(ref_p, p) / abs(v_dot(n, v_muls(wi, -1)))
        if math.isinf(pdf):
            pdf = 0.0
        return p, perr, n, pdf
    wc = v_normalized(v_sub(pCenter, ref_p))
    wcX, wcY = coordinate_system(wc)
    r2 = radius * radius
    sinThetaMax2 = r2 / v_dist2(ref_p, pCenter)
    cosThetaMax = math.sqrt(go_max(0, 1.0 - sinThetaMax2))
    cosTheta = (1.0 - u[0]) + u[0] * cosThetaMax
    sinTheta = math.sqrt(go_max(0, 1 - cosTheta * cosTheta))
    phi = u[1] * 2 * PI
    dc = math.sqrt(v_dist2(ref_p, pCenter))
    ds = dc * cosTheta - math.sqrt(go_max(0, r2 - (dc * dc) * (sinTheta * sinTheta)))
    cosAlpha = (dc * dc + r2 - ds * ds) / (2.0 * dc * radius)
    sinAlpha = math.sqrt(go_max(0, 1.0 - cosAlpha * cosAlpha))
    nWorld = spherical_direction_xyz(sinAlpha, cosAlpha, phi, v_muls(wcX, -1), v_muls(wcY, -1), v_muls(wc, -1), cos, sin)
    pWorld = v_add(pCenter, v_muls(nWorld, radius))
    perr = v_muls(v_abs(pWorld), gamma(5.0))
    n = v_muls(nWorld, -1) if reverse else nWorld
    return pWorld, perr, n, 1.0 / (2.0 * PI * (1.0 - cosThetaMax))   # UniformConePdf sampling.go:169-171


def disk_sample_at(height, radius, inner, m, minv, ref_p, u, cos, sin):  # disk.go:160-170 + shape.go:50-65
    pd = concentric_sample_disk(u, cos, sin)
    pObj = [pd[0] * radius, pd[1] * radius, height]
    n = transform_normal(minv, [0.0, 0.0, 1.0])
    p, perr = transform_point(m, pObj, [0.0, 0.0, 0.0])
    phiMax = radians(clamp(360.0, 0, 360))
    pdf = 1 / (phiMax * 0.5 * (radius * radius - inner * inner))   # disk.go:183-185
    wi = v_sub(p, ref_p)
    if v_len2(wi) == 0.0:
        return p, perr, n, 0.0
    wi = v_normalized(wi)
    pdf *= v_dist2(ref_p, p) / abs(v_dot(n, v_muls(wi, -1)))
    if math.isinf(pdf):
        pdf = 0.0
    return p, perr, n, pdf


# ---------------------------------------------------------------- lights: point.go:44-49, distant.go:36-44, diffuse.go:36-59
def light_sample_li(light, ref_p, ref_perr, ref_n, u, cos, sin):
    """returns Li(3), wi(3), pdf, p1.p(3), p1.perr(3), p1.n(3), isDelta"""
    z3 = [0.0, 0.0, 0.0]
    if light["kind"] == "point":
        pl = light["p"]
        wi = v_normalized(v_sub(pl, ref_p))
        d2 = v_dist2(pl, ref_p)
        return [c / d2 for c in light["I"]], wi, 1.0, list(pl), z3, z3, 1
    if light["kind"] == "distant":
        w = light["w"]
        return list(light["L"]), list(w), 1.0, v_muls(w, 2 * light["world_radius"]), z3, z3, 1
    if light["shape"] == "sphere":
        p, perr, n, pdf = sphere_sample_at(light["radius"], light["m"], light["minv"], False, ref_p, ref_perr, ref_n, u, cos, sin)
    else:
        p, perr, n, pdf = disk_sample_at(light["height"], light["radius"], light["inner"], light["m"], light["minv"], ref_p, u, cos, sin)
    if pdf == 0 or v_len2(v_sub(p, ref_p)) == 0:
        return z3, z3, 0.0, p, perr, n, 0
    wi = v_sub(p, ref_p)   # un-normalised (diffuse.go:55)
    w = v_muls(wi, -1)
    Li = list(light["L"]) if (light["two_sided"] or v_dot(n, w) > 0) else z3
    return Li, wi, pdf, p, perr, n, 0


# ---------------------------------------------------------------- film.go:106-113, 211-248 (box filter: table of ones), spectrum.go:35-41
def film_tile_add_sample(tile_bounds, frx, fry, cropped, pFilm, L):
    """tile_bounds = sample bounds (x0, y0, x1, y1) of the tile; returns (pixel bounds, {(x, y): [r, g, b, w]})"""
    x0, y0, x1, y1 = tile_bounds
    pb = [max(cropped[0], math.ceil(x0 - 0.5 - frx)), max(cropped[1], math.ceil(y0 - 0.5 - fry)),
          min(cropped[2], math.floor(x1 - 0.5 + frx) + 1), min(cropped[3], math.floor(y1 - 0.5 + fry) + 1)]
    dx, dy = pFilm[0] - 0.5, pFilm[1] - 0.5
    p0 = [int(go_max(float(math.ceil(dx - frx)), float(pb[0]))), int(go_max(float(math.ceil(dy - fry)), float(pb[1])))]
    p1 = [int(go_min(float(math.floor(dx + frx)) + 1, float(pb[2]))), int(go_min(float(math.floor(dy + fry)) + 1, float(pb[3])))]
    px = {}
    for y in range(p0[1], p1[1]):
        for x in range(p0[0], p1[0]):
            fw = 1.0  # BoxFilter.Evaluate (filter.go:30-32)
            px[(x, y)] = [L[0] * (1.0 * fw), L[1] * (1.0 * fw), L[2] * (1.0 * fw), fw]
    return pb, px


def rgb_to_xyz(c):  # spectrum.go:35-41
    return [0.412453 * c[0] + 0.357580 * c[1] + 0.180423 * c[2],
            0.212671 * c[0] + 0.715160 * c[1] + 0.072169 * c[2],
            0.019334 * c[0] + 0.119193 * c[1] + 0.950227 * c[2]]


# ---------------------------------------------------------------- rng.go:28-57, sampling.go:101-146, stratified.go:21-48
M64 = (1 << 64) - 1


class Rng:
    def __init__(self):
        self.state, self.inc = 0x853c49e6748fea9b, 0xda3e39cb94b95bdb   # rng.go:9-13

    def set_sequence(self, seed):
        self.state = 0
        self.inc = ((seed << 1) | 1) & M64
        self.u32()
        self.state = (self.state + 0x853c49e6748fea9b) & M64
        self.u32()

    def u32(self):
        old = self.state
        self.state = (old * 0x5851f42d4c957f2d + self.inc) & M64
        xs = (((old >> 18) ^ old) >> 27) & 0xffffffff
        rot = old >> 59
        return ((xs >> rot) | (xs << ((rot + 1) & 31))) & 0xffffffff   # sic: not a rotate

    def u32b(self, b):
        threshold = ((~b + 1) & 0xffffffff) % b
        while True:
            r = self.u32()
            if r >= threshold:
                return r % b

    def uniform(self):
        return go_min(ONE_MINUS_EPSILON, float(self.u32()) * 2.3283064365386963e-10)


def stratified_start_pixel(rng, nx, ny, jitter, ndims):
    """the 1-D tables after Stratified.StartPixel; the 2-D tables stay all zeros (sampling.go:122-124 writes a copy)"""
    n = nx * ny
    t1 = []
    for _ in range(ndims):
        inv = 1.0 / float(n)
        samp = []
        for i in range(n):
            delta = rng.uniform() if jitter else 0.5
            samp.append(go_min((float(i) + delta) * inv, ONE_MINUS_EPSILON))
        for i in range(n):
            other = i + rng.u32b(n - i)
            samp[i], samp[other] = samp[other], samp[i]
        t1.append(samp)
    for _ in range(ndims):
        for _y in range(ny):
            for _x in range(nx):
                if jitter:
                    rng.uniform(); rng.uniform()
        for i in range(n):
            rng.u32b(n - i)
    return t1


# ---------------------------------------------------------------- camera.go:192-242 (ray only)
def camera_ray(r2c, c2w, lens_radius, focal, pFilm, pLens, cos, sin):
    pCamera, _ = transform_point(r2c, [pFilm[0], pFilm[1], 0.0], [0.0, 0.0, 0.0])
    o, d = [0.0, 0.0, 0.0], v_normalized(pCamera)
    if lens_radius > 0:
        pl = concentric_sample_disk(pLens, cos, sin)
        pl = [pl[0] * lens_radius, pl[1] * lens_radius]
        ft = focal / d[2]
        pFocus = v_add(v_muls(d, ft), o)
        o = [pl[0], pl[1], 0.0]
        d = v_normalized(v_sub(pFocus, o))
    return transform_ray(c2w, o, d)


def spawn_ray_to(p, perr, n, q, qerr, qn):  # interaction.go:91-102: origin stays the un-offset point
    origin = offset_ray_origin(p, perr, n, v_sub(q, p))
    target = offset_ray_origin(q, qerr, qn, v_sub(origin, q))
    return list(p), v_sub(target, origin)


# ================================================================ the vectors
def hx(v):
    if isinstance(v, (list, tuple)):
        return [hx(x) for x in v]
    return float(v).hex()


def main():
    ex = (lambda x: {0.0: 1.0}[x], lambda x: {0.0: 0.0}[x])   # cos, sin that only accept the exact argument 0
    gm = (gomath.Cos, gomath.Sin)
    cases = []

    def add(fn, trig, inp, out, cite):
        cases.append(dict(fn=fn, trig=trig, cite=cite, **{"in": hx(inp), "out": hx(out)}))

    for c, ei, et in [(1.0, 1.0, 1.5), (0.5, 1.0, 1.5), (-0.5, 1.0, 1.5), (0.1, 1.0, 1.5), (-0.7453559924999299, 1.0, 1.5), (-0.3, 1.0, 1.5),
                      (0.0, 1.0, 1.5), (1.7, 1.0, 1.33), (0.9999999999999999, 1.0, 2.4), (1e-9, 1.5, 1.0)]:
        add("fr_dielectric", "none", [c, ei, et], [fr_dielectric(c, ei, et)], "reflection.go:21-42")

    wos = [[0.3, 0.4, 0.8660254037844386], [0.0, 0.0, 1.0], [-0.6, 0.1, -0.7937253933193772], [0.7, -0.7, 0.14142135623730964], [1e-5, 0.0, 0.99999999995]]
    wis = [[-0.2, 0.5, 0.8426149773176359], [0.5, 0.5, 0.7071067811865476], [0.0, 0.0, 1.0], [0.9, 0.0, -0.4358898943540673]]
    for sg in (20.0, 90.0, 0.5):
        for wo in wos:
            for wi in wis:
                R = [0.5, 0.25, 0.125]
                add("oren_nayar_f", "none", [sg] + R + wo + wi, oren_nayar(R, sg, wo, wi), "reflection.go:616-652")

    for wo in wos + [[0.0, 0.6, -0.8], [0.99, 0.0, 0.1410673597566272]]:
        for u in ([0.0, 0.3], [0.03, 0.3], [0.5, 0.9], [0.9999, 0.1]):
            R, T, eb = [1.0, 0.9, 0.8], [0.7, 1.0, 0.95], 1.5
            f, wi, pdf, st = fresnel_specular_sample_f(R, T, 1.0, eb, wo, u)
            add("fresnel_specular_sample_f", "none", R + T + [eb] + wo + u, f + wi + [pdf, float(st)], "reflection.go:482-523,106-118")

    for trig, tr, us in (("exact", ex, [[0.5, 0.5]]), ("gomath", gm, [[0.25, 0.75], [0.9, 0.1], [0.5, 0.2], [0.0, 0.0], [0.3, 0.5], [0.999, 0.999]])):
        for u in us:
            add("concentric_sample_disk", trig, u, concentric_sample_disk(u, *tr), "sampling.go:173-192")
            add("cosine_sample_hemisphere", trig, u, cosine_sample_hemisphere(u, *tr), "sampling.go:194-198")
            for wo in ([0.1, 0.2, 0.9746794344808963], [0.1, 0.2, -0.9746794344808963]):
                R = [0.73, 0.5, 0.1]
                f, wi, pdf = lambert_sample_f(R, wo, u, *tr)
                add("lambert_sample_f", trig, R + wo + u, f + wi + [pdf], "reflection.go:183-253,305-314,343-348,589-591")

    for p, pe, n, w in [([0, 0, 0], [5e-324] * 3, [1, 1, 1], [1, 1, 1]), ([1.5, -2.25, 3.0], [1e-320, 0.0, 2e-322], [0.0, 0.6, -0.8], [0.3, -1.0, 0.2]),
                        ([10.0, 10.0, 10.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, -1.0, 0.0]), ([-1.0, 0.99, 0.25], [5e-323, 1e-322, 5e-324], [0.0, -1.0, 0.0], [0.1, -0.5, 0.0])]:
        p, pe, n, w = [float(x) for x in p], [float(x) for x in pe], [float(x) for x in n], [float(x) for x in w]
        add("offset_ray_origin", "none", p + pe + n + w, offset_ray_origin(p, pe, n, w), "ray.go:57-74")

    for v in ([0.0, 0.0, 1.0], [0.6, 0.0, 0.8], [0.2, -0.7, 0.6855654600401044], [-0.9, 0.1, -0.4242640687119285]):
        v2, v3 = coordinate_system(v)
        add("coordinate_system", "none", v, v2 + v3, "geometry.go:47-60")

    for nl in (1, 2, 4, 9):
        for u in (0.0, 0.24999999999999997, 0.25, 0.5, 0.7, 0.9999999999999999):
            off, pdf = sample_discrete([1.0] * nl, u)
            add("sample_discrete_uniform", "none", [float(nl), u], [float(off), pdf], "sampling.go:11-55, math.go:64-80, lightdistribution.go:25-34")

    for c in ([1.0, 1.0, 1.0], [0.25, 0.5, 4.0], [17.0, 12.0, 4.0], [1e-300, 0.0, 3.5]):
        add("rgb_to_xyz", "none", c, rgb_to_xyz(c), "spectrum.go:35-41")

    # FilmTile.AddSample: film 8x6, box radius (0.5, 0.5) and (1.0, 2.0), tileSize 1 and 4
    for (W, H, ts, frx, fry, tile, pf, L) in [(8, 6, 1, 0.5, 0.5, 11, [3.0, 1.0], [0.5, 2.0, 0.125]), (8, 6, 1, 0.5, 0.5, 0, [0.0, 0.0], [1.0, 1.0, 1.0]),
                                               (8, 6, 4, 0.5, 0.5, 1, [6.0, 3.0], [0.1, 0.2, 0.3]), (8, 6, 4, 1.0, 2.0, 3, [7.25, 4.75], [3.0, 0.0, 1.0]),
                                               (8, 6, 1, 0.5, 0.5, 47, [7.0, 5.0], [2.0, 2.0, 2.0]), (8, 6, 4, 0.5, 0.5, 0, [2.5, 2.5], [1.0, 0.5, 0.25])]:
        ntx = (W + ts - 1) // ts
        tx, ty = tile % ntx, tile // ntx
        x0, y0 = tx * ts, ty * ts
        tb = (x0, y0, min(x0 + ts, W), min(y0 + ts, H))
        pb, px = film_tile_add_sample(tb, frx, fry, (0, 0, W, H), pf, L)
        flat = []
        for (x, y), v in sorted(px.items(), key=lambda kv: (kv[0][1], kv[0][0])):
            flat += [float(x), float(y)] + v
        add("film_add_sample", "none", [float(W), float(H), float(ts), frx, fry, float(tile)] + pf + L, [float(b) for b in pb] + flat, "film.go:106-113,211-248, filter.go:30-32")

    r = Rng()
    add("rng_u32", "none", [0.0, 0.0, 8.0], [float(r.u32()) for _ in range(8)], "rng.go:9-13,36-42 (default state)")
    for seed in (0, 1, 7, 119, 2073599):
        r = Rng(); r.set_sequence(seed)
        add("rng_u32", "none", [1.0, float(seed), 8.0], [float(r.u32()) for _ in range(8)], "rng.go:28-42")
        r = Rng(); r.set_sequence(seed)
        add("rng_uniform", "none", [float(seed), 6.0], [r.uniform() for _ in range(6)], "rng.go:55-57")
        r = Rng(); r.set_sequence(seed)
        add("rng_u32b", "none", [float(seed), 16.0, 6.0], [float(r.u32b(16 - k)) for k in range(6)], "rng.go:44-53")
    for seed, nx, ny, jit, nd in ((0, 2, 2, 0, 2), (5, 4, 4, 0, 4), (5, 4, 4, 1, 4), (77, 3, 2, 1, 3)):
        r = Rng(); r.set_sequence(seed)
        t1 = stratified_start_pixel(r, nx, ny, bool(jit), nd)
        after = [r.uniform(), r.uniform()]
        add("stratified_start_pixel", "none", [float(seed), float(nx), float(ny), float(jit), float(nd)], [v for t in t1 for v in t] + after,
            "stratified.go:21-48, sampling.go:101-146 (1-D tables, then the next two raw RNG floats)")

    # lights of a small fixed scene (tests/test_shading_kats.py builds the same one through the host mirror)
    tm, tmi = translate([2.0, 5.0, -1.0])
    lights = [dict(kind="point", p=[1.0, 4.0, 2.0], I=[10.0, 20.0, 30.0]),
              dict(kind="distant", w=[0.0, 0.6, 0.8], L=[0.5, 0.25, 1.0], world_radius=None),
              dict(kind="area", shape="sphere", radius=0.75, m=tm, minv=tmi, L=[5.0, 4.0, 3.0], two_sided=False)]
    # disk light of config 2: o2w = Translate(0, .99, 0) * RotateX(90) with Go's sin/cos (host-side constructor, transform.go:381-394)
    s90, c90 = gomath.Sin(radians(90.0)), gomath.Cos(radians(90.0))
    rx = [[1.0, 0.0, 0.0, 0.0], [0.0, c90, -s90, 0.0], [0.0, s90, c90, 0.0], [0.0, 0.0, 0.0, 1.0]]
    rxi = [[rx[j][i] for j in range(4)] for i in range(4)]
    t2, t2i = translate([0.0, 0.99, 0.0])

    def mmul_plain(a, b):
        return [[sum(a[i][k] * b[k][j] for k in range(4)) for j in range(4)] for i in range(4)]
    # the matrices are INPUT data of the KAT (the quirky Matrix4x4.Mul of the host is not under test here)
    dm, dmi = mmul_plain(t2, rx), mmul_plain(rxi, t2i)
    lights.append(dict(kind="area", shape="disk", height=0.0, radius=0.25, inner=0.0, m=dm, minv=dmi, L=[17.0, 12.0, 4.0], two_sided=False))
    lights.append(dict(kind="area", shape="disk", height=0.0, radius=0.25, inner=0.0, m=dm, minv=dmi, L=[17.0, 12.0, 4.0], two_sided=True))
    refs = [([0.5, 0.0, 0.5], [1e-322, 0.0, 5e-324], [0.0, 1.0, 0.0]), ([-0.9, 0.2, 0.3], [0.0, 2e-323, 1e-323], [1.0, 0.0, 0.0]),
            ([2.1, 5.2, -1.1], [5e-324, 5e-324, 5e-324], [0.0, 0.0, 1.0]), ([0.25, 0.995, 0.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0])]
    for li, l in enumerate(lights):
        for (rp, re, rn) in refs:
            for trig, tr, us in (("exact", ex, [[0.5, 0.0], [0.0, 0.0], [0.9, 0.0]]), ("gomath", gm, [[0.3, 0.6], [0.75, 0.25], [0.99, 0.9]])):
                for u in us:
                    if l["kind"] == "distant":
                        continue   # needs the scene's world radius: covered by the scene-level case below
                    try:
                        Li, wi, pdf, p1, p1e, p1n, delta = light_sample_li(l, rp, re, rn, u, *tr)
                    except KeyError:
                        continue   # the exact-trig table met a non-zero angle: not an "exact" case
                    add("light_sample_li", trig, [float(li)] + rp + re + rn + u, Li + wi + [pdf] + p1 + p1e + p1n + [float(delta)],
                        "point.go:44-49, diffuse.go:36-59, sphere.go:270-344, disk.go:160-170, shape.go:50-65")

    for (p, pe, n, q, qe, qn) in [([0.0, 0.0, 0.0], [0.0] * 3, [0.0] * 3, [10.0, 0.0, 0.0], [0.0] * 3, [0.0] * 3),
                                  ([0.5, 0.0, 0.5], [1e-322, 0.0, 5e-324], [0.0, 1.0, 0.0], [0.1, 0.99, -0.1], [0.0, 3e-323, 0.0], [0.0, -1.0, 0.0])]:
        o, d = spawn_ray_to(p, pe, n, q, qe, qn)
        add("spawn_ray_to", "none", p + pe + n + q + qe + qn, o + d + [1 - 0.0001], "interaction.go:91-102 (light_test.go:10-44 pins the first case)")

    # camera: config-1-like off-axis matrices are host data; here a plain pinhole + thin lens
    r2c = [[0.001, 0.0, 0.0, -0.96], [0.0, -0.001, 0.0, 0.54], [0.0, 0.0, 0.0, 1.0], [0.0, 0.0, 0.0, 1.0]]
    c2w = [[1.0, 0.0, 0.0, 0.5], [0.0, 0.8, -0.6, 1.0], [0.0, 0.6, 0.8, -3.0], [0.0, 0.0, 0.0, 1.0]]
    for lr, fd, pf, plens, trig, tr in [(0.0, 1e6, [0.0, 0.0], [0.0, 0.0], "none", ex), (0.0, 1e6, [960.0, 540.0], [0.0, 0.0], "none", ex),
                                         (0.0, 1e6, [1919.0, 3.0], [0.0, 0.0], "none", ex), (0.05, 4.0, [100.0, 900.0], [0.5, 0.5], "exact", ex),
                                         (0.05, 4.0, [100.0, 900.0], [0.8, 0.3], "gomath", gm)]:
        o, d = camera_ray(r2c, c2w, lr, fd, pf, plens, *tr)
        add("camera_ray", trig, [x for row in r2c for x in row] + [x for row in c2w for x in row] + [lr, fd] + pf + plens, o + d,
            "camera.go:192-242, transform.go:227-300")

    # math.Max / math.Min with their special cases and math.Sin / math.Cos of one argument (the device also
    # evaluates its fused sin+cos, which must equal the two separate functions): out = [Max, Min, Sin(x), Cos(x), Sin(x), Cos(x)]
    inf, nan = float("inf"), float("nan")
    specials = [0.0, -0.0, 1.0, -1.0, inf, -inf, nan, 5e-324, -5e-324, 2.5, 1e300]
    for x in specials:
        for y in specials:
            add("go_math", "gomath", [x, y], [go_max(x, y), go_min(x, y), gomath.Sin(x), gomath.Cos(x), gomath.Sin(x), gomath.Cos(x)],
                "go1.11 src/math/dim.go (Max, Min), sin.go (Sin, Cos)")
    for x in (0.1, -0.1, 0.7853981633974483, 0.7853981633974484, 1.5707963267948966, 2.356194490192345, 3.141592653589793, 4.0, -5.5,
              6.283185307179586, 6.283185307179587, 100.0, -12345.678, 1e9, 536870912.0, 0.3926990816987241, 1e-300, 2.2250738585072014e-308):
        add("go_math", "gomath", [x, -x], [go_max(x, -x), go_min(x, -x), gomath.Sin(x), gomath.Cos(x), gomath.Sin(x), gomath.Cos(x)],
            "go1.11 src/math/dim.go (Max, Min), sin.go (Sin, Cos)")

    out = dict(note="generated by tests/golden/make_shading_kats.py — independent plain-Python restatement of the cited Go lines; values are C99 hex floats",
               lights=[{k: (hx(v) if isinstance(v, list) else v) for k, v in l.items()} for l in lights], cases=cases)
    with open(os.path.join(HERE, "shading_kats.json"), "w") as f:
        json.dump(out, f, indent=0)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
