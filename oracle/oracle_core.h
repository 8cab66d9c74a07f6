// ORACLE — test infrastructure only.  Nothing under go-pbrt_b200/ may include, link or call this.
//
// oracle_core.h: CPU restatement (C++17, -O2 -ffp-contract=off, no fast-math) of the reference's
// geometry / efloat / transform / shape / BVH code on the hot path.
// Each function cites the reference file:line it follows.  Quirks of the reference
// are reproduced, not fixed.
//
// PARITY PIN: the Go toolchain is absent, so the reference cannot be executed here.  This restatement is
// pinned against every golden value the reference's own tests hold for this path (tests/test_oracle_golden.py):
//   pkg/efloat/efloat_test.go:9-13, pkg/pbrt/ray_test.go:10-19, pkg/pbrt/transform_test.go:17-36,66-81,
//   pkg/pbrt/light_test.go:10-44, pkg/accelerator/simple_test.go:40-108, pkg/accelerator/bvh_test.go:43-141.
// (round 2 added pkg/geometry/xyz_test.go, spectrum_test.go, reflection_test.go:9-15, bvh_test.go:143-264.)
// Path.Li, DirectLighting.Li, samplers, RNG, BSDFs, lights, camera, film have NO reference-held golden; they are pinned by
// INDEPENDENT plain-Python restatements of the Go source instead: per function (tests/golden/make_shading_kats.py) and
// composed — whole films of configs 1 and 2, sphere / mixed / partial-shape scenes, both integrators, both samplers and
// FAST mode (tests/golden/make_*_golden.py) — which this oracle reproduces bit for bit, ray counts included
// (tests/test_*_golden.py).  Triangles do not exist in the reference ⇒ "parity unpinned", defined here.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../include/gopbrt_cuda.h"
#include "gomath.h"

namespace oracle {
namespace gm = gomath;

// ---------------------------------------------------------------- pkg/geometry/xyz.go:424-614
struct V3 {
  double x = 0, y = 0, z = 0;
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }  // xyz.go:428-437
  void set(int i, double v) { if (i == 0) x = v; else if (i == 1) y = v; else z = v; }
};
static inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 mul(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 muls(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 divs(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
static inline V3 vabs(V3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // xyz.go:561-563
static inline double absdot(V3 a, V3 b) { return std::fabs(dot(a, b)); }
static inline double len2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
static inline double length(V3 a) { return std::sqrt(len2(a)); }
// xyz.go:579-581: DistanceSquared = (other - xyz).LengthSquared()
static inline double dist2(V3 self, V3 other) { return len2(sub(other, self)); }
static inline double dist(V3 self, V3 other) { return std::sqrt(dist2(self, other)); }
static inline V3 cross(V3 a, V3 b) {  // xyz.go:583-585
  return {(a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)};
}
static inline V3 normalized(V3 a) {  // xyz.go:587-606: multiply by 1/sqrt
  double n2 = len2(a);
  if (n2 > 0) {
    double inv = 1.0 / std::sqrt(n2);
    a.x *= inv; a.y *= inv; a.z *= inv;
  }
  return a;
}
// pkg/pbrt/geometry.go:115-120
static inline V3 faceforward(V3 n1, V3 n2) { return dot(n1, n2) < 0.0 ? muls(n1, -1) : n1; }
// pkg/pbrt/geometry.go:47-60 — divides by the SQUARED length
static inline void coordinate_system(V3 v1, V3* v2, V3* v3) {
  if (std::fabs(v1.x) > std::fabs(v1.y)) {
    double v = v1.x * v1.x + v1.z * v1.z;
    *v2 = V3{-v1.z / v, 0 / v, v1.x / v};
  } else {
    double v = v1.y * v1.y + v1.z * v1.z;
    *v2 = V3{0 / v, v1.z / v, -v1.y / v};
  }
  *v3 = cross(v1, *v2);
}

// ---------------------------------------------------------------- pkg/efloat
struct Counters {
  std::atomic<uint64_t> efloat_panics{0};
  std::atomic<uint64_t> stack_overflows{0};
};
static Counters g_counters;

struct EF { double v, lo, hi; };
// efloat.go:102-111 — the reference panics; the oracle counts and carries on
static inline void ef_check(const EF& f) {
  if (std::isinf(f.lo) || std::isnan(f.lo) || std::isinf(f.hi) || std::isnan(f.hi) || f.lo > f.hi)
    g_counters.efloat_panics.fetch_add(1, std::memory_order_relaxed);
}
static inline EF ef_new(double v, double err) {  // efloat.go:10-22
  EF f{v, v, v};
  if (err != 0) {
    f.lo = gm::NextFloatDown(v - err);
    f.hi = gm::NextFloatUp(v + err);
  }
  ef_check(f);
  return f;
}
static inline EF ef_add(EF f, EF o) {  // efloat.go:35-40
  f.v = f.v + o.v;
  f.lo = gm::NextFloatDown(f.lo + o.lo);
  f.hi = gm::NextFloatUp(f.hi + o.hi);
  ef_check(f);
  return f;
}
static inline EF ef_sub(EF f, EF o) {  // efloat.go:95-100
  f.v = f.v - o.v;
  f.lo = gm::NextFloatDown(f.lo - o.hi);
  f.hi = gm::NextFloatUp(f.hi - o.lo);
  ef_check(f);
  return f;
}
static inline EF ef_mul(EF f, EF o) {  // efloat.go:73-86
  double p0 = f.lo * o.lo, p1 = f.hi * o.lo, p2 = f.lo * o.hi, p3 = f.hi * o.hi;
  f.v = f.v * o.v;
  f.lo = gm::NextFloatDown(gm::Min(gm::Min(p0, p1), gm::Min(p2, p3)));
  f.hi = gm::NextFloatUp(gm::Max(gm::Max(p0, p1), gm::Max(p2, p3)));
  ef_check(f);
  return f;
}
static inline EF ef_div(EF f, EF o) {  // efloat.go:47-66
  f.v = f.v / o.v;
  if (o.lo < 0 && o.hi > 0) {
    f.lo = -gm::Inf;
    f.hi = gm::Inf;
  } else {
    double d0 = f.lo / o.lo, d1 = f.hi / o.lo, d2 = f.lo / o.hi, d3 = f.hi / o.hi;
    f.lo = gm::NextFloatDown(gm::Min(gm::Min(d0, d1), gm::Min(d2, d3)));
    f.hi = gm::NextFloatUp(gm::Max(gm::Max(d0, d1), gm::Max(d2, d3)));
  }
  ef_check(f);
  return f;
}
static inline EF ef_muls(EF f, double s) { return ef_mul(f, ef_new(s, 0.0)); }  // efloat.go:88-90
// efloat/math.go:35-59
static inline bool ef_quadratic(EF a, EF b, EF c, EF* t0, EF* t1) {
  double disc = b.v * b.v - 4. * a.v * c.v;
  if (disc < 0) return false;
  double root = std::sqrt(disc);
  EF fr = ef_new(root, gm::MachineEpsilon() * root);
  EF q = (b.v < 0) ? ef_muls(ef_sub(b, fr), -0.5) : ef_muls(ef_add(b, fr), -0.5);
  *t0 = ef_div(q, a);
  *t1 = ef_div(c, q);
  if (t0->v > t1->v) std::swap(*t0, *t1);
  return true;
}

// ---------------------------------------------------------------- pkg/pbrt/transform.go
struct M4 { double m[4][4]; };
struct Xf { M4 m, inv; };
static inline Xf xf_from(const gopbrt_transform& t) {
  Xf x;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { x.m.m[i][j] = t.m[i * 4 + j]; x.inv.m[i][j] = t.minv[i * 4 + j]; }
  return x;
}
static inline Xf xf_inverse(const Xf& t) { return Xf{t.inv, t.m}; }  // transform.go:175-177
static inline bool xf_is_identity(const Xf& t) {  // transform.go:167-173 (Matrix only)
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) if (t.m.m[i][j] != (i == j ? 1.0 : 0.0)) return false;
  return true;
}
// transform.go:227-247 — note the asymmetric error terms
static inline V3 xf_point(const Xf& t, V3 p, V3 pe, V3* err) {
  const double(*m)[4] = t.m.m;
  double xp = m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z + m[0][3];
  double yp = m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z + m[1][3];
  double zp = m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z + m[2][3];
  double wp = m[3][0] * p.x + m[3][1] * p.y + m[3][2] * p.z + m[3][3];
  double g3 = gm::Gamma(3);
  if (err) {
    double g31 = gm::Gamma(3.0) + 1.0;
    err->x = g31 * (std::fabs(m[0][0]) * pe.x + std::fabs(m[0][1]) * pe.y + std::fabs(m[0][2]) * pe.z) +
             (g3 * (std::fabs(m[0][0] * p.x) + std::fabs(m[0][1]) * p.y + std::fabs(m[0][2] * p.z + std::fabs(m[0][3]))));
    err->y = g31 * (std::fabs(m[1][0]) * pe.x + std::fabs(m[1][1]) * pe.y + std::fabs(m[1][2]) * pe.z) +
             (g3 * (std::fabs(m[1][0] * p.x) + std::fabs(m[1][1]) * p.y + std::fabs(m[1][2] * p.z + std::fabs(m[1][3]))));
    err->z = g31 * (std::fabs(m[2][0]) * pe.x + std::fabs(m[2][1]) * pe.y + std::fabs(m[2][2]) * pe.z) +
             (g3 * (std::fabs(m[2][0] * p.x) + std::fabs(m[2][1]) * p.y + std::fabs(m[2][2] * p.z + std::fabs(m[2][3]))));
  }
  V3 np{xp, yp, zp};
  if (wp == 1.0) return np;
  return divs(np, wp);
}
static inline V3 xf_vector(const Xf& t, V3 v) {  // transform.go:249-255
  const double(*m)[4] = t.m.m;
  return {m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z, m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
          m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z};
}
static inline V3 xf_vector_err(const Xf& t, V3 v, V3* err) {  // transform.go:257-269
  const double(*m)[4] = t.m.m;
  double g3 = gm::Gamma(3);
  err->x = g3 * (std::fabs(m[0][0] * v.x) + std::fabs(m[0][1] * v.y) + std::fabs(m[0][2] * v.z));
  err->y = g3 * (std::fabs(m[1][0] * v.x) + std::fabs(m[1][1] * v.y) + std::fabs(m[1][2] * v.z));
  err->z = g3 * (std::fabs(m[2][0] * v.x) + std::fabs(m[2][1] * v.y) + std::fabs(m[2][2] * v.z));
  return xf_vector(t, v);
}
static inline V3 xf_normal(const Xf& t, V3 n) {  // transform.go:271-277 (transpose of the inverse)
  const double(*i)[4] = t.inv.m;
  return {i[0][0] * n.x + i[1][0] * n.y + i[2][0] * n.z, i[0][1] * n.x + i[1][1] * n.y + i[2][1] * n.z,
          i[0][2] * n.x + i[1][2] * n.y + i[2][2] * n.z};
}

// pkg/pbrt/ray.go:5-15 (differentials are dropped by Path.Li, path.go:35 / ray.go:37-51)
struct Ray { V3 o, d; double tmax = gm::Inf; double time = 0; };
static inline V3 ray_at(const Ray& r, double t) { return add(r.o, muls(r.d, t)); }  // ray.go:53-55

// transform.go:279-300
static inline Ray xf_ray(const Xf& t, const Ray& r, V3* oerr, V3* derr) {
  V3 oe, de;
  V3 o = xf_point(t, r.o, V3{}, &oe);
  V3 d = xf_vector_err(t, r.d, &de);
  double l2 = len2(d);
  if (l2 > 0) {
    double dt = dot(vabs(d), oe) / l2;
    o = add(o, muls(d, dt));
  }
  if (oerr) *oerr = oe;
  if (derr) *derr = de;
  return Ray{o, d, r.tmax, r.time};
}

// ---------------------------------------------------------------- pkg/pbrt/bounds.go
struct B3 { V3 mn, mx; bool valid = false; };
static inline V3 min_point(V3 a, V3 b) { return {gm::Min(a.x, b.x), gm::Min(a.y, b.y), gm::Min(a.z, b.z)}; }  // geometry.go:74-80
static inline V3 max_point(V3 a, V3 b) { return {gm::Max(a.x, b.x), gm::Max(a.y, b.y), gm::Max(a.z, b.z)}; }
static inline void b3_union_point(B3& b, V3 p) {  // bounds.go:209-219
  if (!b.valid) { b.mn = p; b.mx = p; b.valid = true; }
  b.mn = min_point(b.mn, p);
  b.mx = max_point(b.mx, p);
}
static inline void b3_union(B3& b, const B3& o) {  // bounds.go:221-238
  if (!o.valid) return;
  if (!b.valid) { b = o; }
  b.mn = min_point(b.mn, o.mn);
  b.mx = max_point(b.mx, o.mx);
}
static inline V3 b3_corner(const B3& b, int c) {  // bounds.go:114-120
  return {(c & 1) ? b.mx.x : b.mn.x, (c & 2) ? b.mx.y : b.mn.y, (c & 4) ? b.mx.z : b.mn.z};
}
static inline B3 xf_bounds(const Xf& t, const B3& b) {  // transform.go:336-345
  B3 r;
  b3_union_point(r, xf_point(t, b.mn, V3{}, nullptr));
  for (int i = 1; i < 8; i++) b3_union_point(r, xf_point(t, b3_corner(b, i), V3{}, nullptr));
  return r;
}
static inline double b3_area(const B3& b) {  // bounds.go:133-136
  V3 d = sub(b.mx, b.mn);
  return 2 * (d.x * d.y + d.x * d.z + d.y * d.z);
}
static inline int b3_max_extent(const B3& b) {  // bounds.go:138-147
  V3 d = sub(b.mx, b.mn);
  if (d.x > d.y && d.x > d.z) return 0;
  if (d.y > d.z) return 1;
  return 2;
}
static inline V3 b3_offset(const B3& b, V3 p) {  // bounds.go:195-207
  V3 o = sub(p, b.mn);
  if (b.mx.x > b.mn.x) o.x /= b.mx.x - b.mn.x;
  if (b.mx.y > b.mn.y) o.y /= b.mx.y - b.mn.y;
  if (b.mx.z > b.mn.z) o.z /= b.mx.z - b.mn.z;
  return o;
}
// bounds.go:149-185 — the gamma factor is exactly 1.0; keep the if-structure (NaN semantics)
static inline bool b3_intersect_p(const B3& b, const Ray& r, V3 invd, const int neg[3]) {
  double g = 1 + 2 * gm::Gamma(3);
  double tMin = ((neg[0] ? b.mx.x : b.mn.x) - r.o.x) * invd.x;
  double tMax = ((neg[0] ? b.mn.x : b.mx.x) - r.o.x) * invd.x;
  double tyMin = ((neg[1] ? b.mx.y : b.mn.y) - r.o.y) * invd.y;
  double tyMax = ((neg[1] ? b.mn.y : b.mx.y) - r.o.y) * invd.y;
  tMax *= g;
  tyMax *= g;
  if (tMin > tyMax || tyMin > tMax) return false;
  if (tyMin > tMin) tMin = tyMin;
  if (tyMax < tMax) tMax = tyMax;
  double tzMin = ((neg[2] ? b.mx.z : b.mn.z) - r.o.z) * invd.z;
  double tzMax = ((neg[2] ? b.mn.z : b.mx.z) - r.o.z) * invd.z;
  tzMax *= g;
  if (tMin > tzMax || tzMin > tMax) return false;
  if (tzMin > tMin) tMin = tzMin;
  if (tzMax < tMax) tMax = tzMax;
  return tMin < r.tmax && tMax > 0;
}
// bounds.go:105-112
static inline void b3_bounding_sphere(const B3& b, V3* c, double* rad) {
  *c = divs(add(b.mn, b.mx), 2.0);
  *rad = 0;
  bool inside = c->x >= b.mn.x && c->x <= b.mx.x && c->y >= b.mn.y && c->y <= b.mx.y && c->z >= b.mn.z && c->z <= b.mx.z;
  if (inside) *rad = dist(*c, b.mx);
}

// ---------------------------------------------------------------- surface interaction (the fields Path.Li reads)
// pkg/pbrt/interaction.go:23-30,123-148.  Per SURVEY §0.10/Q12 the fields below all end up in world space
// after Shape.Intersect + TransformedPrimitive.Intersect; top-level dpdu.. (primitive space) are never read.
struct Hit {
  V3 p, perr, n, wo;  // interaction.{Point,PointError,Normal,Wo}
  V3 ns, sdpdu;       // Shading.{Normal,dpdu}
  double u = 0, v = 0;
  double time = 0;
  int prim = -1;
};

// transform.go:302-334 restricted to those fields, with the aliasing semantics spelled out:
//   Point/PointError/Normal/Wo go through the shared *interaction; Shading.* through the shared *Shading;
//   Shading.Normal is the UN-normalised transform of the previous Shading.Normal, then FaceForward'ed.
static inline void xf_hit(const Xf& t, Hit& h) {
  V3 perr;
  V3 p = xf_point(t, h.p, h.perr, &perr);
  h.p = p;
  h.perr = perr;
  h.n = normalized(xf_normal(t, h.n));
  h.wo = normalized(xf_vector(t, h.wo));
  h.ns = xf_normal(t, h.ns);
  h.sdpdu = xf_vector(t, h.sdpdu);
  h.ns = faceforward(h.ns, h.n);
}

// ---------------------------------------------------------------- shapes
struct Sphere {
  Xf o2w, w2o;
  bool reverse;
  double radius, zMin, zMax, thetaMin, thetaMax, phiMax;
};
static inline Sphere make_sphere(const gopbrt_sphere& s, const Xf& o2w) {  // sphere.go:19-36
  Sphere r;
  r.o2w = o2w;
  r.w2o = xf_inverse(o2w);
  r.reverse = s.reverse_orientation != 0;
  r.radius = s.radius;
  r.zMin = gm::Clamp(gm::Min(s.z_min, s.z_max), -s.radius, s.radius);
  r.zMax = gm::Clamp(gm::Max(s.z_min, s.z_max), -s.radius, s.radius);
  r.thetaMin = gm::Acos(gm::Clamp(gm::Min(s.z_min, s.z_max) / s.radius, -1, 1));
  r.thetaMax = gm::Acos(gm::Clamp(gm::Max(s.z_min, s.z_max) / s.radius, -1, 1));
  r.phiMax = gm::Pi / 180.0 * gm::Clamp(s.phi_max_deg, 0, 360);
  return r;
}
static inline B3 sphere_object_bound(const Sphere& s) {  // sphere.go:46-51
  B3 b; b.mn = {-s.radius, -s.radius, s.zMin}; b.mx = {s.radius, s.radius, s.zMax}; b.valid = true; return b;
}
static inline double sphere_area(const Sphere& s) { return s.phiMax * s.radius * (s.zMax - s.zMin); }  // sphere.go:42-44

// sphere.go:64-188 (full==true) and :190-268 (full==false: IntersectP)
static inline bool sphere_intersect(const Sphere& s, const Ray& r, double* tHit, Hit* hit) {
  V3 oe, de;
  Ray ray = xf_ray(s.w2o, r, &oe, &de);
  EF ox = ef_new(ray.o.x, oe.x), oy = ef_new(ray.o.y, oe.y), oz = ef_new(ray.o.z, oe.z);
  EF dx = ef_new(ray.d.x, de.x), dy = ef_new(ray.d.y, de.y), dz = ef_new(ray.d.z, de.z);
  EF a = ef_add(ef_add(ef_mul(dx, dx), ef_mul(dy, dy)), ef_mul(dz, dz));
  EF b = ef_muls(ef_add(ef_add(ef_mul(dx, ox), ef_mul(dy, oy)), ef_mul(dz, oz)), 2.0);
  EF c = ef_sub(ef_add(ef_add(ef_mul(ox, ox), ef_mul(oy, oy)), ef_mul(oz, oz)), ef_muls(ef_new(s.radius, 0), s.radius));
  EF t0, t1;
  if (!ef_quadratic(a, b, c, &t0, &t1)) return false;
  if (t0.hi > ray.tmax || t1.lo <= 0) return false;
  EF tShape = t0;
  bool isT1 = false;
  if (tShape.lo <= 0) {
    tShape = t1;
    isT1 = true;
    if (tShape.hi > ray.tmax) return false;
  }
  V3 pHit = ray_at(ray, tShape.v);
  pHit = muls(pHit, s.radius / dist(pHit, V3{}));
  if (pHit.x == 0.0 && pHit.y == 0.0) pHit.x = 1e-5 * s.radius;
  double phi = gm::Atan2(pHit.y, pHit.x);
  if (phi < 0.0) phi += 2 * gm::Pi;
  if ((s.zMin > -s.radius && pHit.z < s.zMin) || (s.zMax < s.radius && pHit.z > s.zMax) || phi > s.phiMax) {
    if (isT1) return false;  // pointer compare tShapeHit == t1
    if (t1.hi > ray.tmax) return false;
    tShape = t1;
    pHit = ray_at(ray, tShape.v);
    pHit = muls(pHit, s.radius / dist(pHit, V3{}));
    if (pHit.x == 0.0 && pHit.y == 0.0) pHit.x = 1e-5 * s.radius;
    double phi2 = gm::Atan2(pHit.y, pHit.x);  // `phi :=` shadows: the outer phi keeps its first value (sphere.go:127)
    if (phi2 < 0.0) phi2 += 2 * gm::Pi;
    if ((s.zMin > -s.radius && pHit.z < s.zMin) || (s.zMax < s.radius && pHit.z > s.zMax) || phi2 > s.phiMax) return false;
  }
  *tHit = tShape.v;
  if (!hit) return true;
  double u = phi / s.phiMax;
  double theta = gm::Acos(gm::Clamp(pHit.z / s.radius, -1, 1));
  double v = (theta - s.thetaMin) / (s.thetaMax - s.thetaMin);
  double zRadius = std::sqrt(pHit.x * pHit.x + pHit.y * pHit.y);
  double invZ = 1.0 / zRadius;
  double cosPhi = pHit.x * invZ, sinPhi = pHit.y * invZ;
  V3 dpdu{-s.phiMax * pHit.y, s.phiMax * pHit.x, 0};
  V3 dpdv = muls(V3{pHit.z * cosPhi, pHit.z * sinPhi, -s.radius * gm::Sin(theta)}, s.thetaMax - s.thetaMin);
  V3 pError = muls(vabs(pHit), gm::Gamma(5));
  // NewSurfaceInteractionWith (interaction.go:176-207)
  V3 n = normalized(cross(dpdu, dpdv));
  if (s.reverse) n = muls(n, -1);  // ReverseOrientation() != TransformSwapsHandedness() (always false)
  hit->p = pHit; hit->perr = pError; hit->n = n; hit->wo = muls(ray.d, -1);
  hit->ns = n; hit->sdpdu = dpdu; hit->u = u; hit->v = v; hit->time = ray.time;
  xf_hit(s.o2w, *hit);  // sphere.go:185
  return true;
}

struct Disk {
  Xf o2w, w2o;
  bool reverse;
  double height, radius, innerRadius, phiMax;
};
static inline Disk make_disk(const gopbrt_disk& d, const Xf& o2w) {  // disk.go:22-35
  Disk r;
  r.o2w = o2w; r.w2o = xf_inverse(o2w); r.reverse = d.reverse_orientation != 0;
  r.height = d.height; r.radius = d.radius; r.innerRadius = d.inner_radius;
  r.phiMax = gm::Pi / 180.0 * gm::Clamp(d.phi_max_deg, 0, 360);
  return r;
}
static inline B3 disk_object_bound(const Disk& d) {  // disk.go:41-54
  B3 b; b.mn = {-d.radius, -d.radius, d.height}; b.mx = {d.radius, d.radius, d.height}; b.valid = true; return b;
}
static inline double disk_area(const Disk& d) { return d.phiMax * 0.5 * (d.radius * d.radius - d.innerRadius * d.innerRadius); }  // disk.go:183-185
// disk.go:64-126 / :127-159
static inline bool disk_intersect(const Disk& d, const Ray& r, double* tHit, Hit* hit) {
  Ray ray = xf_ray(d.w2o, r, nullptr, nullptr);
  if (ray.d.z == 0) return false;
  double t = (d.height - ray.o.z) / ray.d.z;
  if (t <= 0 || t >= ray.tmax) return false;
  V3 pHit = ray_at(ray, t);
  double dist2v = pHit.x * pHit.x + pHit.y * pHit.y;
  if (dist2v > d.radius * d.radius || dist2v < d.innerRadius * d.innerRadius) return false;
  double phi = gm::Atan2(pHit.y, pHit.x);
  if (phi < 0) phi += 2 * gm::Pi;
  if (phi > d.phiMax) return false;
  *tHit = t;
  if (!hit) return true;
  double u = phi / d.phiMax;
  double rHit = std::sqrt(dist2v);
  double oneMinusV = (rHit - d.innerRadius) / (d.radius - d.innerRadius);
  double v = 1 - oneMinusV;
  V3 dpdu{-d.phiMax * pHit.y, d.phiMax * pHit.x, 0};
  V3 dpdv = muls(V3{pHit.x, pHit.y, 0}, (d.radius - d.innerRadius) / rHit);  // normal comes out -z
  pHit.z = d.height;
  V3 n = normalized(cross(dpdu, dpdv));
  if (d.reverse) n = muls(n, -1);
  hit->p = pHit; hit->perr = V3{}; hit->n = n; hit->wo = muls(ray.d, -1);
  hit->ns = n; hit->sdpdu = dpdu; hit->u = u; hit->v = v; hit->time = ray.time;
  xf_hit(d.o2w, *hit);
  return true;
}

// Triangle — NOT in the reference (parity unpinned).  Defined here: world-space vertices, the pbrt-v3
// watertight test (translate, permute, shear; edge functions) carried out in float64 with this repo's
// conventions: reject `t <= 0 || t >= tMax` like disk.go:75, hit point = barycentric interpolation,
// pError = Gamma(7)*sum|b_i p_i| with this repo's (denormal) Gamma, default uv (0,0),(1,0),(1,1),
// n = normalize(cross(dp02, dp12)) flipped by reverseOrientation, shading normal = n, Shading.dpdu = dpdu
// (or a CoordinateSystem(n) tangent when the uv system is degenerate).
struct Tri { V3 p0, p1, p2; bool reverse; };
static inline int max_dim(V3 v) { return (v.x > v.y) ? ((v.x > v.z) ? 0 : 2) : ((v.y > v.z) ? 1 : 2); }
static inline V3 permute(V3 v, int x, int y, int z) { return {v[x], v[y], v[z]}; }
static inline bool tri_intersect(const Tri& tr, const Ray& ray, double* tHit, Hit* hit) {
  V3 p0t = sub(tr.p0, ray.o), p1t = sub(tr.p1, ray.o), p2t = sub(tr.p2, ray.o);
  int kz = max_dim(vabs(ray.d));
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  V3 d = permute(ray.d, kx, ky, kz);
  p0t = permute(p0t, kx, ky, kz); p1t = permute(p1t, kx, ky, kz); p2t = permute(p2t, kx, ky, kz);
  double Sx = -d.x / d.z, Sy = -d.y / d.z, Sz = 1.0 / d.z;
  p0t.x += Sx * p0t.z; p0t.y += Sy * p0t.z;
  p1t.x += Sx * p1t.z; p1t.y += Sy * p1t.z;
  p2t.x += Sx * p2t.z; p2t.y += Sy * p2t.z;
  double e0 = p1t.x * p2t.y - p1t.y * p2t.x;
  double e1 = p2t.x * p0t.y - p2t.y * p0t.x;
  double e2 = p0t.x * p1t.y - p0t.y * p1t.x;
  if ((e0 < 0 || e1 < 0 || e2 < 0) && (e0 > 0 || e1 > 0 || e2 > 0)) return false;
  double det = e0 + e1 + e2;
  if (det == 0) return false;
  p0t.z *= Sz; p1t.z *= Sz; p2t.z *= Sz;
  double tScaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
  if (det < 0 && (tScaled >= 0 || tScaled < ray.tmax * det)) return false;
  if (det > 0 && (tScaled <= 0 || tScaled > ray.tmax * det)) return false;
  double invDet = 1 / det;
  double b0 = e0 * invDet, b1 = e1 * invDet, b2 = e2 * invDet;
  double t = tScaled * invDet;
  if (t <= 0 || t >= ray.tmax) return false;
  *tHit = t;
  if (!hit) return true;
  V3 dp02 = sub(tr.p0, tr.p2), dp12 = sub(tr.p1, tr.p2);
  // uv = (0,0),(1,0),(1,1): duv02 = (-1,-1), duv12 = (0,-1); determinant = 1
  double du02 = -1, dv02 = -1, du12 = 0, dv12 = -1;
  double determinant = du02 * dv12 - dv02 * du12;
  V3 dpdu, dpdv;
  bool degenerate = std::fabs(determinant) < 1e-8;
  if (!degenerate) {
    double invdet = 1 / determinant;
    dpdu = muls(sub(muls(dp02, dv12), muls(dp12, dv02)), invdet);
    dpdv = muls(add(muls(dp02, -du12), muls(dp12, du02)), invdet);
  }
  V3 n = normalized(cross(dp02, dp12));
  if (degenerate || len2(cross(dpdu, dpdv)) == 0) {
    V3 tmp;
    coordinate_system(n, &dpdu, &tmp);
  }
  V3 pAbs = add(add(vabs(muls(tr.p0, b0)), vabs(muls(tr.p1, b1))), vabs(muls(tr.p2, b2)));
  V3 pError = muls(pAbs, gm::Gamma(7));
  V3 pHit = add(add(muls(tr.p0, b0), muls(tr.p1, b1)), muls(tr.p2, b2));
  if (tr.reverse) n = muls(n, -1);
  hit->p = pHit; hit->perr = pError; hit->n = n; hit->wo = muls(ray.d, -1);
  hit->ns = n; hit->sdpdu = dpdu;
  hit->u = b0 * 0 + b1 * 1 + b2 * 1; hit->v = b0 * 0 + b1 * 0 + b2 * 1;
  hit->time = ray.time;
  return true;
}

// ---------------------------------------------------------------- scene + primitives
struct Prim { int kind, index, material, p2w; };

struct Scene {
  std::vector<Xf> xf;
  std::vector<Sphere> spheres;
  std::vector<Disk> disks;
  std::vector<V3> verts;
  std::vector<gopbrt_triangle> tris;
  std::vector<Prim> prims;
  std::vector<gopbrt_material> materials;
  std::vector<gopbrt_texture> textures;
  std::vector<gopbrt_light> lights;
  std::vector<B3> prim_bounds;  // Primitive.WorldBound()
  B3 world;
  // light distribution (lightdistribution.go:25-42, sampling.go:11-40)
  std::vector<double> light_cdf;
  double light_func_int = 0;
  // Distant.Preprocess (distant.go:36-38)
  V3 world_center;
  double world_radius = 0;

  // accelerator
  // 0 = reference RecursiveBuild(SplitSAH) + [64] stack (config 1 only: the reference tree is O(N) deep);
  // 1 = oracle's own median tree, 2 = brute force — both apply the PARITY SPEC of SURVEY §8a: a primitive is a
  //     candidate iff its OWN fp64 world bound passes Bounds3.IntersectP with the running tMax (own_bound_test);
  // 3 / 4 = as 1 / 2 without the own-bound test (analysis only).
  int accel_mode = 1;
  bool own_bound_test = true;
  struct Node { B3 b; uint64_t primOffset = 0, second = 0, nPrims = 0; uint8_t axis = 0; };
  std::vector<Node> nodes;
  std::vector<int> ordered;  // orderedPrims → original primitive index
  bool brute() const { return accel_mode == 2 || accel_mode == 4; }
  int max_prims = 2;
};

static inline Tri scene_tri(const Scene& sc, int i) {
  const gopbrt_triangle& t = sc.tris[i];
  return Tri{sc.verts[t.v[0]], sc.verts[t.v[1]], sc.verts[t.v[2]], t.reverse_orientation != 0};
}

// GeometricPrimitive.Intersect (primitive.go:46-61) under an optional TransformedPrimitive (primitive.go:94-109).
// On a hit sets r.tmax = tHit exactly as the reference does.
static inline bool prim_intersect(const Scene& sc, int pi, Ray& r, Hit* hit) {
  const Prim& p = sc.prims[pi];
  Ray ray = r;
  if (p.p2w >= 0) ray = xf_ray(xf_inverse(sc.xf[p.p2w]), r, nullptr, nullptr);
  double t;
  bool ok = false;
  switch (p.kind) {
    case GOPBRT_SHAPE_SPHERE: ok = sphere_intersect(sc.spheres[p.index], ray, &t, hit); break;
    case GOPBRT_SHAPE_DISK: ok = disk_intersect(sc.disks[p.index], ray, &t, hit); break;
    case GOPBRT_SHAPE_TRIANGLE: ok = tri_intersect(scene_tri(sc, p.index), ray, &t, hit); break;
  }
  if (!ok) return false;
  r.tmax = t;  // primitive.go:51 then :102
  if (hit) {
    hit->prim = pi;
    if (p.p2w >= 0 && !xf_is_identity(sc.xf[p.p2w])) xf_hit(sc.xf[p.p2w], *hit);  // primitive.go:104-106, SURVEY Q12
  }
  return true;
}
static inline bool prim_intersect_p(const Scene& sc, int pi, const Ray& r) {  // primitive.go:42-44,111-115
  Ray tmp = r;
  return prim_intersect(sc, pi, tmp, nullptr);
}

static inline B3 prim_world_bound(const Scene& sc, const Prim& p) {
  B3 b;
  switch (p.kind) {
    case GOPBRT_SHAPE_SPHERE: b = xf_bounds(sc.spheres[p.index].o2w, sphere_object_bound(sc.spheres[p.index])); break;  // sphere.go:53-55
    case GOPBRT_SHAPE_DISK: b = xf_bounds(sc.disks[p.index].o2w, disk_object_bound(sc.disks[p.index])); break;          // disk.go:55-57
    case GOPBRT_SHAPE_TRIANGLE: {
      Tri t = scene_tri(sc, p.index);
      b3_union_point(b, t.p0); b3_union_point(b, t.p1); b3_union_point(b, t.p2);
    } break;
  }
  if (p.p2w >= 0) b = xf_bounds(sc.xf[p.p2w], b);  // primitive.go:127-129 → transform.go:584-590
  return b;
}

// ---- reference BVH build: bvh.go:223-411 with SplitSAH, bugs included ----
struct BuildInfo { int prim; B3 b; V3 c; };
struct BuildNode { B3 b; BuildNode* ch[2] = {nullptr, nullptr}; uint8_t axis = 0; int64_t first = 0, n = 0; };

template <class F>
static int64_t partition_at(std::vector<BuildInfo>& in, int64_t start, int64_t end, int64_t pivot, F f) {  // bvh.go:163-175
  BuildInfo pv = in[pivot];
  std::swap(in[pivot], in[end]);
  for (int64_t i = start; i < end; i++)
    if (f(in[i], pv)) { std::swap(in[start], in[i]); start++; }
  std::swap(in[end], in[start]);
  return start;
}

static BuildNode* ref_recursive_build(Scene& sc, std::vector<BuildInfo>& info, int64_t start, int64_t end, int64_t* total,
                                      std::vector<BuildNode*>& pool) {
  BuildNode* node = new BuildNode();
  pool.push_back(node);
  (*total)++;
  B3 bounds;
  for (int64_t i = start; i < end; i++) b3_union(bounds, info[i].b);
  int64_t n = end - start;
  auto make_leaf = [&]() {
    node->first = (int64_t)sc.ordered.size();
    for (int64_t i = start; i < end; i++) sc.ordered.push_back(info[i].prim);
    node->n = n;
    node->b = bounds;
    return node;
  };
  if (n == 1) return make_leaf();
  B3 cb;
  for (int64_t i = start; i < end; i++) b3_union_point(cb, info[i].c);
  int dim = b3_max_extent(cb);
  int64_t mid = (start + end) / 2;
  if (cb.mx[dim] == cb.mn[dim]) return make_leaf();
  // SplitSAH (bvh.go:341-408)
  if (n <= 2) {
    partition_at(info, start, end - 1, mid, [&](const BuildInfo& a, const BuildInfo& b) { return a.c[dim] < b.c[dim]; });
  } else {
    const int nBuckets = 12;
    struct Bucket { int count = 0; B3 b; } buckets[nBuckets];
    auto bucket_of = [&](const BuildInfo& a) {
      int b = nBuckets * (int)(b3_offset(cb, a.c)[dim]);  // int() BEFORE the multiply: only buckets 0 and 12→11
      if (b == nBuckets) b = nBuckets - 1;
      return b;
    };
    for (int64_t i = start; i < end; i++) {
      int b = bucket_of(info[i]);
      buckets[b].count++;
      b3_union(buckets[b].b, info[i].b);
    }
    double cost[nBuckets - 1];
    for (int i = 0; i < nBuckets - 1; i++) {
      B3 b0, b1;
      int c0 = 0, c1 = 0;
      for (int j = 0; j <= i; j++) { b3_union(b0, buckets[j].b); c0 += buckets[j].count; }
      for (int j = i + 1; j < nBuckets; j++) { b3_union(b1, buckets[j].b); c1 += buckets[j].count; }
      // SurfaceArea of a nil-bounds Bounds3 would nil-deref in Go only if Diagonal() is called on nil Min/Max;
      // b0 always holds bucket 0 (the min-centroid primitive), b1 always bucket 11 (the max) when n>=2 and extent>0.
      double a0 = b0.valid ? b3_area(b0) : 0.0, a1 = b1.valid ? b3_area(b1) : 0.0;
      cost[i] = 1.0 + ((double)c0 * a0 + (double)c1 * a1) / b3_area(bounds);
    }
    double minCost = cost[0];
    int minB = 0;
    for (int i = 1; i < nBuckets - 1; i++)
      if (cost[i] < minCost) { minCost = cost[i]; minB = i; }
    double leafCost = (double)n;
    if (n > (int64_t)sc.max_prims || minCost < leafCost) {
      mid = partition_at(info, start, end - 1, end - 1, [&](const BuildInfo& a, const BuildInfo&) { return bucket_of(a) <= minB; });
    } else {
      return make_leaf();
    }
  }
  // mid == start or mid == end makes the reference recurse on an empty range and nil-deref (bvh.go:278-296);
  // the oracle stops with a leaf instead (never reached by the scenes under test).
  if (mid <= start || mid >= end) return make_leaf();
  BuildNode* c0 = ref_recursive_build(sc, info, start, mid, total, pool);
  BuildNode* c1 = ref_recursive_build(sc, info, mid, end, total, pool);
  node->ch[0] = c0; node->ch[1] = c1;
  node->b = c0->b;
  b3_union(node->b, c1->b);  // bvh.go:51-63
  node->axis = (uint8_t)dim;
  node->n = 0;
  return node;
}

// the oracle's own tree for large scenes (the reference tree is O(N) deep): median split on the
// largest centroid extent, leaves of <= max_prims.  Same per-node/per-primitive arithmetic as the reference.
static BuildNode* sane_build(Scene& sc, std::vector<BuildInfo>& info, int64_t start, int64_t end, int64_t* total,
                             std::vector<BuildNode*>& pool) {
  BuildNode* node = new BuildNode();
  pool.push_back(node);
  (*total)++;
  B3 bounds;
  for (int64_t i = start; i < end; i++) b3_union(bounds, info[i].b);
  int64_t n = end - start;
  B3 cb;
  for (int64_t i = start; i < end; i++) b3_union_point(cb, info[i].c);
  int dim = b3_max_extent(cb);
  if (n <= sc.max_prims || cb.mx[dim] == cb.mn[dim]) {
    node->first = (int64_t)sc.ordered.size();
    for (int64_t i = start; i < end; i++) sc.ordered.push_back(info[i].prim);
    node->n = n;
    node->b = bounds;
    return node;
  }
  int64_t mid = (start + end) / 2;
  std::nth_element(info.begin() + start, info.begin() + mid, info.begin() + end,
                   [dim](const BuildInfo& a, const BuildInfo& b) { return a.c[dim] < b.c[dim]; });
  BuildNode* c0 = sane_build(sc, info, start, mid, total, pool);
  BuildNode* c1 = sane_build(sc, info, mid, end, total, pool);
  node->ch[0] = c0; node->ch[1] = c1;
  node->b = c0->b;
  b3_union(node->b, c1->b);
  node->axis = (uint8_t)dim;
  return node;
}

static uint64_t flatten(Scene& sc, BuildNode* n, uint64_t* offset) {  // bvh.go:632-651
  uint64_t my = (*offset)++;
  sc.nodes[my].b = n->b;
  if (n->n > 0) {
    sc.nodes[my].primOffset = (uint64_t)n->first;
    sc.nodes[my].nPrims = (uint64_t)n->n;
  } else {
    sc.nodes[my].axis = n->axis;
    sc.nodes[my].nPrims = 0;
    flatten(sc, n->ch[0], offset);
    sc.nodes[my].second = flatten(sc, n->ch[1], offset);
  }
  return my;
}

static inline void build_accel(Scene& sc) {
  sc.nodes.clear();
  sc.ordered.clear();
  if (sc.prims.empty() || sc.brute()) return;
  std::vector<BuildInfo> info(sc.prims.size());
  for (size_t i = 0; i < sc.prims.size(); i++) {
    info[i].prim = (int)i;
    info[i].b = sc.prim_bounds[i];
    info[i].c = add(muls(sc.prim_bounds[i].mn, 0.5), muls(sc.prim_bounds[i].mx, 0.5));  // bvh.go:33
  }
  int64_t total = 0;
  std::vector<BuildNode*> pool;
  BuildNode* root = sc.accel_mode == 0 ? ref_recursive_build(sc, info, 0, (int64_t)info.size(), &total, pool)
                                       : sane_build(sc, info, 0, (int64_t)info.size(), &total, pool);
  sc.nodes.resize(total);
  uint64_t off = 0;
  flatten(sc, root, &off);
  for (auto* p : pool) delete p;
}

struct TravStats { uint64_t nodes = 0, prims = 0; };

// Closest hit of accel modes 1 (own tree) and 2 (brute force): the ORDER-INDEPENDENT statement of the parity spec
//  — over all primitives whose own float64 bound and shape test pass with the ray's ORIGINAL tMax,
// the minimum tHit; bit-identical tHit goes to the lower primitive index.  It equals the reference's running-tMax rule
// (primitive.go:51, kept verbatim in accel mode 0 below) except where two candidates' distances differ by less than the
// rounding of the slab / shape arithmetic — coplanar faces, grazing hits within the EFloat bound — where the reference's
// answer depends on its own tree's visit order.  The running best distance only culls what cannot win, with a 2^-20 slack.
static inline bool scene_intersect_canonical(const Scene& sc, Ray& ray, Hit* hit, TravStats* ts, V3 invd, const int neg[3]) {
  const double t_orig = ray.tmax, slack = 1.0 + 9.5367431640625e-07;
  double t_best = t_orig;
  int best = -1;
  Hit best_hit;
  auto cull_ray = [&]() {
    Ray rc = ray;
    double c = t_best * slack;
    rc.tmax = (best >= 0 && c < t_orig) ? c : t_orig;
    return rc;
  };
  auto candidate = [&](int pi) {
    if (!b3_intersect_p(sc.prim_bounds[pi], cull_ray(), invd, neg)) return;
    if (ts) ts->prims++;
    Ray r2 = ray;  // original tMax
    Hit h;
    if (prim_intersect(sc, pi, r2, &h)) {
      if (r2.tmax < t_best || (r2.tmax == t_best && pi < best)) { t_best = r2.tmax; best = pi; best_hit = h; }
    }
  };
  if (sc.brute()) {
    for (size_t i = 0; i < sc.prims.size(); i++) candidate((int)i);
  } else if (!sc.nodes.empty()) {
    uint64_t toVisit = 0, cur = 0;
    uint64_t stack[256];
    for (;;) {
      const Scene::Node& node = sc.nodes[cur];
      if (ts) ts->nodes++;
      if (b3_intersect_p(node.b, cull_ray(), invd, neg)) {
        if (node.nPrims > 0) {
          for (uint64_t i = 0; i < node.nPrims; i++) candidate(sc.ordered[node.primOffset + i]);
          if (toVisit == 0) break;
          cur = stack[--toVisit];
        } else {
          if (toVisit >= 256) { g_counters.stack_overflows.fetch_add(1); break; }
          if (neg[node.axis]) { stack[toVisit++] = cur + 1; cur = node.second; }
          else { stack[toVisit++] = node.second; cur = cur + 1; }
        }
      } else {
        if (toVisit == 0) break;
        cur = stack[--toVisit];
      }
    }
  }
  if (best < 0) return false;
  ray.tmax = t_best;
  if (hit) *hit = best_hit;
  return true;
}

// BVH.Intersect (bvh.go:659-712).  accel_mode 0 is the reference itself: running tMax, its own visit order and its fixed
// [64] stack (an overflow is where Go would panic with index-out-of-range — counted, traversal abandoned).
static inline bool scene_intersect(const Scene& sc, Ray& ray, Hit* hit, TravStats* ts = nullptr) {
  V3 invd{1 / ray.d.x, 1 / ray.d.y, 1 / ray.d.z};
  int neg[3] = {invd.x < 0, invd.y < 0, invd.z < 0};
  if (sc.own_bound_test) return scene_intersect_canonical(sc, ray, hit, ts, invd, neg);
  if (sc.brute()) {  // accelerator.Simple ordering is by distance; here: plain loop with running tMax
    bool any = false;
    for (size_t i = 0; i < sc.prims.size(); i++) {
      if (sc.own_bound_test && !b3_intersect_p(sc.prim_bounds[i], ray, invd, neg)) continue;
      if (ts) ts->prims++;
      if (prim_intersect(sc, (int)i, ray, hit)) any = true;
    }
    return any;
  }
  if (sc.nodes.empty()) return false;
  bool any = false;
  uint64_t toVisit = 0, cur = 0;
  const size_t cap = sc.accel_mode == 0 ? 64 : 256;
  uint64_t stack[256];
  for (;;) {
    const Scene::Node& node = sc.nodes[cur];
    if (ts) ts->nodes++;
    if (b3_intersect_p(node.b, ray, invd, neg)) {
      if (node.nPrims > 0) {
        for (uint64_t i = 0; i < node.nPrims; i++) {
          int pi = sc.ordered[node.primOffset + i];
          if (sc.own_bound_test && !b3_intersect_p(sc.prim_bounds[pi], ray, invd, neg)) continue;
          if (ts) ts->prims++;
          if (prim_intersect(sc, pi, ray, hit)) any = true;
        }
        if (toVisit == 0) break;
        cur = stack[--toVisit];
      } else {
        if (toVisit >= cap) { g_counters.stack_overflows.fetch_add(1); break; }
        if (neg[node.axis]) { stack[toVisit++] = cur + 1; cur = node.second; }
        else { stack[toVisit++] = node.second; cur = cur + 1; }
      }
    } else {
      if (toVisit == 0) break;
      cur = stack[--toVisit];
    }
  }
  return any;
}

// BVH.IntersectP (bvh.go:713-765)
static inline bool scene_intersect_p(const Scene& sc, const Ray& ray, TravStats* ts = nullptr) {
  V3 invd{1 / ray.d.x, 1 / ray.d.y, 1 / ray.d.z};
  int neg[3] = {invd.x < 0, invd.y < 0, invd.z < 0};
  if (sc.brute()) {
    for (size_t i = 0; i < sc.prims.size(); i++) {
      if (sc.own_bound_test && !b3_intersect_p(sc.prim_bounds[i], ray, invd, neg)) continue;
      if (ts) ts->prims++;
      if (prim_intersect_p(sc, (int)i, ray)) return true;
    }
    return false;
  }
  if (sc.nodes.empty()) return false;
  uint64_t toVisit = 0, cur = 0;
  const size_t cap = sc.accel_mode == 0 ? 64 : 256;
  uint64_t stack[256];
  for (;;) {
    const Scene::Node& node = sc.nodes[cur];
    if (ts) ts->nodes++;
    if (b3_intersect_p(node.b, ray, invd, neg)) {
      if (node.nPrims > 0) {
        for (uint64_t i = 0; i < node.nPrims; i++) {
          int pi = sc.ordered[node.primOffset + i];
          if (sc.own_bound_test && !b3_intersect_p(sc.prim_bounds[pi], ray, invd, neg)) continue;
          if (ts) ts->prims++;
          if (prim_intersect_p(sc, pi, ray)) return true;
        }
        if (toVisit == 0) return false;
        cur = stack[--toVisit];
      } else {
        if (toVisit >= cap) { g_counters.stack_overflows.fetch_add(1); return false; }
        if (neg[node.axis]) { stack[toVisit++] = cur + 1; cur = node.second; }
        else { stack[toVisit++] = node.second; cur = cur + 1; }
      }
    } else {
      if (toVisit == 0) return false;
      cur = stack[--toVisit];
    }
  }
}

static inline bool valid_index(int64_t i, int64_t n) { return i >= 0 && i < n; }

// NewUniformLightDistribution + NewDistribution1D (lightdistribution.go:25-34, sampling.go:11-40): func[i] = 1
static inline void uniform_light_distribution(int n, std::vector<double>* cdf, double* func_int) {
  cdf->assign(n + 1, 0.0);
  for (int i = 1; i < n + 1; i++) (*cdf)[i] = (*cdf)[i - 1] + 1.0 / (double)n;
  *func_int = n ? (*cdf)[n] : 0.0;
  if (*func_int == 0.0) {
    for (int i = 1; i < n + 1; i++) (*cdf)[i] = (double)i / (double)n;
  } else {
    for (int i = 1; i < n + 1; i++) (*cdf)[i] /= *func_int;
  }
}

// accelerator.NewBVH + pbrt.NewScene (bvh.go:223-270, scene.go:16-36)
static inline Scene* scene_from_desc(const gopbrt_scene_desc* d, int accel_mode) {
  Scene* sc = new Scene();
  sc->accel_mode = accel_mode;
  sc->own_bound_test = (accel_mode == 1 || accel_mode == 2);
  sc->max_prims = d->max_prims_in_node > 0 ? std::min(255, d->max_prims_in_node) : 4;
  for (int i = 0; i < d->n_transforms; i++) sc->xf.push_back(xf_from(d->transforms[i]));
  for (int i = 0; i < d->n_spheres; i++) {
    if (!valid_index(d->spheres[i].object_to_world, d->n_transforms)) { delete sc; return nullptr; }
    sc->spheres.push_back(make_sphere(d->spheres[i], sc->xf[d->spheres[i].object_to_world]));
  }
  for (int i = 0; i < d->n_disks; i++) {
    if (!valid_index(d->disks[i].object_to_world, d->n_transforms)) { delete sc; return nullptr; }
    sc->disks.push_back(make_disk(d->disks[i], sc->xf[d->disks[i].object_to_world]));
  }
  for (int64_t i = 0; i < d->n_vertices; i++) sc->verts.push_back(V3{d->vertices[3 * i], d->vertices[3 * i + 1], d->vertices[3 * i + 2]});
  for (int64_t i = 0; i < d->n_triangles; i++) sc->tris.push_back(d->triangles[i]);
  for (int64_t i = 0; i < d->n_primitives; i++) {
    const gopbrt_primitive& p = d->primitives[i];
    sc->prims.push_back(Prim{p.shape_kind, p.shape_index, p.material, p.prim_to_world});
  }
  for (int i = 0; i < d->n_materials; i++) sc->materials.push_back(d->materials[i]);
  for (int i = 0; i < d->n_textures; i++) sc->textures.push_back(d->textures[i]);
  for (int i = 0; i < d->n_lights; i++) sc->lights.push_back(d->lights[i]);
  sc->prim_bounds.resize(sc->prims.size());
  for (size_t i = 0; i < sc->prims.size(); i++) {
    sc->prim_bounds[i] = prim_world_bound(*sc, sc->prims[i]);
    b3_union(sc->world, sc->prim_bounds[i]);
  }
  build_accel(*sc);
  if (sc->world.valid) b3_bounding_sphere(sc->world, &sc->world_center, &sc->world_radius);
  uniform_light_distribution((int)sc->lights.size(), &sc->light_cdf, &sc->light_func_int);
  return sc;
}

}  // namespace oracle
