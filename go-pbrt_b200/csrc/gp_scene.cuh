// gp_scene.cuh — device-resident scene layout and the shape routines of the hot path.
//
// HBM layout (uploaded once by gopbrt_scene_create):
//   nodes      float4[2*N]   32-byte BVH nodes: {min.xyz (f32, rounded down), a}, {max.xyz (f32, rounded up), b}
//                            interior: a = second child, b = axis;  leaf: a = first record, b = nPrims<<8 | axis
//                            (same information as LinearBVHNode, pkg/accelerator/bvh.go:80-87; first child = n+1)
//   recs       80-byte primitive records in LEAF order: {u32 kind|flags, u32 primitive id, double d[9]}
//                            sphere (translation-only transforms): d = radius, w2o translation, p2w^-1 translation
//                            triangle: d = p0, p1, p2 (world space);  anything else: looked up through prims/shape tables
//   rec_bounds double[6] per record: the primitive's own float64 world bound (spheres/disks) for the parity-spec
//                            candidate test; triangles derive it from their vertices
//   prims      int4 per primitive id: {shape kind, shape index, material, prim_to_world transform or -1}
//   xf         32 doubles per transform (m, minv) + flags;  sphere/disk/material/texture/light tables
#pragma once
#include "gp_math.cuh"

namespace gp {

enum : uint32_t {
  RK_SPHERE = 0, RK_DISK = 1, RK_TRIANGLE = 2, RK_KIND_MASK = 3,
  RF_FAST = 4,        // record carries everything the intersection needs (translation-only sphere / triangle)
  RF_HAS_P2W = 8,     // wrapped in a TransformedPrimitive
  RF_REVERSE = 16,    // reverseOrientation
  RF_FULL = 32,       // full sphere / full disk: the clip test can never fire (phiMax == 2Pi exactly, zMin == -r, zMax == r)
  RF_P2W_IDENTITY = 64,
  RF_CLASS_SHIFT = 8, RF_CLASS_MASK = 3u << 8  // shade class: bit 0 = material is not plain Lambert, bit 1 = sphere/disk hit
};
enum : int { XF_TRANSLATION = 1, XF_IDENTITY = 2 };

struct __align__(16) PrimRec {
  uint32_t flags;
  uint32_t prim;
  double d[9];
};
static_assert(sizeof(PrimRec) == 80, "PrimRec must be 80 bytes");

struct SphereDev { double radius, zMin, zMax, thetaMin, thetaMax, phiMax; int xf; int flags; };
struct DiskDev { double height, radius, innerRadius, phiMax; int xf; int flags; };
struct MaterialDev { int kind, tex_a, tex_b, pad; double sigma, eta, u_rough, v_rough; };
struct TextureDev { int kind, mapping, tex1, tex2; double rgb[3], vs[3], vt[3], ds, dt, su, sv, du, dv; };
struct LightDev { int kind, shape_kind, shape_index, two_sided; double rgb[3], v[3]; };

struct DevScene {
  const float4* nodes;
  const PrimRec* recs;
  const double* rec_bounds;
  const int4* prims;
  const double* xf;        // 32 doubles each
  const int* xf_flags;
  const SphereDev* spheres;
  const DiskDev* disks;
  const MaterialDev* materials;
  const TextureDev* textures;
  const LightDev* lights;
  const double* light_cdf;
  int n_lights;
  double light_func_int;
  double world_radius;     // Distant.Preprocess (distant.go:36-38)
  int n_nodes;
  // flat aggregate of small scenes (k_trace_flat): per primitive two float4 {min.xyz f32 rounded down, record index},
  // {max.xyz f32 rounded up, record flags}; triangles first (bit k of flat_tri_mask set), then spheres / disks
  const float4* flat;   // n_flat entries, then n_flat_groups box groups {box, mask of the entries that share it}
  int n_flat, n_flat_groups;
  unsigned long long flat_tri_mask;
};

// load a transform's Matrix (inv=false) or MatrixInverse (inv=true) into registers.  Translation-only transforms are
// synthesised from their three translation entries with literal 1.0 / 0.0 so the arithmetic below is the very same
// sequence of IEEE operations as the general case (0*y terms are kept: they matter for NaN/Inf and signed zeros).
GP_D M4 load_m4(const DevScene& sc, int xf, bool inv) {
  const double* b = sc.xf + (size_t)xf * 32 + (inv ? 16 : 0);
  M4 r;
  if (sc.xf_flags[xf] & XF_TRANSLATION) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) r.m[i][j] = (i == j) ? 1.0 : 0.0;
    r.m[0][3] = b[3]; r.m[1][3] = b[7]; r.m[2][3] = b[11];
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) r.m[i][j] = b[i * 4 + j];
  }
  return r;
}
// plain 16-load variant (no translation special case): same values, half the code; used by the shade stage
GP_D M4 load_m4_plain(const DevScene& sc, int xf, bool inv) {
  const double2* b = (const double2*)(sc.xf + (size_t)xf * 32 + (inv ? 16 : 0));
  M4 r;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double2 lo = __ldg(b + 2 * i), hi = __ldg(b + 2 * i + 1);
    r.m[i][0] = lo.x; r.m[i][1] = lo.y; r.m[i][2] = hi.x; r.m[i][3] = hi.y;
  }
  return r;
}
GP_HD M4 translation_m4(double x, double y, double z) {
  M4 r;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) r.m[i][j] = (i == j) ? 1.0 : 0.0;
  r.m[0][3] = x; r.m[1][3] = y; r.m[2][3] = z;
  return r;
}

// ---- Sphere.Intersect / IntersectP root finding (sphere.go:64-95,190-218) on the OBJECT-space ray ----
// returns false on miss; on success *t0/*t1 are the interval roots and *tsel the selected root value, *sel1 whether
// the selected root is t1 (the reference compares pointers).
GP_HD bool sphere_roots(const Ray& ray, V3 oe, V3 de, double radius, EF* t0, EF* t1, int& bad) {
  // Early miss.  efloat.Quadratic returns false as soon as the PLAIN discriminant b.v*b.v - 4*a.v*c.v is negative
  // (efloat/math.go:36-40), and the .v components of a, b, c are ordinary float64 arithmetic in the order written below
  // (efloat.go Add/Sub/Mul keep .v = plain op), so a miss is decided bit-exactly without any interval bound.  It is
  // taken only when every operand is far from overflow, which rules out every efloat.Check panic on the skipped path.
  {
    double ox = ray.o.x, oy = ray.o.y, oz = ray.o.z, dx = ray.d.x, dy = ray.d.y, dz = ray.d.z;
    double m = fmax(fmax(fmax(fabs(ox), fabs(oy)), fmax(fabs(oz), fabs(dx))), fmax(fmax(fabs(dy), fabs(dz)), fabs(radius)));
    double me = fmax(fmax(fmax(fabs(oe.x), fabs(oe.y)), fmax(fabs(oe.z), fabs(de.x))), fmax(fabs(de.y), fabs(de.z)));
    if (m < 1e100 && me < 1e100) {
      double av = ((dx * dx) + (dy * dy)) + (dz * dz);
      double bv = (((dx * ox) + (dy * oy)) + (dz * oz)) * 2.0;
      double cv = (((ox * ox) + (oy * oy)) + (oz * oz)) - (radius * radius);
      if (bv * bv - 4. * av * cv < 0) return false;
    }
  }
  EF ox = ef_new(ray.o.x, oe.x, bad), oy = ef_new(ray.o.y, oe.y, bad), oz = ef_new(ray.o.z, oe.z, bad);
  EF dx = ef_new(ray.d.x, de.x, bad), dy = ef_new(ray.d.y, de.y, bad), dz = ef_new(ray.d.z, de.z, bad);
  EF a = ef_add(ef_add(ef_mul(dx, dx, bad), ef_mul(dy, dy, bad), bad), ef_mul(dz, dz, bad), bad);
  EF b = ef_muls(ef_add(ef_add(ef_mul(dx, ox, bad), ef_mul(dy, oy, bad), bad), ef_mul(dz, oz, bad), bad), 2.0, bad);
  EF c = ef_sub(ef_add(ef_add(ef_mul(ox, ox, bad), ef_mul(oy, oy, bad), bad), ef_mul(oz, oz, bad), bad),
                ef_muls(ef_new(radius, 0, bad), radius, bad), bad);
  return ef_quadratic(a, b, c, t0, t1, bad);
}
// ---- full spheres without interval arithmetic, whenever the interval bounds cannot change a decision ----
// Sphere.Intersect reads the EFloat bounds of its two roots in four comparisons only (sphere.go:83-93: t0.hi > tMax, t1.lo <= 0,
// t0.lo <= 0, t1.hi > tMax); the hit distance itself is the plain value t.v, and every .v is ordinary float64 arithmetic in the
// order written here (efloat.go keeps .v = plain op).  So the ~1000-instruction interval evaluation is needed only when a root
// lies within its own error bound of 0 or tMax — a ray leaving the sphere it starts on, a grazing hit.  Everywhere else the
// comparisons are decided by v and a RIGOROUS UPPER BOUND W of the interval's half-width:
//   every EFloat operation returns the hull of the exact operation on its input intervals, each end rounded to nearest and
//   stepped one ulp outward: at most 4u|x| per end (u = 2^-53).  The inputs are (v, err) with err of the order 1e-323 |v|
//   (the reference's MachineEpsilon is the smallest denormal), i.e. one ulp each side.  Propagating:
//     squares and products            W <= 12u |x y|
//     a = dx^2+dy^2+dz^2              W(a) <= 24u a
//     b = 2 (dx ox + dy oy + dz oz)   W(b) <= 64u S,          S  = |dx ox| + |dy oy| + |dz oz|
//     c = |o|^2 - r^2                 W(c) <= 32u (O2 + r^2), O2 = |o|^2
//     root = EF(sqrt(disc), eps root) W <= 4u root            (the reference does NOT propagate a, b, c into the root)
//     q = -(b +- root) / 2            W(q) <= 32u S + 12u |q|
//     t0 = q / a                      W <= (W(q) + |t0| W(a)) / (a - W(a)) + 4u |t0|
//     t1 = c / q                      W <= (W(c) + |t1| W(q)) / (|q| - W(q)) + 4u |t1|
//   The code below uses FOUR TIMES these constants, and gives up (returns -1: the caller runs the interval path) whenever a
//   comparison falls inside the bound, an operand is huge, an input error is not negligible, or q's interval could touch 0.
//   A decided comparison is the comparison the interval path makes, so hit / miss, the root chosen and tHit are bit-identical.
// Returns 1 = hit (*tHit, *which as sphere_select), 0 = miss, -1 = undecided.
GP_HD int sphere_full_fast(const Ray& ray, V3 oe, V3 de, double radius, double* tHit, int* which, double* dbg = nullptr) {
  const double ox = ray.o.x, oy = ray.o.y, oz = ray.o.z, dx = ray.d.x, dy = ray.d.y, dz = ray.d.z;
  const double m = fmax(fmax(fmax(fabs(ox), fabs(oy)), fmax(fabs(oz), fabs(dx))), fmax(fmax(fabs(dy), fabs(dz)), fabs(radius)));
  const double me = fmax(fmax(fmax(fabs(oe.x), fabs(oe.y)), fmax(fabs(oe.z), fabs(de.x))), fmax(fabs(de.y), fabs(de.z)));
  if (!(m < 1e100) || !(me <= 1e-290)) return -1;
  const double av = ((dx * dx) + (dy * dy)) + (dz * dz);
  const double bv = (((dx * ox) + (dy * oy)) + (dz * oz)) * 2.0;
  const double cv = (((ox * ox) + (oy * oy)) + (oz * oz)) - (radius * radius);
  const double disc = bv * bv - 4. * av * cv;
  if (disc < 0) return 0;  // efloat.Quadratic's own early return (efloat/math.go:36-40)
  if (!(av > 1e-200)) return -1;
  const double root = sqrt(disc);
  const double q = (bv < 0) ? (bv - root) * -0.5 : (bv + root) * -0.5;
  double r0 = q / av, r1 = cv / q;
  const double u = 1.1102230246251565e-16, tiny = 1e-280;
  const double S = fabs(dx * ox) + fabs(dy * oy) + fabs(dz * oz);
  const double O2 = ox * ox + oy * oy + oz * oz;
  const double Wq = 128 * u * S + 48 * u * fabs(q) + tiny;
  const double Wc = 128 * u * (O2 + radius * radius) + tiny;
  if (!(fabs(q) > 2 * Wq)) return -1;
  double W0 = 1.01 * (Wq / av) + 128 * u * fabs(r0) + tiny;  // |t0| W(a) / a <= 96u |t0|, + 16u |t0| of rounding, + 1 / (1 - 96u)
  double W1 = 2 * (Wc + fabs(r1) * Wq) / fabs(q) + 16 * u * fabs(r1) + tiny;
  if (r0 > r1) { double t = r0; r0 = r1; r1 = t; t = W0; W0 = W1; W1 = t; }  // efloat/math.go:55-57 (swap on .v)
  if (dbg) { dbg[0] = r0; dbg[1] = W0; dbg[2] = r1; dbg[3] = W1; }  // (verification build: the roots and their assumed bounds)
  const double tmax = ray.tmax;
  if (r0 > tmax) return 0;                  // t0.hi >= t0.v > tMax
  if (!(r0 + W0 <= tmax)) return -1;
  if (r1 <= 0) return 0;                    // t1.lo <= t1.v <= 0
  if (!(r1 - W1 > 0)) return -1;
  if (r0 <= 0) {                            // t0.lo <= t0.v <= 0: the second root is the candidate
    if (r1 > tmax) return 0;
    if (!(r1 + W1 <= tmax)) return -1;
    *tHit = r1;
    if (which) *which = 1;
    return 1;
  }
  if (!(r0 - W0 > 0)) return -1;
  *tHit = r0;
  if (which) *which = 0;
  return 1;
}

GP_HD V3 sphere_refine(const Ray& ray, double t, double radius) {  // sphere.go:98-104
  V3 pHit = ray.o + ray.d * t;
  pHit = pHit * (radius / sqrt(dist2(pHit, mk3(0, 0, 0))));
  if (pHit.x == 0.0 && pHit.y == 0.0) pHit.x = 1e-5 * radius;
  return pHit;
}
GP_HD double phi_of(V3 pHit) {
  double phi = go_atan2(pHit.y, pHit.x);
  if (phi < 0.0) phi += 2 * kPi;
  return phi;
}
// full root selection incl. the clip test.  which: 0 = t0 selected directly, 1 = t1 because t0.lo <= 0, 2 = t1 after
// t0 was clipped (then the reference keeps the FIRST phi for u: `phi :=` shadows, sphere.go:127).
GP_HD bool sphere_select(const Ray& ray, EF t0, EF t1, double radius, double zMin, double zMax, double phiMax, bool full, double* tHit,
                         int* which) {
  if (t0.hi > ray.tmax || t1.lo <= 0) return false;
  EF ts = t0;
  int w = 0;
  if (ts.lo <= 0) {
    ts = t1;
    w = 1;
    if (ts.hi > ray.tmax) return false;
  }
  if (!full) {
    V3 pHit = sphere_refine(ray, ts.v, radius);
    double phi = phi_of(pHit);
    if ((zMin > -radius && pHit.z < zMin) || (zMax < radius && pHit.z > zMax) || phi > phiMax) {
      if (w == 1) return false;
      if (t1.hi > ray.tmax) return false;
      ts = t1;
      w = 2;
      pHit = sphere_refine(ray, ts.v, radius);
      double phi2 = phi_of(pHit);
      if ((zMin > -radius && pHit.z < zMin) || (zMax < radius && pHit.z > zMax) || phi2 > phiMax) return false;
    }
  }
  *tHit = ts.v;
  if (which) *which = w;
  return true;
}

// ---- Disk.Intersect / IntersectP (disk.go:64-93,127-159) on the object-space ray ----
GP_HD bool disk_test(const Ray& ray, double height, double radius, double innerRadius, double phiMax, double* tHit) {
  if (ray.d.z == 0) return false;
  double t = (height - ray.o.z) / ray.d.z;
  if (t <= 0 || t >= ray.tmax) return false;
  V3 pHit = ray.o + ray.d * t;
  double d2 = pHit.x * pHit.x + pHit.y * pHit.y;
  if (d2 > radius * radius || d2 < innerRadius * innerRadius) return false;
  double phi = go_atan2(pHit.y, pHit.x);
  if (phi < 0) phi += 2 * kPi;
  if (phi > phiMax) return false;
  *tHit = t;
  return true;
}

// ---- Triangle (new shape, defined by this backend + its oracle): watertight test in float64 ----
GP_HD int max_dim(V3 v) { return (v.x > v.y) ? ((v.x > v.z) ? 0 : 2) : ((v.y > v.z) ? 1 : 2); }
GP_HD V3 permute(V3 v, int x, int y, int z) { return mk3(comp(v, x), comp(v, y), comp(v, z)); }
// The watertight test only ever permutes cyclically — (kx, ky, kz) = (kz+1, kz+2, kz) mod 3 — so the permutation is a
// three-way choice keyed by kz.  Done with bit masks on the float64 bit patterns: a pure selection (bit-exact), and
// branch-free by construction (the generic comp() chain compiled to ~370 instructions of divergent code per triangle).
GP_HD V3 permute_cyclic(V3 v, int kz) {
  const uint64_t m0 = (uint64_t)0 - (uint64_t)(kz == 0), m1 = (uint64_t)0 - (uint64_t)(kz == 1), m2 = ~(m0 | m1);
  const uint64_t x = f2b(v.x), y = f2b(v.y), z = f2b(v.z);
  return mk3(b2f((y & m0) | (z & m1) | (x & m2)),   // component kx: y, z, x for kz = 0, 1, 2
             b2f((z & m0) | (x & m1) | (y & m2)),   // component ky
             b2f((x & m0) | (y & m1) | (z & m2)));  // component kz
}
// per-ray part of the watertight test (depends on the ray direction only): permutation and shear constants
struct TriRay { int kx, ky, kz; double Sx, Sy, Sz; };
GP_HD TriRay tri_ray_setup(V3 dir) {
  TriRay tr;
  tr.kz = max_dim(vabs(dir));
  tr.kx = tr.kz + 1; if (tr.kx == 3) tr.kx = 0;
  tr.ky = tr.kx + 1; if (tr.ky == 3) tr.ky = 0;
  V3 d = permute_cyclic(dir, tr.kz);
  tr.Sx = -d.x / d.z; tr.Sy = -d.y / d.z; tr.Sz = 1.0 / d.z;
  return tr;
}
GP_HD bool tri_test_pre(V3 p0, V3 p1, V3 p2, const Ray& ray, const TriRay& tr, double* tHit, double* bary) {
  V3 p0t = p0 - ray.o, p1t = p1 - ray.o, p2t = p2 - ray.o;
  int kx = tr.kx, ky = tr.ky, kz = tr.kz;
  p0t = permute_cyclic(p0t, kz); p1t = permute_cyclic(p1t, kz); p2t = permute_cyclic(p2t, kz);
  (void)kx; (void)ky;
  double Sx = tr.Sx, Sy = tr.Sy, Sz = tr.Sz;
  p0t.x += Sx * p0t.z; p0t.y += Sy * p0t.z;
  p1t.x += Sx * p1t.z; p1t.y += Sy * p1t.z;
  p2t.x += Sx * p2t.z; p2t.y += Sy * p2t.z;
  double e0 = p1t.x * p2t.y - p1t.y * p2t.x;
  double e1 = p2t.x * p0t.y - p2t.y * p0t.x;
  double e2 = p0t.x * p1t.y - p0t.y * p1t.x;
  if (((e0 < 0) | (e1 < 0) | (e2 < 0)) & ((e0 > 0) | (e1 > 0) | (e2 > 0))) return false;  // same truth table, no short-circuit branches
  double det = e0 + e1 + e2;
  if (det == 0) return false;
  p0t.z *= Sz; p1t.z *= Sz; p2t.z *= Sz;
  double tScaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
  if (det < 0 && (tScaled >= 0 || tScaled < ray.tmax * det)) return false;
  if (det > 0 && (tScaled <= 0 || tScaled > ray.tmax * det)) return false;
  double invDet = 1 / det;
  double t = tScaled * invDet;
  if (t <= 0 || t >= ray.tmax) return false;
  *tHit = t;
  if (bary) { bary[0] = e0 * invDet; bary[1] = e1 * invDet; bary[2] = e2 * invDet; }
  return true;
}
// The same test with the vertices read through the permutation: v = the triangle's nine coordinates {p0, p1, p2} in memory
// (the flat aggregate keeps them in shared memory), fetched as v[kx], v[ky], v[kz] — nine indexed loads instead of five
// vector loads and 27 mask operations on 64-bit patterns; permute(p - o) == permute(p) - permute(o) component by component,
// so every subtraction is the one tri_test_pre makes.
GP_D bool tri_test_idx(const double* v, const Ray& ray, const TriRay& tr, double* tHit) {
  const int kx = tr.kx, ky = tr.ky, kz = tr.kz;
  const V3 o = permute_cyclic(ray.o, kz);
  V3 p0t = mk3(v[kx] - o.x, v[ky] - o.y, v[kz] - o.z);
  V3 p1t = mk3(v[3 + kx] - o.x, v[3 + ky] - o.y, v[3 + kz] - o.z);
  V3 p2t = mk3(v[6 + kx] - o.x, v[6 + ky] - o.y, v[6 + kz] - o.z);
  double Sx = tr.Sx, Sy = tr.Sy, Sz = tr.Sz;
  p0t.x += Sx * p0t.z; p0t.y += Sy * p0t.z;
  p1t.x += Sx * p1t.z; p1t.y += Sy * p1t.z;
  p2t.x += Sx * p2t.z; p2t.y += Sy * p2t.z;
  double e0 = p1t.x * p2t.y - p1t.y * p2t.x;
  double e1 = p2t.x * p0t.y - p2t.y * p0t.x;
  double e2 = p0t.x * p1t.y - p0t.y * p1t.x;
  if (((e0 < 0) | (e1 < 0) | (e2 < 0)) & ((e0 > 0) | (e1 > 0) | (e2 > 0))) return false;
  double det = e0 + e1 + e2;
  if (det == 0) return false;
  p0t.z *= Sz; p1t.z *= Sz; p2t.z *= Sz;
  double tScaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
  if (det < 0 && (tScaled >= 0 || tScaled < ray.tmax * det)) return false;
  if (det > 0 && (tScaled <= 0 || tScaled > ray.tmax * det)) return false;
  double invDet = 1 / det;
  double t = tScaled * invDet;
  if (t <= 0 || t >= ray.tmax) return false;
  *tHit = t;
  return true;
}
GP_HD bool tri_test(V3 p0, V3 p1, V3 p2, const Ray& ray, double* tHit, double* bary) {
  return tri_test_pre(p0, p1, p2, ray, tri_ray_setup(ray.d), tHit, bary);
}

// ---- the surface interaction fields Path.Li reads (pkg/pbrt/interaction.go:23-30,123-148) ----
struct Hit {
  V3 p, perr, n, wo, ns, sdpdu;
  double u, v;
};
// Transform.TransformSurfaceInteraction (transform.go:302-334) on those fields, with the reference's pointer-aliasing
// outcome: Shading.Normal is the un-normalised transform of the previous one, then FaceForward'ed.
GP_HD void xf_hit(const M4& m, const M4& inv, Hit& h) {
  V3 perr;
  V3 p = xf_point(m, h.p, h.perr, &perr);
  h.p = p; h.perr = perr;
  h.n = normalized(xf_normal_inv(inv, h.n));
  h.wo = normalized(xf_vector(m, h.wo));
  h.ns = xf_normal_inv(inv, h.ns);
  h.sdpdu = xf_vector(m, h.sdpdu);
  h.ns = faceforward(h.ns, h.n);
}

// GP_SPHERE_FAST_PATH: the decision of sphere_full_fast when it has one.  -DGP_CHECK_FAST_SPHERE (verification build) runs the
// interval path as well and raises `bad` (counted as efloat_panics, which every test asserts to be 0) if the two ever disagree
// or an interval is wider than the bound the fast path assumed; -DGP_NO_FAST_SPHERE disables the fast path.
#if defined(GP_NO_FAST_SPHERE)
#define GP_SPHERE_FAST_PATH(ray, oe, de, radius, tHit, bad)
#elif defined(GP_CHECK_FAST_SPHERE)
#define GP_SPHERE_FAST_PATH(ray, oe, de, radius, tHit, bad)                                                   \
  {                                                                                                           \
    double tf_ = 0, dbg_[4] = {0, -1, 0, -1};                                                                 \
    int wf_ = -1;                                                                                             \
    const int f_ = sphere_full_fast(ray, oe, de, radius, &tf_, &wf_, dbg_);                                   \
    EF c0_, c1_;                                                                                              \
    double ts_ = 0;                                                                                           \
    int ws_ = -1;                                                                                             \
    const bool roots_ = sphere_roots(ray, oe, de, radius, &c0_, &c1_, bad);                                   \
    const bool h_ = roots_ && sphere_select(ray, c0_, c1_, radius, 0, 0, 0, true, &ts_, &ws_);                \
    if (f_ >= 0 && (h_ != (f_ == 1) || (h_ && (ts_ != tf_ || ws_ != wf_)))) bad = 1;                          \
    if (roots_ && dbg_[1] >= 0 &&                                                                             \
        (c0_.v != dbg_[0] || c1_.v != dbg_[2] || !(c0_.hi - c0_.v <= dbg_[1]) || !(c0_.v - c0_.lo <= dbg_[1]) ||  \
         !(c1_.hi - c1_.v <= dbg_[3]) || !(c1_.v - c1_.lo <= dbg_[3])))                                         \
      bad = 1;                                                                                                \
  }
#else
#define GP_SPHERE_FAST_PATH(ray, oe, de, radius, tHit, bad)                                                   \
  {                                                                                                           \
    const int f_ = sphere_full_fast(ray, oe, de, radius, tHit, nullptr);                                      \
    if (f_ >= 0) return f_ == 1;                                                                              \
  }
#endif

// generic primitive test used by both traversal kernels.  Returns true and the hit t (the new r.TMax,
// primitive.go:51,102) if the primitive is hit within ray.tmax.  `bad` collects efloat.Check panics.
// sphere / disk test (the long float64 + EFloat path); triangles are tested inline by the traversal kernel
GP_D bool quadric_test(const DevScene& sc, const PrimRec* rec, uint32_t flags, const Ray& wray, double* tHit, int& bad) {
  uint32_t kind = flags & RK_KIND_MASK;
  Ray ray = wray;
  if (kind == RK_SPHERE && (flags & RF_FAST)) {
    const double* d = rec->d;
    // TransformedPrimitive.Intersect: ray through primitiveToWorld.Inverse() (primitive.go:94-96)
    if (flags & RF_HAS_P2W) ray = xf_ray(translation_m4(d[4], d[5], d[6]), ray, nullptr, nullptr);
    V3 oe, de;
    ray = xf_ray(translation_m4(d[1], d[2], d[3]), ray, &oe, &de);  // worldToObject.TransformRay (sphere.go:65)
    GP_SPHERE_FAST_PATH(ray, oe, de, d[0], tHit, bad)
    EF t0, t1;
    if (!sphere_roots(ray, oe, de, d[0], &t0, &t1, bad)) return false;
    return sphere_select(ray, t0, t1, d[0], 0, 0, 0, true, tHit, nullptr);
  }
  // general path: look the shape up
  int4 pr = sc.prims[rec->prim];
  if (pr.w >= 0) ray = xf_ray(load_m4(sc, pr.w, true), ray, nullptr, nullptr);
  if (kind == RK_SPHERE) {
    SphereDev s = sc.spheres[pr.y];
    V3 oe, de;
    ray = xf_ray(load_m4(sc, s.xf, true), ray, &oe, &de);
    if (flags & RF_FULL) { GP_SPHERE_FAST_PATH(ray, oe, de, s.radius, tHit, bad) }
    EF t0, t1;
    if (!sphere_roots(ray, oe, de, s.radius, &t0, &t1, bad)) return false;
    return sphere_select(ray, t0, t1, s.radius, s.zMin, s.zMax, s.phiMax, (flags & RF_FULL) != 0, tHit, nullptr);
  }
  DiskDev dk = sc.disks[pr.y];
  ray = xf_ray(load_m4(sc, dk.xf, true), ray, nullptr, nullptr);
  return disk_test(ray, dk.height, dk.radius, dk.innerRadius, dk.phiMax, tHit);
}

// Rebuild the SurfaceInteraction of a known hit (world ray + primitive id + tHit).  The traversal kernel keeps only
// (primitive, t); everything here is a deterministic function of those, so recomputing it in the shade stage gives
// the same bits the reference computes inside Shape.Intersect (sphere.go:137-185, disk.go:95-123) followed by
// TransformedPrimitive.Intersect (primitive.go:104-106).
// tri_only (a compile-time constant at the call site): the hit's shade class says it is a triangle
GP_D void hit_record(const DevScene& sc, int rec_index, const Ray& wray_in, double tHit, Hit* h, int* prim_out, int& bad, bool tri_only = false) {
  const PrimRec* rec = sc.recs + rec_index;
  uint32_t rflags = rec->flags;
  int prim = (int)rec->prim;
  *prim_out = prim;
  int4 pr = sc.prims[prim];
  Ray ray = wray_in;
  if (tri_only || pr.x == RK_TRIANGLE) {
    const double* d = rec->d;
    V3 p0 = mk3(d[0], d[1], d[2]), p1 = mk3(d[3], d[4], d[5]), p2 = mk3(d[6], d[7], d[8]);
    double t, b[3];
    Ray r2 = ray;
    r2.tmax = d_inf();
    // barycentrics do not depend on tMax; the test cannot fail for a ray that hit this triangle at tHit
    tri_test(p0, p1, p2, r2, &t, b);
    V3 dp02 = p0 - p2, dp12 = p1 - p2;
    double du02 = -1, dv02 = -1, du12 = 0, dv12 = -1;
    double determinant = du02 * dv12 - dv02 * du12;
    double invdet = 1 / determinant;
    V3 dpdu = (dp02 * dv12 - dp12 * dv02) * invdet;
    V3 dpdv = (dp02 * -du12 + dp12 * du02) * invdet;
    V3 n = normalized(cross(dp02, dp12));
    if (len2(cross(dpdu, dpdv)) == 0) { V3 tmp; coordinate_system(n, &dpdu, &tmp); }
    V3 pAbs = vabs(p0 * b[0]) + vabs(p1 * b[1]) + vabs(p2 * b[2]);
    h->perr = pAbs * gamma_n(7);
    h->p = p0 * b[0] + p1 * b[1] + p2 * b[2];
    if (rflags & RF_REVERSE) n = n * -1.0;
    h->n = n;
    h->wo = ray.d * -1.0;
    h->ns = n;
    h->sdpdu = dpdu;
    h->u = b[0] * 0 + b[1] * 1 + b[2] * 1;
    h->v = b[0] * 0 + b[1] * 0 + b[2] * 1;
    return;
  }
  M4 p2w_inv;
  if (pr.w >= 0) { p2w_inv = load_m4_plain(sc, pr.w, true); ray = xf_ray(p2w_inv, ray, nullptr, nullptr); }
  int sxf;
  if (pr.x == RK_SPHERE) {
    SphereDev s = sc.spheres[pr.y];
    sxf = s.xf;
    M4 w2o = load_m4_plain(sc, s.xf, true);
    V3 oe, de;
    ray = xf_ray(w2o, ray, &oe, &de);
    V3 pHit = sphere_refine(ray, tHit, s.radius);
    double phi = phi_of(pHit);
    if (!(s.flags & RF_FULL)) {
      // partial sphere: if the reference reached tHit through the clip branch, u keeps the phi of the clipped t0 root
      EF t0, t1;
      if (sphere_roots(ray, oe, de, s.radius, &t0, &t1, bad) && t0.lo > 0 && tHit != t0.v) phi = phi_of(sphere_refine(ray, t0.v, s.radius));
    }
    double u = phi / s.phiMax;
    double theta = go_acos(go_clamp(pHit.z / s.radius, -1, 1));
    double v = (theta - s.thetaMin) / (s.thetaMax - s.thetaMin);
    double zRadius = sqrt(pHit.x * pHit.x + pHit.y * pHit.y);
    double invZ = 1.0 / zRadius;
    double cosPhi = pHit.x * invZ, sinPhi = pHit.y * invZ;
    V3 dpdu = mk3(-s.phiMax * pHit.y, s.phiMax * pHit.x, 0);
    V3 dpdv = mk3(pHit.z * cosPhi, pHit.z * sinPhi, -s.radius * go_sin(theta)) * (s.thetaMax - s.thetaMin);
    V3 n = normalized(cross(dpdu, dpdv));
    if (s.flags & RF_REVERSE) n = n * -1.0;  // interaction.go:179-181
    h->p = pHit; h->perr = vabs(pHit) * gamma_n(5); h->n = n; h->wo = ray.d * -1.0; h->ns = n; h->sdpdu = dpdu; h->u = u; h->v = v;
  } else {
    DiskDev dk = sc.disks[pr.y];
    sxf = dk.xf;
    ray = xf_ray(load_m4_plain(sc, dk.xf, true), ray, nullptr, nullptr);
    V3 pHit = ray.o + ray.d * tHit;
    double d2 = pHit.x * pHit.x + pHit.y * pHit.y;
    double phi = phi_of(pHit);
    double u = phi / dk.phiMax;
    double rHit = sqrt(d2);
    double oneMinusV = (rHit - dk.innerRadius) / (dk.radius - dk.innerRadius);
    double v = 1 - oneMinusV;
    V3 dpdu = mk3(-dk.phiMax * pHit.y, dk.phiMax * pHit.x, 0);
    V3 dpdv = mk3(pHit.x, pHit.y, 0) * ((dk.radius - dk.innerRadius) / rHit);  // normal comes out -z
    pHit.z = dk.height;
    V3 n = normalized(cross(dpdu, dpdv));
    if (dk.flags & RF_REVERSE) n = n * -1.0;
    h->p = pHit; h->perr = mk3(0, 0, 0); h->n = n; h->wo = ray.d * -1.0; h->ns = n; h->sdpdu = dpdu; h->u = u; h->v = v;
  }
  xf_hit(load_m4_plain(sc, sxf, false), load_m4_plain(sc, sxf, true), *h);  // sphere.go:185 / disk.go:110
  if (pr.w >= 0 && !(sc.xf_flags[pr.w] & XF_IDENTITY)) xf_hit(load_m4_plain(sc, pr.w, false), p2w_inv, *h);  // primitive.go:104-106
}

}  // namespace gp
