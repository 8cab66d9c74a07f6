// gp_math.cuh — float64 numeric core of libgopbrt_cuda (host + device).
//
// Everything on the intersection / spawn / shading path is float64 and must round exactly like the Go reference on
// amd64 (which never contracts x*y+z): this translation unit is compiled with -fmad=false (device) and
// -ffp-contract=off (host).  Go's math.Min/Max/Nextafter and its software Sin/Cos/Atan2/Acos (golang:1.11,
// Cephes-derived, reached through pkg/math/math.go:21-144) are reproduced instruction for instruction so that
// hit/miss, primitive ids, t and the shaded film agree with the reference semantics bit for bit, not just to an ulp.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#define GP_HD __host__ __device__ __forceinline__
#define GP_D __device__ __forceinline__
#define GP_HD_NOINLINE __host__ __device__ __noinline__

namespace gp {

constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double kPiOver2 = kPi / 2;  // exact scalings of the float64 Pi (pkg/math/math.go:13-14 compute the same values)
constexpr double kPiOver4 = kPi / 4;
constexpr double k2Pi = 2 * kPi;
constexpr double kInvPi = 1.0 / kPi;  // float64 division, as pkg/math/math.go:10

GP_HD double d_inf() {
#ifdef __CUDA_ARCH__
  return __longlong_as_double(0x7ff0000000000000LL);
#else
  return INFINITY;
#endif
}
GP_HD uint64_t f2b(double x) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
GP_HD double b2f(uint64_t u) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double x; memcpy(&x, &u, 8); return x;
#endif
}
GP_HD bool is_nan(double x) { return x != x; }
GP_HD bool is_inf(double x) { return fabs(x) == d_inf(); }
GP_HD bool sign_bit(double x) { return (f2b(x) >> 63) != 0; }
GP_HD double copy_sign(double mag, double sgn) { return b2f((f2b(mag) & 0x7fffffffffffffffULL) | (f2b(sgn) & 0x8000000000000000ULL)); }

// math.Nextafter restricted to the two call shapes of pkg/math/math.go:122-128, branch-free.
// NextFloatUp(v) = Nextafter(v, v+1): v+1 == v (|v| >= 2^53 on the positive side, +-Inf) returns v unchanged — NOT the
// true next float; 0 (either sign) -> smallest positive denormal; NaN stays NaN (its payload is irrelevant).
// On the device the step itself is ONE directed-rounding add of the smallest denormal: v + 2^-1074 rounded up IS the next
// representable number above any finite v (exact for zero and denormals, a round-up of an inexact sum otherwise), which
// replaces ten 64-bit integer instructions per call — the interval arithmetic of pkg/efloat calls this twice per operation
// (config 2 frame 145.3 -> 136.9 ms, films bit-identical).  One representable difference: stepping from -+2^-1074 onto
// zero yields the zero of the other sign; no comparison can see it and the next step erases it.
GP_HD double next_up(double v) {
#ifdef __CUDA_ARCH__
  double stepped = __dadd_ru(v, 4.9406564584124654e-324);
#else
  double stepped = (v < 0) ? b2f(f2b(v) - 1) : b2f(f2b(fabs(v)) + 1);
#endif
  return (v + 1 == v) ? v : stepped;
}
GP_HD double next_down(double v) {
#ifdef __CUDA_ARCH__
  double stepped = __dadd_rd(v, -4.9406564584124654e-324);
#else
  double stepped = (v > 0) ? b2f(f2b(v) - 1) : b2f(f2b(-fabs(v)) + 1);
#endif
  return (v - 1 == v) ? v : stepped;
}

// math.Min / math.Max (Go): -Inf/+Inf first, NaN-propagating, signed-zero aware — NOT fmin/fmax.
// On the device: the hardware minimum / maximum (DMNMX: -0 < +0, a NaN operand loses) is math.Min / math.Max in every case
// but one — a NaN operand must win unless the other operand is the infinity that Go tests for first — so the Go function is
// the hardware instruction plus one select (the if-chain compiled to a dozen instructions, and the shade kernels spend 6 % of
// theirs here).  Pinned with every pair of special values by the "go_math" known-answer test.
GP_HD double go_min(double x, double y) {
#ifdef __CUDA_ARCH__
  const double r = fmin(x, y);
  return ((is_nan(x) || is_nan(y)) && r != -d_inf()) ? b2f(0x7ff8000000000001ULL) : r;
#else
  if (x == -d_inf() || y == -d_inf()) return -d_inf();
  if (is_nan(x) || is_nan(y)) return b2f(0x7ff8000000000001ULL);
  if (x == 0 && x == y) return sign_bit(x) ? x : y;
  return x < y ? x : y;
#endif
}
GP_HD double go_max(double x, double y) {
#ifdef __CUDA_ARCH__
  const double r = fmax(x, y);
  return ((is_nan(x) || is_nan(y)) && r != d_inf()) ? b2f(0x7ff8000000000001ULL) : r;
#else
  if (x == d_inf() || y == d_inf()) return d_inf();
  if (is_nan(x) || is_nan(y)) return b2f(0x7ff8000000000001ULL);
  if (x == 0 && x == y) return sign_bit(x) ? y : x;
  return x > y ? x : y;
#endif
}
GP_HD double go_clamp(double v, double lo, double hi) {  // pkg/math/math.go:42-50
  if (v < lo) return lo;
  if (v > hi) return hi;
  return v;
}

// pkg/math/math.go:17-19,82-84: MachineEpsilon = NextFloatUp(0) = 4.94e-324.  The arithmetic is carried
// out for real (denormals are honoured by fp64 on the GPU): gamma(n) = (n*eps)/(1 - n*eps).
GP_HD double machine_epsilon() { return b2f(1); }
GP_HD double one_minus_epsilon() { return b2f(0x3fefffffffffffffULL); }
GP_HD double gamma_n(double n) { double e = machine_epsilon(); return (n * e) / (1 - n * e); }

// ---- Go 1.11 src/math/sin.go (Cephes sin.c): Cody-Waite Pi/4 reduction + degree-6 polynomials ----
#define GP_PI4A 7.85398125648498535156e-1
#define GP_PI4B 3.77489470793079817668e-8
#define GP_PI4C 2.69515142907905952645e-15
#define GP_M4PI 1.273239544735162542821171882678754627704620361328125

GP_HD double poly_sin(double z, double zz) {
  return z + z * zz * ((((((1.58962301576546568060e-10 * zz) + -2.50507477628578072866e-8) * zz + 2.75573136213857245213e-6) * zz +
                          -1.98412698295895385996e-4) * zz + 8.33333333332211858878e-3) * zz + -1.66666666666666307295e-1);
}
GP_HD double poly_cos(double zz) {
  return 1.0 - 0.5 * zz + zz * zz * ((((((-1.13585365213876817300e-11 * zz) + 2.08757008419747316778e-9) * zz + -2.75573141792967388112e-7) * zz +
                                       2.48015872888517045348e-5) * zz + -1.38888888888730564116e-3) * zz + 4.16666666666665929218e-2);
}
GP_HD_NOINLINE static double go_cos(double x) {
  if (is_nan(x) || is_inf(x)) return b2f(0x7ff8000000000001ULL);
  bool sign = false;
  x = fabs(x);
  uint64_t j = (uint64_t)(long long)(x * GP_M4PI);
  double y = (double)(long long)j;
  if (j & 1) { j++; y++; }
  j &= 7;
  if (j > 3) { j -= 4; sign = !sign; }
  if (j > 1) sign = !sign;
  double z = ((x - y * GP_PI4A) - y * GP_PI4B) - y * GP_PI4C;
  double zz = z * z;
  y = (j == 1 || j == 2) ? poly_sin(z, zz) : poly_cos(zz);
  return sign ? -y : y;
}
GP_HD_NOINLINE static double go_sin(double x) {
  if (x == 0 || is_nan(x)) return x;
  if (is_inf(x)) return b2f(0x7ff8000000000001ULL);
  bool sign = false;
  if (x < 0) { x = -x; sign = true; }
  uint64_t j = (uint64_t)(long long)(x * GP_M4PI);
  double y = (double)(long long)j;
  if (j & 1) { j++; y++; }
  j &= 7;
  if (j > 3) { sign = !sign; j -= 4; }
  double z = ((x - y * GP_PI4A) - y * GP_PI4B) - y * GP_PI4C;
  double zz = z * z;
  y = (j == 1 || j == 2) ? poly_cos(zz) : poly_sin(z, zz);
  return sign ? -y : y;
}
// math.Sin(x) and math.Cos(x) of the SAME argument (the sampling warps take both of one angle): sin.go's two functions make the
// same range reduction — |x|, the octant j, the three-term Cody-Waite remainder z — and each then evaluates one of the same
// two polynomials.  Here the reduction and both polynomials are computed once and each result selects its own; every value is
// produced by exactly the operations its own function performs, so both are bit-identical to go_sin / go_cos, at half the
// arithmetic and without the octant branch (a warp's lanes fall in different octants: both polynomials ran anyway, twice).
// (inlined, results in registers: as an out-of-line function its return sequence alone was 2.7 % of the Lambert shade class's
// instructions at 6 of 32 lanes, and pointer results would travel through the thread's local-memory stack)
struct SinCos { double sn, cs; };
GP_HD SinCos go_sincos(double x) {
  const double qnan = b2f(0x7ff8000000000001ULL);
  const double xa = fabs(x);
  uint64_t j = (uint64_t)(long long)(xa * GP_M4PI);
  double y = (double)(long long)j;
  if (j & 1) { j++; y++; }
  j &= 7;
  bool sign_s = x < 0, sign_c = false;
  if (j > 3) { j -= 4; sign_s = !sign_s; sign_c = !sign_c; }
  if (j > 1) sign_c = !sign_c;
  const double z = ((xa - y * GP_PI4A) - y * GP_PI4B) - y * GP_PI4C;
  const double zz = z * z;
  const double ps = poly_sin(z, zz), pc = poly_cos(zz);
  const bool swap = (j == 1 || j == 2);
  const double ys = swap ? pc : ps, yc = swap ? ps : pc;
  const bool bad = is_nan(x) || is_inf(x);
  SinCos r;
  r.cs = bad ? qnan : (sign_c ? -yc : yc);
  r.sn = (x == 0 || is_nan(x)) ? x : (is_inf(x) ? qnan : (sign_s ? -ys : ys));
  return r;
}
// ---- Go src/math/atan.go (Cephes atan.c) ----
GP_HD double go_xatan(double x) {
  double z = x * x;
  z = z * ((((-8.750608600031904122785e-01 * z + -1.615753718733365076637e+01) * z + -7.500855792314704667340e+01) * z +
            -1.228866684490136173410e+02) * z + -6.485021904942025371773e+01) /
      (((((z + 2.485846490142306297962e+01) * z + 1.650270098316988542046e+02) * z + 4.328810604912902668951e+02) * z +
        4.853903996359136964868e+02) * z + 1.945506571482613964425e+02);
  z = x * z + x;
  return z;
}
// The three range cases of satan evaluate the same polynomial on different arguments.  Written with selects — one argument,
// ONE polynomial evaluation, the case's own additions afterwards — the lanes of a warp stay together instead of running the
// polynomial three times at a third of the lanes each; every case performs exactly the operations of its branch in atan.go,
// in the same order, so the result is bit-identical (ncu: the sphere shade classes spent 12 % of their instructions here at
// 6 of 32 lanes).
GP_HD double go_satan(double x) {
  const double Morebits = 6.123233995736765886130e-17;
  const double Tan3pio8 = 2.41421356237309504880;
  const bool lo = x <= 0.66, hi = x > Tan3pio8;   // (a NaN takes the last form, as it falls through both ifs in atan.go)
  const double num = hi ? 1.0 : (x - 1), den = hi ? x : (x + 1);
  const double r = go_xatan(lo ? x : num / den);
  const double r_hi = kPiOver2 - r + Morebits;
  const double r_mid = kPiOver4 + r + 0.5 * Morebits;
  return lo ? r : (hi ? r_hi : r_mid);
}
GP_HD double go_atan(double x) {
  const double r = go_satan(fabs(x));              // x > 0: satan(x); x < 0: -satan(-x); x == 0: x itself (keeps the zero's sign)
  return (x == 0) ? x : ((x > 0) ? r : -r);
}
GP_HD_NOINLINE static double go_atan2(double y, double x) {  // src/math/atan2.go
  if (is_nan(y) || is_nan(x)) return b2f(0x7ff8000000000001ULL);
  if (y == 0) {
    if (x >= 0 && !sign_bit(x)) return copy_sign(0.0, y);
    return copy_sign(kPi, y);
  }
  if (x == 0) return copy_sign(kPiOver2, y);
  if (is_inf(x)) {
    if (x > 0) return is_inf(y) ? copy_sign(kPiOver4, y) : copy_sign(0.0, y);
    return is_inf(y) ? copy_sign(3 * kPi / 4, y) : copy_sign(kPi, y);
  }
  if (is_inf(y)) return copy_sign(kPiOver2, y);
  double q = go_atan(y / x);
  if (x < 0) return q <= 0 ? q + kPi : q - kPi;
  return q;
}
GP_HD double go_asin(double x) {  // src/math/asin.go
  if (x == 0) return x;
  bool sign = false;
  if (x < 0) { x = -x; sign = true; }
  if (x > 1) return b2f(0x7ff8000000000001ULL);
  double temp = sqrt(1 - x * x);
  const bool big = x > 0.7;                       // one satan call for both cases (see go_satan)
  const double s = go_satan((big ? temp : x) / (big ? x : temp));
  temp = big ? kPiOver2 - s : s;
  return sign ? -temp : temp;
}
GP_HD_NOINLINE static double go_acos(double x) { return kPiOver2 - go_asin(x); }

// ---- pkg/geometry/xyz.go:424-614 ----
struct V3 { double x, y, z; };
GP_HD V3 mk3(double x, double y, double z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
GP_HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
GP_HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
GP_HD V3 operator*(V3 a, double s) { return mk3(a.x * s, a.y * s, a.z * s); }
GP_HD V3 operator/(V3 a, double s) { return mk3(a.x / s, a.y / s, a.z / s); }
GP_HD V3 vabs(V3 a) { return mk3(fabs(a.x), fabs(a.y), fabs(a.z)); }
GP_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GP_HD double len2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
GP_HD double dist2(V3 self, V3 other) { return len2(other - self); }  // xyz.go:579-581
GP_HD V3 cross(V3 a, V3 b) { return mk3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)); }
GP_HD V3 normalized(V3 a) {  // xyz.go:587-606: times 1/sqrt
  double n2 = len2(a);
  if (n2 > 0) {
    double inv = 1.0 / sqrt(n2);
    a.x *= inv; a.y *= inv; a.z *= inv;
  }
  return a;
}
GP_HD V3 faceforward(V3 n1, V3 n2) { return dot(n1, n2) < 0.0 ? n1 * -1.0 : n1; }  // geometry.go:115-120
GP_HD double comp(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
// geometry.go:47-60 — divides by the squared length
GP_HD void coordinate_system(V3 v1, V3* v2, V3* v3) {
  if (fabs(v1.x) > fabs(v1.y)) {
    double v = v1.x * v1.x + v1.z * v1.z;
    *v2 = mk3(-v1.z / v, 0 / v, v1.x / v);
  } else {
    double v = v1.y * v1.y + v1.z * v1.z;
    *v2 = mk3(0 / v, v1.z / v, -v1.y / v);
  }
  *v3 = cross(v1, *v2);
}

// ---- pkg/efloat (efloat.go, math.go) ----
// `bad` accumulates the conditions on which efloat.Check panics (efloat.go:102-111): Inf/NaN bounds or Low > High; the
// kernels count them (the reference process would be dead at that point, so nothing after a panic needs to agree).
// The interval bounds use fmin/fmax instead of Go's math.Min/Max: they differ only for NaN operands (a panic either
// way) and in the sign of a zero result, which next_down/next_up erase (both zeros step to the same denormal).
struct EF { double v, lo, hi; };
GP_HD void ef_check(const EF& f, int& bad) {
  if (!(f.lo <= f.hi && fabs(f.lo) < d_inf() && fabs(f.hi) < d_inf())) bad = 1;
}
GP_HD EF ef_new(double v, double err, int& bad) {  // efloat.go:10-22
  EF f; f.v = v;
  f.lo = (err != 0) ? next_down(v - err) : v;
  f.hi = (err != 0) ? next_up(v + err) : v;
  ef_check(f, bad);
  return f;
}
GP_HD EF ef_add(EF f, EF o, int& bad) {
  f.v = f.v + o.v; f.lo = next_down(f.lo + o.lo); f.hi = next_up(f.hi + o.hi);
  ef_check(f, bad);
  return f;
}
GP_HD EF ef_sub(EF f, EF o, int& bad) {
  f.v = f.v - o.v; f.lo = next_down(f.lo - o.hi); f.hi = next_up(f.hi - o.lo);
  ef_check(f, bad);
  return f;
}
GP_HD EF ef_mul(EF f, EF o, int& bad) {
  double p0 = f.lo * o.lo, p1 = f.hi * o.lo, p2 = f.lo * o.hi, p3 = f.hi * o.hi;
  f.v = f.v * o.v;
  f.lo = next_down(fmin(fmin(p0, p1), fmin(p2, p3)));
  f.hi = next_up(fmax(fmax(p0, p1), fmax(p2, p3)));
  if (is_nan(p0) || is_nan(p1) || is_nan(p2) || is_nan(p3)) bad = 1;  // Go's Min/Max would have propagated the NaN into Check
  ef_check(f, bad);
  return f;
}
GP_HD EF ef_div(EF f, EF o, int& bad) {
  f.v = f.v / o.v;
  if (o.lo < 0 && o.hi > 0) { f.lo = -d_inf(); f.hi = d_inf(); bad = 1; }
  else {
    double d0 = f.lo / o.lo, d1 = f.hi / o.lo, d2 = f.lo / o.hi, d3 = f.hi / o.hi;
    f.lo = next_down(fmin(fmin(d0, d1), fmin(d2, d3)));
    f.hi = next_up(fmax(fmax(d0, d1), fmax(d2, d3)));
    if (is_nan(d0) || is_nan(d1) || is_nan(d2) || is_nan(d3)) bad = 1;
  }
  ef_check(f, bad);
  return f;
}
GP_HD EF ef_muls(EF f, double s, int& bad) { return ef_mul(f, ef_new(s, 0.0, bad), bad); }
GP_HD bool ef_quadratic(EF a, EF b, EF c, EF* t0, EF* t1, int& bad) {  // efloat/math.go:35-59
  double disc = b.v * b.v - 4. * a.v * c.v;
  if (disc < 0) return false;
  double root = sqrt(disc);
  EF fr = ef_new(root, machine_epsilon() * root, bad);
  EF q = (b.v < 0) ? ef_muls(ef_sub(b, fr, bad), -0.5, bad) : ef_muls(ef_add(b, fr, bad), -0.5, bad);
  EF r0 = ef_div(q, a, bad), r1 = ef_div(c, q, bad);
  if (r0.v > r1.v) { EF t = r0; r0 = r1; r1 = t; }
  *t0 = r0; *t1 = r1;
  return true;
}

// ---- pkg/pbrt/transform.go: 4x4 row-major matrix held in registers ----
struct M4 { double m[4][4]; };

// TransformPoint (transform.go:227-247) incl. the asymmetric error expression
GP_HD V3 xf_point(const M4& t, V3 p, V3 pe, V3* err) {
  double xp = t.m[0][0] * p.x + t.m[0][1] * p.y + t.m[0][2] * p.z + t.m[0][3];
  double yp = t.m[1][0] * p.x + t.m[1][1] * p.y + t.m[1][2] * p.z + t.m[1][3];
  double zp = t.m[2][0] * p.x + t.m[2][1] * p.y + t.m[2][2] * p.z + t.m[2][3];
  double wp = t.m[3][0] * p.x + t.m[3][1] * p.y + t.m[3][2] * p.z + t.m[3][3];
  if (err) {
    double g3 = gamma_n(3), g31 = gamma_n(3.0) + 1.0;
    err->x = g31 * (fabs(t.m[0][0]) * pe.x + fabs(t.m[0][1]) * pe.y + fabs(t.m[0][2]) * pe.z) +
             (g3 * (fabs(t.m[0][0] * p.x) + fabs(t.m[0][1]) * p.y + fabs(t.m[0][2] * p.z + fabs(t.m[0][3]))));
    err->y = g31 * (fabs(t.m[1][0]) * pe.x + fabs(t.m[1][1]) * pe.y + fabs(t.m[1][2]) * pe.z) +
             (g3 * (fabs(t.m[1][0] * p.x) + fabs(t.m[1][1]) * p.y + fabs(t.m[1][2] * p.z + fabs(t.m[1][3]))));
    err->z = g31 * (fabs(t.m[2][0]) * pe.x + fabs(t.m[2][1]) * pe.y + fabs(t.m[2][2]) * pe.z) +
             (g3 * (fabs(t.m[2][0] * p.x) + fabs(t.m[2][1]) * p.y + fabs(t.m[2][2] * p.z + fabs(t.m[2][3]))));
  }
  V3 np = mk3(xp, yp, zp);
  if (wp == 1.0) return np;
  return np / wp;
}
GP_HD V3 xf_vector(const M4& t, V3 v) {  // transform.go:249-255
  return mk3(t.m[0][0] * v.x + t.m[0][1] * v.y + t.m[0][2] * v.z, t.m[1][0] * v.x + t.m[1][1] * v.y + t.m[1][2] * v.z,
             t.m[2][0] * v.x + t.m[2][1] * v.y + t.m[2][2] * v.z);
}
GP_HD V3 xf_vector_err(const M4& t, V3 v, V3* err) {  // transform.go:257-269
  double g3 = gamma_n(3);
  err->x = g3 * (fabs(t.m[0][0] * v.x) + fabs(t.m[0][1] * v.y) + fabs(t.m[0][2] * v.z));
  err->y = g3 * (fabs(t.m[1][0] * v.x) + fabs(t.m[1][1] * v.y) + fabs(t.m[1][2] * v.z));
  err->z = g3 * (fabs(t.m[2][0] * v.x) + fabs(t.m[2][1] * v.y) + fabs(t.m[2][2] * v.z));
  return xf_vector(t, v);
}
// TransformNormal (transform.go:271-277): transpose of the INVERSE matrix — pass the inverse here
GP_HD V3 xf_normal_inv(const M4& inv, V3 n) {
  return mk3(inv.m[0][0] * n.x + inv.m[1][0] * n.y + inv.m[2][0] * n.z, inv.m[0][1] * n.x + inv.m[1][1] * n.y + inv.m[2][1] * n.z,
             inv.m[0][2] * n.x + inv.m[1][2] * n.y + inv.m[2][2] * n.z);
}

struct Ray { V3 o, d; double tmax; };

// TransformRay (transform.go:279-300)
GP_HD Ray xf_ray(const M4& t, const Ray& r, V3* oerr, V3* derr) {
  V3 oe, de;
  V3 o = xf_point(t, r.o, mk3(0, 0, 0), &oe);
  V3 d = xf_vector_err(t, r.d, &de);
  double l2 = len2(d);
  if (l2 > 0) {
    double dt = dot(vabs(d), oe) / l2;
    o = o + d * dt;
  }
  if (oerr) *oerr = oe;
  if (derr) *derr = de;
  Ray out; out.o = o; out.d = d; out.tmax = r.tmax;
  return out;
}

// Bounds3.IntersectP (bounds.go:149-185): same if-structure (NaN semantics); the robustness factor
// 1 + 2*gamma(3) is exactly 1.0 and is multiplied for real.
GP_HD bool slab_test(double bx0, double by0, double bz0, double bx1, double by1, double bz1, V3 o, V3 invd, int nx, int ny, int nz,
                     double rtmax) {
  double g = 1 + 2 * gamma_n(3);
  double tMin = ((nx ? bx1 : bx0) - o.x) * invd.x;
  double tMax = ((nx ? bx0 : bx1) - o.x) * invd.x;
  double tyMin = ((ny ? by1 : by0) - o.y) * invd.y;
  double tyMax = ((ny ? by0 : by1) - o.y) * invd.y;
  tMax *= g;
  tyMax *= g;
  if (tMin > tyMax || tyMin > tMax) return false;
  if (tyMin > tMin) tMin = tyMin;
  if (tyMax < tMax) tMax = tyMax;
  double tzMin = ((nz ? bz1 : bz0) - o.z) * invd.z;
  double tzMax = ((nz ? bz0 : bz1) - o.z) * invd.z;
  tzMax *= g;
  if (tMin > tzMax || tzMin > tMax) return false;
  if (tzMin > tMin) tMin = tzMin;
  if (tzMax < tMax) tMax = tzMax;
  return tMin < rtmax && tMax > 0;
}

// ---- conservative float32 slab test for BVH NODES ----
// Node boxes are float32, rounded outward from the float64 union of their primitives' bounds.  Exact float64 decisions
// are made per primitive (own-bound test + shape test), so a node test only has to be a SUPERSET of the reference's
// float64 slab test (bounds.go:149-185) applied to any primitive bound inside the box: it may pass a box the exact test
// would reject, never the other way round.  Per axis, with s = sign(1/d):
//   dn = RD(s*b_near - s*o) <= s*(B_near - o),  lb = RD(dn * RD|1/d|)      (a lower bound of the exact tNear when dn >= 0)
//   df = RU(s*b_far  - s*o) >= s*(B_far  - o),  ub = RU(df * RU|1/d|)      (an upper bound of the exact tFar  when df >= 0)
// for every primitive bound B inside the node box b.  The two cases the magnitudes do not cover are harmless:
//   dn < 0: lb is merely <= 0; a non-positive tNear can only reject through "tNear > tFar" with tFar < 0, where the exact
//           test fails its own "tMax > 0" — or through "tNear >= r.TMax", which needs r.TMax <= 0: such rays are retired
//           before traversal (no shape test can return a hit with t <= 0);
//   df < 0: every primitive inside has an exact tFar < 0 on this axis and fails "tMax > 0" in the reference.
// NaNs (0*inf, inf-inf) compare false and therefore pass, as they do in the reference's comparisons.
struct RayF32 {
  float c_lo[3], c_hi[3];    // RD(-s*o), RU(-s*o)
  float ia_lo[3], ia_hi[3];  // RD|1/d|, RU|1/d|
  float sgn[3];              // s = -1 where 1/d < 0 (bounds.go:150: dirIsNeg), else +1
};
GP_D RayF32 ray_f32(V3 o, V3 invd) {
  RayF32 r;
  const double oo[3] = {o.x, o.y, o.z}, ii[3] = {invd.x, invd.y, invd.z};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const bool neg = ii[k] < 0;
    const double so = neg ? oo[k] : -oo[k];  // -s*o
    r.c_lo[k] = __double2float_rd(so);
    r.c_hi[k] = __double2float_ru(so);
    const double ia = fabs(ii[k]);
    r.ia_lo[k] = __double2float_rd(ia);
    r.ia_hi[k] = __double2float_ru(ia);
    r.sgn[k] = neg ? -1.f : 1.f;
  }
  return r;
}
GP_D bool slab_test_f32_maybe(float4 n0, float4 n1, const RayF32& r, int nx, int ny, int nz, float tmax_ub) {
  // near / far plane per axis (bounds.go:151-152: bounds[dirIsNeg] / bounds[1 - dirIsNeg])
  const float nxp = nx ? n1.x : n0.x, fxp = nx ? n0.x : n1.x;
  const float nyp = ny ? n1.y : n0.y, fyp = ny ? n0.y : n1.y;
  const float nzp = nz ? n1.z : n0.z, fzp = nz ? n0.z : n1.z;
  const float a0 = __fmul_rd(__fmaf_rd(nxp, r.sgn[0], r.c_lo[0]), r.ia_lo[0]);
  const float a1 = __fmul_rd(__fmaf_rd(nyp, r.sgn[1], r.c_lo[1]), r.ia_lo[1]);
  const float a2 = __fmul_rd(__fmaf_rd(nzp, r.sgn[2], r.c_lo[2]), r.ia_lo[2]);
  const float b0 = __fmul_ru(__fmaf_ru(fxp, r.sgn[0], r.c_hi[0]), r.ia_hi[0]);
  const float b1 = __fmul_ru(__fmaf_ru(fyp, r.sgn[1], r.c_hi[1]), r.ia_hi[1]);
  const float b2 = __fmul_ru(__fmaf_ru(fzp, r.sgn[2], r.c_hi[2]), r.ia_hi[2]);
  const float tnear = fmaxf(fmaxf(a0, a1), a2);  // fmaxf/fminf drop NaNs: an unknown axis only loosens the bracket
  const float tfar = fminf(fminf(b0, b1), b2);
  // reject only what the exact test is certain to reject: near > far on some axis pair, tMin >= r.TMax, or tMax <= 0
  return !(tnear > tfar) & !(tnear >= tmax_ub) & !(tfar <= 0.f);
}

// OffsetRayOrigin (ray.go:57-74)
GP_HD V3 offset_ray_origin(V3 p, V3 pError, V3 n, V3 w) {
  double d = dot(vabs(n), pError) * 1024.0;
  V3 off = n * d;
  if (dot(w, n) < 0) off = off * -1.0;
  V3 po = p + off;
  if (off.x > 0) po.x = next_up(po.x); else if (off.x < 0) po.x = next_down(po.x);
  if (off.y > 0) po.y = next_up(po.y); else if (off.y < 0) po.y = next_down(po.y);
  if (off.z > 0) po.z = next_up(po.z); else if (off.z < 0) po.z = next_down(po.z);
  return po;
}

}  // namespace gp
