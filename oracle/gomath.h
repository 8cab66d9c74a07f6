// ORACLE — test infrastructure only.  Nothing under go-pbrt_b200/ may include, link or call this.
//
// gomath.h: CPU restatement of the Go standard library `math` routines the reference's hot path
// reaches through pkg/math/math.go:21-144.  The Go toolchain (golang:1.11 per cloudbuild.yaml:12) is a
// third-party dependency absent from /root/reference; on amd64 its Sin/Cos/Tan/Atan/Atan2/Asin/Acos are
// the pure-Go Cephes-derived routines (src/math/{sin,tan,atan,atan2,asin}.go — "Cephes Math Library
// Release 2.8", S. Moshier), restated here from the published Cephes algorithm: 3-part Cody-Waite
// reduction by Pi/4 (DP1..DP3), degree-6 minimax polynomials (sincof/coscof), atan via P/Q rational on
// three ranges.  Sqrt/Floor/Ceil/Abs/Nextafter/Min/Max are IEEE-exact and restated directly.
// Pinned by the one reference test that touches this boundary: pkg/pbrt/transform_test.go:77-81
// (RotateY(90): cos(Pi/180*90) must be 6.123233995736757e-17, which glibc does not return).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace gomath {

constexpr double Pi = 3.14159265358979323846264338327950288419716939937510582097494459;
static const double Inf = INFINITY;

static inline uint64_t bits(double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; }
static inline double frombits(uint64_t u) { double x; std::memcpy(&x, &u, 8); return x; }

// math.Nextafter (src/math/nextafter.go)
static inline double Nextafter(double x, double y) {
  if (std::isnan(x) || std::isnan(y)) return NAN;
  if (x == y) return x;
  if (x == 0) return std::copysign(frombits(1), y);
  if ((y > x) == (x > 0)) return frombits(bits(x) + 1);
  return frombits(bits(x) - 1);
}
// pkg/math/math.go:122-128
static inline double NextFloatUp(double v) { return Nextafter(v, v + 1); }
static inline double NextFloatDown(double v) { return Nextafter(v, v - 1); }

// math.Min / math.Max (src/math/dim.go): NaN-propagating, signed-zero aware, ±Inf first
static inline double Min(double x, double y) {
  if ((std::isinf(x) && x < 0) || (std::isinf(y) && y < 0)) return -Inf;
  if (std::isnan(x) || std::isnan(y)) return NAN;
  if (x == 0 && x == y) return std::signbit(x) ? x : y;
  return x < y ? x : y;
}
static inline double Max(double x, double y) {
  if ((std::isinf(x) && x > 0) || (std::isinf(y) && y > 0)) return Inf;
  if (std::isnan(x) || std::isnan(y)) return NAN;
  if (x == 0 && x == y) return std::signbit(x) ? y : x;
  return x > y ? x : y;
}
// pkg/math/math.go:42-50
static inline double Clamp(double v, double lo, double hi) {
  if (v < lo) return lo;
  if (v > hi) return hi;
  return v;
}

// pkg/math/math.go:17-19,82-84 — MachineEpsilon is the smallest denormal
static inline double MachineEpsilon() { return NextFloatUp(0.0); }
static inline double OneMinusEpsilon() { return NextFloatDown(1.0); }
static inline double Gamma(double n) { double e = MachineEpsilon(); return (n * e) / (1 - n * e); }

// ---- Cephes sin/cos as in Go src/math/sin.go ----
static const double PI4A = 7.85398125648498535156e-1;   // 0x3fe921fb40000000
static const double PI4B = 3.77489470793079817668e-8;   // 0x3e64442d00000000
static const double PI4C = 2.69515142907905952645e-15;  // 0x3ce8469898cc5170
static const double M4PI = 1.273239544735162542821171882678754627704620361328125;  // Go 1.11 sin.go literal = 0x1.45f306dc9c882p+0
// (Go >= 1.12 writes 4/Pi, which rounds to ...883, and adds Payne-Hanek reduction for |x| >= 2^29; the
// reference pins golang:1.11 in cloudbuild.yaml:12, and hot-path arguments are all < 2*Pi.)

static const double sincof[6] = {
    1.58962301576546568060e-10, -2.50507477628578072866e-8, 2.75573136213857245213e-6,
    -1.98412698295895385996e-4, 8.33333333332211858878e-3,  -1.66666666666666307295e-1};
static const double coscof[6] = {
    -1.13585365213876817300e-11, 2.08757008419747316778e-9, -2.75573141792967388112e-7,
    2.48015872888517045348e-5,   -1.38888888888730564116e-3, 4.16666666666665929218e-2};

static inline double Cos(double x) {
  if (std::isnan(x) || std::isinf(x)) return NAN;
  bool sign = false;
  if (x < 0) x = -x;
  uint64_t j = (uint64_t)(x * M4PI);
  double y = (double)j;
  if (j & 1) { j++; y++; }
  j &= 7;
  if (j > 3) { j -= 4; sign = !sign; }
  if (j > 1) sign = !sign;
  double z = ((x - y * PI4A) - y * PI4B) - y * PI4C;
  double zz = z * z;
  if (j == 1 || j == 2)
    y = z + z * zz * ((((((sincof[0] * zz) + sincof[1]) * zz + sincof[2]) * zz + sincof[3]) * zz + sincof[4]) * zz + sincof[5]);
  else
    y = 1.0 - 0.5 * zz + zz * zz * ((((((coscof[0] * zz) + coscof[1]) * zz + coscof[2]) * zz + coscof[3]) * zz + coscof[4]) * zz + coscof[5]);
  return sign ? -y : y;
}

static inline double Sin(double x) {
  if (x == 0 || std::isnan(x)) return x;
  if (std::isinf(x)) return NAN;
  bool sign = false;
  if (x < 0) { x = -x; sign = true; }
  uint64_t j = (uint64_t)(x * M4PI);
  double y = (double)j;
  if (j & 1) { j++; y++; }
  j &= 7;
  if (j > 3) { sign = !sign; j -= 4; }
  double z = ((x - y * PI4A) - y * PI4B) - y * PI4C;
  double zz = z * z;
  if (j == 1 || j == 2)
    y = 1.0 - 0.5 * zz + zz * zz * ((((((coscof[0] * zz) + coscof[1]) * zz + coscof[2]) * zz + coscof[3]) * zz + coscof[4]) * zz + coscof[5]);
  else
    y = z + z * zz * ((((((sincof[0] * zz) + sincof[1]) * zz + sincof[2]) * zz + sincof[3]) * zz + sincof[4]) * zz + sincof[5]);
  return sign ? -y : y;
}

// ---- Cephes tan as in Go src/math/tan.go (host-side only: Perspective, transform.go:500) ----
static inline double Tan(double x) {
  static const double P[3] = {-1.30936939181383777646e4, 1.15351664838587416140e6, -1.79565251976484877988e7};
  static const double Q[5] = {1.0, 1.36812963470692954678e4, -1.32089234440210967447e6, 2.50083801823357915839e7,
                              -5.38695755929454629881e7};
  if (x == 0 || std::isnan(x)) return x;
  if (std::isinf(x)) return NAN;
  bool sign = false;
  if (x < 0) { x = -x; sign = true; }
  uint64_t j = (uint64_t)(x * M4PI);
  double y = (double)j;
  if (j & 1) { j++; y++; }
  double z = ((x - y * PI4A) - y * PI4B) - y * PI4C;
  double zz = z * z;
  if (zz > 1e-14)
    y = z + z * (zz * (((P[0] * zz) + P[1]) * zz + P[2]) / ((((zz + Q[1]) * zz + Q[2]) * zz + Q[3]) * zz + Q[4]));
  else
    y = z;
  if ((j & 2) == 2) y = -1 / y;
  return sign ? -y : y;
}

// ---- Cephes atan as in Go src/math/atan.go ----
static inline double xatan(double x) {
  const double P0 = -8.750608600031904122785e-01, P1 = -1.615753718733365076637e+01, P2 = -7.500855792314704667340e+01,
               P3 = -1.228866684490136173410e+02, P4 = -6.485021904942025371773e+01;
  const double Q0 = +2.485846490142306297962e+01, Q1 = +1.650270098316988542046e+02, Q2 = +4.328810604912902668951e+02,
               Q3 = +4.853903996359136964868e+02, Q4 = +1.945506571482613964425e+02;
  double z = x * x;
  z = z * ((((P0 * z + P1) * z + P2) * z + P3) * z + P4) / (((((z + Q0) * z + Q1) * z + Q2) * z + Q3) * z + Q4);
  z = x * z + x;
  return z;
}
static inline double satan(double x) {
  const double Morebits = 6.123233995736765886130e-17;
  const double Tan3pio8 = 2.41421356237309504880;
  if (x <= 0.66) return xatan(x);
  if (x > Tan3pio8) return Pi / 2 - xatan(1 / x) + Morebits;
  return Pi / 4 + xatan((x - 1) / (x + 1)) + 0.5 * Morebits;
}
static inline double Atan(double x) {
  if (x == 0) return x;
  if (x > 0) return satan(x);
  return -satan(-x);
}
// src/math/atan2.go
static inline double Atan2(double y, double x) {
  if (std::isnan(y) || std::isnan(x)) return NAN;
  if (y == 0) {
    if (x >= 0 && !std::signbit(x)) return std::copysign(0.0, y);
    return std::copysign(Pi, y);
  }
  if (x == 0) return std::copysign(Pi / 2, y);
  if (std::isinf(x)) {
    if (x > 0) return std::isinf(y) ? std::copysign(Pi / 4, y) : std::copysign(0.0, y);
    return std::isinf(y) ? std::copysign(3 * Pi / 4, y) : std::copysign(Pi, y);
  }
  if (std::isinf(y)) return std::copysign(Pi / 2, y);
  double q = Atan(y / x);
  if (x < 0) return q <= 0 ? q + Pi : q - Pi;
  return q;
}
// src/math/asin.go
static inline double Asin(double x) {
  if (x == 0) return x;
  bool sign = false;
  if (x < 0) { x = -x; sign = true; }
  if (x > 1) return NAN;
  double temp = std::sqrt(1 - x * x);
  if (x > 0.7) temp = Pi / 2 - satan(temp / x);
  else temp = satan(x / temp);
  return sign ? -temp : temp;
}
static inline double Acos(double x) { return Pi / 2 - Asin(x); }

}  // namespace gomath
