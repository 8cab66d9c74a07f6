// ORACLE — test infrastructure only.  Nothing under go-pbrt_b200/ may include, link or call this.
//
// oracle_render.h: CPU restatement of the reference's sampler / camera / BSDF / light / Path.Li / DirectLighting.Li /
// film code.  The reference holds no golden value for any of this and
// cannot be run here; this file follows the cited lines character by character, quirks included.  PARITY PIN: independent
// plain-Python restatements of the same Go source — per function (tests/golden/make_shading_kats.py) and as whole films
// (tests/golden/make_*_golden.py: configs 1 and 2, all shapes / materials / lights / samplers, both integrators, FAST mode) — which
// this file reproduces bit for bit with identical ray counts (tests/test_shading_kats.py, tests/test_*_golden.py).
#pragma once
#include <mutex>
#include <thread>

#include "oracle_core.h"

namespace oracle {

struct P2 { double x = 0, y = 0; };
struct RGB {
  double c[3] = {0, 0, 0};
  RGB() {}
  explicit RGB(double v) { c[0] = c[1] = c[2] = v; }
  RGB(double r, double g, double b) { c[0] = r; c[1] = g; c[2] = b; }
};
// pkg/pbrt/spectrum.go
static inline RGB smul(RGB a, RGB b) { return RGB(a.c[0] * b.c[0], a.c[1] * b.c[1], a.c[2] * b.c[2]); }
static inline RGB smuls(RGB a, double s) { return RGB(a.c[0] * s, a.c[1] * s, a.c[2] * s); }
static inline RGB sdivs(RGB a, double s) { return RGB(a.c[0] / s, a.c[1] / s, a.c[2] / s); }
static inline RGB sadd(RGB a, RGB b) { return RGB(a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2]); }
static inline bool sblack(RGB a) { return !(a.c[0] != 0.0) && !(a.c[1] != 0.0) && !(a.c[2] != 0.0); }  // spectrum.go:209-216
static inline bool snan(RGB a) { return std::isnan(a.c[0]) || std::isnan(a.c[1]) || std::isnan(a.c[2]); }
static inline double smax(RGB a) { return gm::Max(gm::Max(a.c[0], a.c[1]), a.c[2]); }  // spectrum.go:185-191
static inline RGB sclamp(RGB a, double lo, double hi) { return RGB(gm::Clamp(a.c[0], lo, hi), gm::Clamp(a.c[1], lo, hi), gm::Clamp(a.c[2], lo, hi)); }

// ---------------------------------------------------------------- pkg/pbrt/rng.go (NOT standard PCG32)
struct Rng {
  uint64_t state = 0x853c49e6748fea9bULL, inc = 0xda3e39cb94b95bdbULL;
  uint32_t u32() {  // rng.go:36-42
    uint64_t old = state;
    state = old * 0x5851f42d4c957f2dULL + inc;
    uint32_t xs = (uint32_t)(((old >> 18) ^ old) >> 27);
    uint32_t rot = (uint32_t)(old >> 59);
    return (xs >> rot) | (xs << ((rot + 1u) & 31));
  }
  void set_sequence(uint64_t seed) {  // rng.go:28-34
    state = 0;
    inc = (seed << 1) | 1;
    u32();
    state += 0x853c49e6748fea9bULL;
    u32();
  }
  uint32_t u32b(uint32_t b) {  // rng.go:44-53
    uint32_t threshold = (~b + 1u) % b;
    for (;;) {
      uint32_t r = u32();
      if (r >= threshold) return r % b;
    }
  }
  double uniform() { return gm::Min(gm::OneMinusEpsilon(), (double)u32() * 2.3283064365386963e-10); }  // rng.go:55-57
};

// Kensler's stateless permutation, used only by the FAST (counter-based) mode, which is this backend's own
// definition (SURVEY §7 "fast"); the Go host needs the matching Sampler.
static inline uint32_t kensler_permute(uint32_t i, uint32_t l, uint32_t p) {
  uint32_t w = l - 1;
  w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
  do {
    i ^= p; i *= 0xe170893d; i ^= p >> 16; i ^= (i & w) >> 4; i ^= p >> 8; i *= 0x0929eb3f; i ^= p >> 23;
    i ^= (i & w) >> 1; i *= 1 | p >> 27; i *= 0x6935fa69; i ^= (i & w) >> 11; i *= 0x74dcb303; i ^= (i & w) >> 2;
    i *= 0x9e501cc3; i ^= (i & w) >> 2; i *= 0xc860a3df; i &= w; i ^= i >> 5;
  } while (i >= l);
  return (i + p) % l;
}
static inline uint32_t hash_u32(uint64_t a, uint32_t b) {
  uint64_t x = a * 0x9E3779B97F4A7C15ULL + (uint64_t)b * 0xD1B54A32D192ED03ULL + 0x632BE59BD9B4E019ULL;
  x ^= x >> 32; x *= 0xD6E8FEB86659FD93ULL; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ULL; x ^= x >> 32;
  return (uint32_t)x;
}

// ---------------------------------------------------------------- pkg/sampler
struct Sampler {
  gopbrt_sampler cfg;
  int spp = 0;
  std::vector<std::vector<double>> s1d;  // pixel.go:28-31
  std::vector<std::vector<P2>> s2d;
  int cur1 = 0, cur2 = 0;  // current1DDimension / current2DDimension
  int idx = 0;             // currentPixelSampleIndex
  Rng rng;
  // FAST mode
  uint64_t fast_pixel = 0;

  void init(const gopbrt_sampler& c) {
    cfg = c;
    spp = c.kind == GOPBRT_SAMPLER_STRATIFIED ? c.x_samples * c.y_samples : c.x_samples;
    int nd = c.kind == GOPBRT_SAMPLER_STRATIFIED ? c.n_sampled_dimensions : 0;
    s1d.assign(nd, std::vector<double>(spp, 0.0));
    s2d.assign(nd, std::vector<P2>(spp));
  }
  void clone_seed(uint64_t seed) { rng = Rng(); rng.set_sequence(seed); }  // pixel.go:34-42, random.go:35-39

  // stratified.go:21-48 (array requests are never made by Path: sampler.go:43-51 unused)
  void start_pixel() {
    if (cfg.kind == GOPBRT_SAMPLER_STRATIFIED && cfg.mode == GOPBRT_MODE_STRICT) {
      int n = spp;
      for (size_t i = 0; i < s1d.size(); i++) {
        // StratifiedSample1D (sampling.go:101-110)
        double inv = 1.0 / (double)n;
        for (int k = 0; k < n; k++) {
          double delta = 0.5;
          if (cfg.jitter) delta = rng.uniform();
          s1d[i][k] = gm::Min(((double)k + delta) * inv, gm::OneMinusEpsilon());
        }
        // ShuffleSamples1D (sampling.go:129-136)
        for (int k = 0; k < n; k++) {
          int other = k + (int)rng.u32b((uint32_t)(n - k));
          std::swap(s1d[i][k], s1d[i][other]);
        }
      }
      for (size_t i = 0; i < s2d.size(); i++) {
        // StratifiedSample2D (sampling.go:112-127): writes into a COPY (`s := samp[..]`), so the table stays all
        // zeros; jitter still consumes two RNG draws per sample.
        for (int y = 0; y < cfg.y_samples; y++)
          for (int x = 0; x < cfg.x_samples; x++)
            if (cfg.jitter) { rng.uniform(); rng.uniform(); }
        // ShuffleSamples2D (sampling.go:138-146)
        for (int k = 0; k < n; k++) {
          int other = k + (int)rng.u32b((uint32_t)(n - k));
          std::swap(s2d[i][k], s2d[i][other]);
        }
      }
    }
    idx = 0;  // sampler.go:21-27
  }
  bool start_next_sample() {  // pixel.go:48-52 + sampler.go:29-34 (increments FIRST: spp-1 samples)
    cur1 = 0;
    cur2 = 0;
    idx += 1;
    if (idx < spp && cfg.mode == GOPBRT_MODE_FAST) rng.set_sequence(fast_pixel * (uint64_t)spp + (uint64_t)idx);
    return idx < spp;
  }
  double get1d() {  // pixel.go:60-69, random.go:21-23
    if (cfg.mode == GOPBRT_MODE_FAST) {
      if (cfg.kind == GOPBRT_SAMPLER_STRATIFIED && cur1 < cfg.n_sampled_dimensions) {
        uint32_t j = kensler_permute((uint32_t)idx, (uint32_t)spp, hash_u32(fast_pixel, (uint32_t)cur1));
        cur1++;
        double delta = 0.5;
        if (cfg.jitter) delta = rng.uniform();
        return gm::Min(((double)j + delta) * (1.0 / (double)spp), gm::OneMinusEpsilon());
      }
      return rng.uniform();
    }
    if (cur1 < (int)s1d.size()) return s1d[cur1++][idx];
    return rng.uniform();
  }
  P2 get2d() {  // pixel.go:71-80, random.go:25-27
    if (cfg.mode == GOPBRT_MODE_FAST) {
      if (cfg.kind == GOPBRT_SAMPLER_STRATIFIED && cur2 < cfg.n_sampled_dimensions) { cur2++; return P2{}; }
      P2 p; p.x = rng.uniform(); p.y = rng.uniform(); return p;
    }
    if (cur2 < (int)s2d.size()) return s2d[cur2++][idx];
    P2 p;
    p.x = rng.uniform();
    p.y = rng.uniform();
    return p;
  }
};

// ---------------------------------------------------------------- pkg/pbrt/sampling.go warps
static inline P2 concentric_sample_disk(P2 u) {  // sampling.go:173-192
  double ox = u.x * 2.0 - 1, oy = u.y * 2.0 - 1;
  if (ox == 0 && oy == 0) return P2{};
  double theta, r;
  double piOver4 = gm::Pi / 4.0, piOver2 = gm::Pi / 2.0;
  if (std::fabs(ox) > std::fabs(oy)) {
    r = ox;
    theta = piOver4 * (oy / ox);
  } else {
    r = oy;
    theta = piOver2 - piOver4 * (ox / oy);
  }
  return P2{gm::Cos(theta) * r, gm::Sin(theta) * r};
}
static inline V3 cosine_sample_hemisphere(P2 u) {  // sampling.go:194-198
  P2 d = concentric_sample_disk(u);
  double z = std::sqrt(gm::Max(0.0, 1.0 - d.x * d.x - d.y * d.y));
  return V3{d.x, d.y, z};
}
static inline V3 uniform_sample_sphere(P2 u) {  // sampling.go:158-163
  double z = 1.0 - 2.0 * u.x;
  double r = std::sqrt(gm::Max(0, 1 - z * z));
  double phi = 2 * gm::Pi * u.y;
  return V3{r * gm::Cos(phi), r * gm::Sin(phi), z};
}
static inline double uniform_cone_pdf(double cosThetaMax) { return 1.0 / (2.0 * gm::Pi * (1.0 - cosThetaMax)); }  // sampling.go:169-171
static inline double power_heuristic(double fPdf, double gPdf) {  // sampling.go:208-212 with nf = ng = 1
  double f = 1.0 * fPdf, g = 1.0 * gPdf;
  return (f * f) / (f * f + g * g);
}

// ---------------------------------------------------------------- pkg/pbrt/ray.go:57-74, interaction.go:68-102
static inline V3 offset_ray_origin(V3 p, V3 pError, V3 n, V3 w) {
  double d = dot(vabs(n), pError) * 1024.0;
  V3 offset = muls(n, d);
  if (dot(w, n) < 0) offset = muls(offset, -1);
  V3 po = add(p, offset);
  for (int i = 0; i < 3; i++) {
    if (offset[i] > 0) po.set(i, gm::NextFloatUp(po[i]));
    else if (offset[i] < 0) po.set(i, gm::NextFloatDown(po[i]));
  }
  return po;
}
struct Intr { V3 p, perr, n; };  // the fields of pbrt.interaction the spawn rules read
static inline Ray spawn_ray(const Intr& i, V3 d, double time) {  // interaction.go:68-77
  return Ray{offset_ray_origin(i.p, i.perr, i.n, d), d, gm::Inf, time};
}
static inline Ray spawn_ray_to(const Intr& i, const Intr& to, double time) {  // interaction.go:91-102
  V3 origin = offset_ray_origin(i.p, i.perr, i.n, sub(to.p, i.p));
  V3 target = offset_ray_origin(to.p, to.perr, to.n, sub(origin, to.p));
  V3 d = sub(target, origin);
  return Ray{i.p, d, 1 - 0.0001, time};
}

// ---------------------------------------------------------------- textures (texture.go, checkerboard.go)
static inline RGB tex_eval(const Scene& sc, int id, const Hit& h) {
  for (int guard = 0; guard < 64; guard++) {
    const gopbrt_texture& t = sc.textures[id];
    if (t.kind == GOPBRT_TEX_CONSTANT) return RGB(t.rgb[0], t.rgb[1], t.rgb[2]);
    double s, tt;
    if (t.mapping == GOPBRT_MAP_PLANAR) {  // texture.go:42-46
      s = t.ds + dot(h.p, V3{t.vs[0], t.vs[1], t.vs[2]});
      tt = t.dt + dot(h.p, V3{t.vt[0], t.vt[1], t.vt[2]});
    } else {  // texture.go:22-26
      s = t.su * h.u + t.du;
      tt = t.sv * h.v + t.dv;
    }
    // checkerboard.go:30-36: int(floor(s)+floor(t)) % 2 == 0 (Go % keeps the sign; -1 % 2 = -1 != 0)
    int64_t k = (int64_t)(std::floor(s) + std::floor(tt));
    id = (k % 2 == 0) ? t.tex1 : t.tex2;
  }
  return RGB(0);
}

// ---------------------------------------------------------------- BSDF (reflection.go)
enum { BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16, BSDF_ALL = 31 };
enum { BX_LAMBERT = 0, BX_OREN_NAYAR = 1, BX_SPEC_REFL_NOOP = 2, BX_FRESNEL_SPECULAR = 3, BX_SPEC_REFL_DIELECTRIC = 4, BX_SPEC_TRANS = 5 };
struct BxDF { int kind, type; RGB r, t; double a = 0, b = 0, etaA = 1, etaB = 1; };
struct BSDF {
  double eta = 1;
  V3 ns, ng, ss, ts;
  int n = 0;
  BxDF bx[2];
};
static inline bool matches(int t, int flags) { return (t & flags) == t; }  // reflection.go:301-303
// NewOrenNayar (reflection.go:616-626) (b = 0.45 s2 / (s2 * 0.09))
static inline BxDF make_oren_nayar(RGB r, double sigma_deg) {
  BxDF x;
  x.kind = BX_OREN_NAYAR;
  x.type = BSDF_REFLECTION | BSDF_DIFFUSE;
  x.r = r;
  double s = gm::Pi / 180.0 * sigma_deg;
  double s2 = s * s;
  x.a = 1.0 - (s2 / (2.0 * (s2 + 0.33)));
  x.b = 0.45 * s2 / (s2 * 0.09);
  return x;
}
static inline double fr_dielectric(double cosThetaI, double etaI, double etaT) {  // reflection.go:21-42
  cosThetaI = gm::Clamp(cosThetaI, -1, 1);
  bool entering = cosThetaI > 0;
  if (!entering) { std::swap(etaI, etaT); cosThetaI = std::fabs(cosThetaI); }
  double sinThetaI = std::sqrt(gm::Max(0, 1 - cosThetaI * cosThetaI));
  double sinThetaT = etaI / etaT * sinThetaI;
  if (sinThetaT >= 1) return 1;
  double cosThetaT = std::sqrt(gm::Max(0, 1 - sinThetaT * sinThetaT));
  double Rparl = ((etaT * cosThetaI) - (etaI * cosThetaT)) / ((etaT * cosThetaI) + (etaI * cosThetaT));
  double Rperp = ((etaI * cosThetaI) - (etaT * cosThetaT)) / ((etaI * cosThetaI) + (etaT * cosThetaT));
  return (Rparl * Rparl + Rperp * Rperp) / 2;
}
static inline double sin2theta(V3 w) { return gm::Max(0, 1 - w.z * w.z); }
static inline double sintheta(V3 w) { return std::sqrt(sin2theta(w)); }
static inline double cosphi(V3 w) { double s = sintheta(w); return s == 0 ? 1 : gm::Clamp(w.x / s, -1, 1); }  // reflection.go:76-92
static inline double sinphi(V3 w) { double s = sintheta(w); return s == 0 ? 0 : gm::Clamp(w.y / s, -1, 1); }

static inline RGB bxdf_f(const BxDF& b, V3 wo, V3 wi) {
  const double invPi = 1.0 / gm::Pi;
  switch (b.kind) {
    case BX_LAMBERT: return smuls(b.r, invPi);  // reflection.go:589-591
    case BX_OREN_NAYAR: {                        // reflection.go:628-652 (tanBeta uses wo in both branches)
      double sinThetaI = sintheta(wi), sinThetaO = sintheta(wo);
      double maxCos = 0.0;
      if (sinThetaI > 1e-4 && sinThetaO > 1e-4) {
        double sinPhiI = sinphi(wi), cosPhiI = cosphi(wi), sinPhiO = sinphi(wo), cosPhiO = cosphi(wo);
        double dCos = cosPhiI * cosPhiO + sinPhiI * sinPhiO;
        maxCos = gm::Max(0.0, dCos);
      }
      double sinAlpha, tanBeta;
      if (std::fabs(wi.z) > std::fabs(wo.z)) { sinAlpha = sinThetaO; tanBeta = sinThetaO / std::fabs(wo.z); }
      else { sinAlpha = sinThetaI; tanBeta = sinThetaO / std::fabs(wo.z); }
      return smuls(b.r, invPi * (b.a + b.b * maxCos * sinAlpha * tanBeta));
    }
    default: return RGB(0);  // SpecularReflection.F / FresnelSpecular.F (reflection.go:479-481,553-555)
  }
}
static inline double bxdf_pdf(const BxDF& b, V3 wo, V3 wi) {
  if (b.kind == BX_LAMBERT || b.kind == BX_OREN_NAYAR) {  // reflection.go:343-348
    if (wo.z * wi.z > 0) return std::fabs(wi.z) * (1.0 / gm::Pi);
    return 0;
  }
  return 0;
}
// returns false for the "Refract failed" return (pdf 0)
static inline void bxdf_sample_f(const BxDF& b, V3 wo, P2 u, RGB* f, V3* wi, double* pdf, int* sampled) {
  switch (b.kind) {
    case BX_LAMBERT:
    case BX_OREN_NAYAR: {  // sampleF (reflection.go:305-314): sampledType = 0
      V3 w = cosine_sample_hemisphere(u);
      if (wo.z < 0) w.z *= -1;
      *wi = w;
      *pdf = bxdf_pdf(b, wo, w);
      *f = bxdf_f(b, wo, w);
      *sampled = 0;
      return;
    }
    case BX_SPEC_REFL_NOOP: {  // reflection.go:557-562 with FresnelNoOp (:379-384)
      V3 w{-wo.x, -wo.y, wo.z};
      *wi = w;
      *pdf = 1.0;
      *f = sdivs(smul(RGB(1.0), b.r), std::fabs(w.z));
      *sampled = 0;
      return;
    }
    case BX_FRESNEL_SPECULAR: {  // reflection.go:482-523
      double F = fr_dielectric(wo.z, b.etaA, b.etaB);
      if (u.x < F) {
        V3 w{-wo.x, -wo.y, wo.z};
        *wi = w;
        *f = sdivs(smuls(b.r, F), std::fabs(w.z));
        *pdf = F;
        *sampled = BSDF_SPECULAR | BSDF_REFLECTION;
        return;
      }
      bool entering = wo.z > 0;
      double etaI = entering ? b.etaA : b.etaB, etaT = entering ? b.etaB : b.etaA;
      // Refract(wo, FaceForward((0,0,1), wo), etaI/etaT) (reflection.go:106-118)
      V3 n = faceforward(V3{0, 0, 1}, wo);
      double eta = etaI / etaT;
      double cosThetaI = dot(n, wo);
      double sin2ThetaI = gm::Max(0, 1 - cosThetaI * cosThetaI);
      double sin2ThetaT = eta * eta * sin2ThetaI;
      if (sin2ThetaT >= 1) { *f = RGB(0); *wi = V3{}; *pdf = 0; *sampled = 0; return; }
      double cosThetaT = std::sqrt(1 - sin2ThetaT);
      V3 w = add(muls(wo, -eta), muls(n, eta * cosThetaI - cosThetaT));
      RGB ft = smuls(b.t, 1 - F);
      ft = smuls(ft, (etaI * etaI) / (etaT / etaT));  // mode == Radiance; (etaT/etaT) sic
      *wi = w;
      *f = sdivs(ft, std::fabs(w.z));
      *pdf = 1 - F;
      *sampled = BSDF_SPECULAR | BSDF_TRANSMISSION;
      return;
    }
    case BX_SPEC_REFL_DIELECTRIC: {  // SpecularReflection with FresnelDielectric (reflection.go:557-562, :398-403)
      V3 w{-wo.x, -wo.y, wo.z};
      *wi = w;
      *pdf = 1.0;
      *f = sdivs(smul(RGB(fr_dielectric(w.z, b.etaA, b.etaB)), b.r), std::fabs(w.z));
      *sampled = 0;
      return;
    }
    case BX_SPEC_TRANS: {  // SpecularTransmission.SampleF (reflection.go:428-451), mode == Radiance
      bool entering = wo.z > 0;
      double etaI = entering ? b.etaA : b.etaB, etaT = entering ? b.etaB : b.etaA;
      V3 n = faceforward(V3{0, 0, 1}, wo);
      double eta = etaI / etaT;
      double cosThetaI = dot(n, wo);
      double sin2ThetaI = gm::Max(0, 1 - cosThetaI * cosThetaI);
      double sin2ThetaT = eta * eta * sin2ThetaI;
      if (sin2ThetaT >= 1) { *f = RGB(0); *wi = V3{}; *pdf = 0; *sampled = 0; return; }
      double cosThetaT = std::sqrt(1 - sin2ThetaT);
      V3 w = add(muls(wo, -eta), muls(n, eta * cosThetaI - cosThetaT));
      double F = fr_dielectric(w.z, b.etaA, b.etaB);
      RGB ft = smul(b.t, RGB(1.0 - F));
      ft = smuls(ft, (etaI * etaI) / (etaT * etaT));
      *wi = w;
      *f = sdivs(ft, std::fabs(w.z));
      *pdf = 1;
      *sampled = 0;  // sic (reflection.go:451)
      return;
    }
  }
}

static inline V3 to_local(const BSDF& b, V3 v) { return V3{dot(v, b.ss), dot(v, b.ts), dot(v, b.ns)}; }  // reflection.go:142-144
static inline V3 to_world(const BSDF& b, V3 v) {  // reflection.go:146-152
  return V3{b.ss.x * v.x + b.ts.x * v.y + b.ns.x * v.z, b.ss.y * v.x + b.ts.y * v.y + b.ns.y * v.z,
            b.ss.z * v.x + b.ts.z * v.y + b.ns.z * v.z};
}
static inline int bsdf_num_components(const BSDF& b, int flags) {  // reflection.go:154-162
  int n = 0;
  for (int i = 0; i < b.n; i++) if (matches(b.bx[i].type, flags)) n++;
  return n;
}
static inline RGB bsdf_f(const BSDF& b, V3 woW, V3 wiW, int flags) {  // reflection.go:164-181
  V3 wi = to_local(b, wiW), wo = to_local(b, woW);
  if (wo.z == 0.0) return RGB(0);
  bool reflect = dot(wiW, b.ng) * dot(woW, b.ng) > 0;
  RGB f(0);
  for (int i = 0; i < b.n; i++)
    if (matches(b.bx[i].type, flags) &&
        ((reflect && (b.bx[i].type & BSDF_REFLECTION) > 0) || (!reflect && (b.bx[i].type & BSDF_TRANSMISSION) > 0)))
      f = sadd(f, bxdf_f(b.bx[i], wo, wi));
  return f;
}
static inline double bsdf_pdf(const BSDF& b, V3 woW, V3 wiW, int flags) {  // reflection.go:255-278
  if (b.n == 0) return 0;
  V3 wo = to_local(b, woW), wi = to_local(b, wiW);
  if (wo.z == 0) return 0;
  double pdf = 0;
  int m = 0;
  for (int i = 0; i < b.n; i++)
    if (matches(b.bx[i].type, flags)) { m++; pdf += bxdf_pdf(b.bx[i], wo, wi); }
  if (m <= 0) return 0;
  return pdf / (double)m;
}
// reflection.go:183-253 — returns the LOCAL wi
static inline void bsdf_sample_f(const BSDF& b, V3 woWorld, P2 u, int type, RGB* f, V3* wi, double* pdf, int* sampled) {
  *f = RGB(0); *wi = V3{}; *pdf = 0; *sampled = 0;
  int m = bsdf_num_components(b, type);
  if (m == 0) return;
  double comp = gm::Min(std::floor(u.x * (double)m), (double)m - 1);
  double count = comp;
  int chosen = -1;
  for (int i = 0; i < b.n; i++)
    if (matches(b.bx[i].type, type)) {
      if (count == 0) { chosen = i; break; }
      count--;
    }
  P2 ur{gm::Min(u.x * (double)m - comp, gm::OneMinusEpsilon()), u.y};
  V3 wo = to_local(b, woWorld);
  if (wo.z == 0.0) return;
  RGB ff; V3 w; double p = 0; int st = 0;
  if (chosen < 0) return;
  bxdf_sample_f(b.bx[chosen], wo, ur, &ff, &w, &p, &st);
  if (p == 0.0) return;
  V3 wiWorld = to_world(b, w);
  if (((b.bx[chosen].type & BSDF_SPECULAR) <= 0) && m > 1)
    for (int i = 0; i < b.n; i++)
      if (i != chosen && matches(b.bx[i].type, type)) p += bxdf_pdf(b.bx[i], wo, w);
  if (m > 1) p /= (double)m;
  if (((b.bx[chosen].type & BSDF_SPECULAR) == 0) && m > 1) {
    bool reflect = dot(wiWorld, b.ng) * dot(woWorld, b.ng) > 0;
    ff = RGB(0);
    for (int i = 0; i < b.n; i++)
      if (matches(b.bx[i].type, type) &&
          ((reflect && (b.bx[i].type & BSDF_REFLECTION) > 0) || (!reflect && (b.bx[i].type & BSDF_TRANSMISSION) > 0)))
        ff = sadd(ff, bxdf_f(b.bx[i], wo, w));
  }
  *f = ff; *wi = w; *pdf = p; *sampled = st;
}

// debug aid (ORACLE_TRACE=<file>, single-threaded renders only): one line per scene.Intersect / EstimateDirect event
static FILE* g_trace = nullptr;

struct RenderStats {
  uint64_t camera_rays = 0, closest_rays = 0, shadow_rays = 0, dead_mis_rays = 0;
  uint64_t nodes = 0, prims = 0, snodes = 0, sprims = 0;
  uint64_t radiance_gt10 = 0, nan_samples = 0, unsupported_material = 0;
  void add(const RenderStats& o) {
    camera_rays += o.camera_rays; closest_rays += o.closest_rays; shadow_rays += o.shadow_rays; dead_mis_rays += o.dead_mis_rays;
    nodes += o.nodes; prims += o.prims; snodes += o.snodes; sprims += o.sprims;
    radiance_gt10 += o.radiance_gt10; nan_samples += o.nan_samples; unsupported_material += o.unsupported_material;
  }
};

// Material.ComputeScatteringFunctions (matte.go:21-37, mirror.go:21-32, glass.go:27-75) + NewBSDF (reflection.go:128-140)
static inline bool compute_scattering(const Scene& sc, const Hit& h, BSDF* b, RenderStats* st, bool allowMultipleLobes = true) {
  int mi = sc.prims[h.prim].material;
  if (mi < 0) { st->unsupported_material++; return false; }  // primitive.go:73-75 panics
  const gopbrt_material& m = sc.materials[mi];
  b->ns = h.ns;
  b->ng = h.n;
  b->ss = normalized(h.sdpdu);
  b->ts = cross(b->ns, b->ss);
  b->n = 0;
  b->eta = 1.0;
  switch (m.kind) {
    case GOPBRT_MAT_MATTE: {
      RGB r = sclamp(tex_eval(sc, m.tex_a, h), 0, gm::Inf);
      double sig = gm::Clamp(m.sigma, 0, 90);
      if (!sblack(r)) {
        BxDF x;
        x.type = BSDF_REFLECTION | BSDF_DIFFUSE;
        x.r = r;
        if (sig == 0) x.kind = BX_LAMBERT;
        else x = make_oren_nayar(r, sig);
        b->bx[b->n++] = x;
      }
      return true;
    }
    case GOPBRT_MAT_MIRROR: {
      RGB r = sclamp(tex_eval(sc, m.tex_a, h), 0.0, gm::Inf);
      if (!sblack(r)) {
        BxDF x;
        x.kind = BX_SPEC_REFL_NOOP;
        x.type = BSDF_REFLECTION | BSDF_DIFFUSE;  // sic: typed Reflection|Diffuse (reflection.go:540)
        x.r = r;
        b->bx[b->n++] = x;
      }
      return true;
    }
    case GOPBRT_MAT_GLASS: {
      b->eta = m.eta;
      RGB R = sclamp(tex_eval(sc, m.tex_a, h), 0, 1), T = sclamp(tex_eval(sc, m.tex_b, h), 0, 1);
      if (sblack(R) && sblack(T)) return true;
      bool isSpecular = m.u_rough == 0 && m.v_rough == 0;
      if (!isSpecular) { st->unsupported_material++; return false; }  // microfacet branch panics in the reference
      if (allowMultipleLobes) {
        BxDF x;
        x.kind = BX_FRESNEL_SPECULAR;
        x.type = BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_SPECULAR;
        x.r = R; x.t = T; x.etaA = 1.0; x.etaB = m.eta;
        b->bx[b->n++] = x;
        return true;
      }
      // glass.go:57-72: separate lobes.  SpecularReflection is typed Reflection|Diffuse (reflection.go:540)
      if (!sblack(R)) {
        BxDF x;
        x.kind = BX_SPEC_REFL_DIELECTRIC;
        x.type = BSDF_REFLECTION | BSDF_DIFFUSE;
        x.r = R; x.etaA = 1.0; x.etaB = m.eta;
        b->bx[b->n++] = x;
      }
      if (!sblack(T)) {
        BxDF x;
        x.kind = BX_SPEC_TRANS;
        x.type = BSDF_TRANSMISSION | BSDF_SPECULAR;
        x.t = T; x.etaA = 1.0; x.etaB = m.eta;
        b->bx[b->n++] = x;
      }
      return true;
    }
  }
  st->unsupported_material++;
  return false;
}

// ---------------------------------------------------------------- lights
// Sphere.Sample (sphere.go:270-285)
static inline void sphere_sample(const Sphere& s, P2 u, Intr* it, double* pdf) {
  V3 pObj = muls(uniform_sample_sphere(u), s.radius);
  V3 n = normalized(xf_normal(s.o2w, pObj));
  if (s.reverse) n = muls(n, -1);
  pObj = muls(pObj, s.radius / dist(pObj, V3{}));
  V3 pObjError = muls(vabs(pObj), gm::Gamma(5));
  it->p = xf_point(s.o2w, pObj, pObjError, &it->perr);
  it->n = n;
  *pdf = 1.0 / sphere_area(s);
}
// Sphere.SampleAtInteraction (sphere.go:287-344)
static inline void sphere_sample_at(const Sphere& s, const Intr& ref, P2 u, Intr* it, double* pdf) {
  V3 pCenter = xf_point(s.o2w, V3{}, V3{}, nullptr);
  V3 pOrigin = offset_ray_origin(ref.p, ref.perr, ref.n, sub(pCenter, ref.p));
  if (dist2(pOrigin, pCenter) <= s.radius * s.radius) {
    double p;
    sphere_sample(s, u, it, &p);
    V3 wi = sub(it->p, ref.p);
    if (len2(wi) == 0) p = 0;
    else {
      wi = normalized(wi);
      p *= dist2(ref.p, it->p) / absdot(it->n, muls(wi, -1));
    }
    if (std::isinf(p)) p = 0.0;
    *pdf = p;
    return;
  }
  V3 wc = normalized(sub(pCenter, ref.p));
  V3 wcX, wcY;
  coordinate_system(wc, &wcX, &wcY);
  double radius2 = s.radius * s.radius;
  double sinThetaMax2 = radius2 / dist2(ref.p, pCenter);
  double cosThetaMax = std::sqrt(gm::Max(0, 1.0 - sinThetaMax2));
  double cosTheta = (1.0 - u.x) + u.x * cosThetaMax;
  double sinTheta = std::sqrt(gm::Max(0, 1 - cosTheta * cosTheta));
  double phi = u.y * 2 * gm::Pi;
  double dc = dist(ref.p, pCenter);
  double ds = dc * cosTheta - std::sqrt(gm::Max(0, radius2 - (dc * dc) * (sinTheta * sinTheta)));
  double cosAlpha = (dc * dc + radius2 - ds * ds) / (2.0 * dc * s.radius);
  double sinAlpha = std::sqrt(gm::Max(0, 1.0 - cosAlpha * cosAlpha));
  // SphericalDirectionXYZ (geometry.go:66-70) with (-wcX, -wcY, -wc)
  V3 x = muls(wcX, -1), y = muls(wcY, -1), z = muls(wc, -1);
  V3 nWorld = add(add(muls(x, sinAlpha * gm::Cos(phi)), muls(y, sinAlpha * gm::Sin(phi))), muls(z, cosAlpha));
  V3 pWorld = add(pCenter, muls(nWorld, s.radius));
  it->p = pWorld;
  it->perr = muls(vabs(pWorld), gm::Gamma(5.0));
  it->n = nWorld;
  if (s.reverse) it->n = muls(it->n, -1);
  *pdf = uniform_cone_pdf(cosThetaMax);
}
// Disk.Sample (disk.go:160-170) + pbrt.SampleAtInteraction (shape.go:50-65)
static inline void disk_sample_at(const Disk& d, const Intr& ref, P2 u, Intr* it, double* pdf) {
  P2 pd = concentric_sample_disk(u);
  V3 pObj{pd.x * d.radius, pd.y * d.radius, d.height};
  V3 n = xf_normal(d.o2w, V3{0, 0, 1});
  if (d.reverse) n = muls(n, -1);
  it->n = n;
  it->p = xf_point(d.o2w, pObj, V3{}, &it->perr);
  double p = 1 / disk_area(d);
  V3 wi = sub(it->p, ref.p);
  if (len2(wi) == 0.0) { *pdf = 0; return; }
  wi = normalized(wi);
  p *= dist2(ref.p, it->p) / absdot(it->n, muls(wi, -1));
  if (std::isinf(p)) p = 0;
  *pdf = p;
}

struct LightSample { RGB Li; V3 wi; double pdf = 0; Intr p1; bool delta = false; };
// point.go:44-49, distant.go:40-44, diffuse.go:47-59
static inline void light_sample_li(const Scene& sc, const gopbrt_light& l, const Intr& ref, P2 u, LightSample* ls) {
  RGB E(l.rgb[0], l.rgb[1], l.rgb[2]);
  V3 v{l.v[0], l.v[1], l.v[2]};
  switch (l.kind) {
    case GOPBRT_LIGHT_POINT:
      ls->wi = normalized(sub(v, ref.p));
      ls->pdf = 1.0;
      ls->p1 = Intr{v, V3{}, V3{}};
      ls->Li = sdivs(E, dist2(v, ref.p));
      ls->delta = true;
      return;
    case GOPBRT_LIGHT_DISTANT:
      ls->p1 = Intr{muls(v, 2 * sc.world_radius), V3{}, V3{}};
      ls->Li = E;
      ls->wi = v;
      ls->pdf = 1;
      ls->delta = true;
      return;
    case GOPBRT_LIGHT_DIFFUSE_AREA: {
      ls->delta = false;
      Intr ps;
      double pdf;
      if (l.shape_kind == GOPBRT_SHAPE_SPHERE) sphere_sample_at(sc.spheres[l.shape_index], ref, u, &ps, &pdf);
      else disk_sample_at(sc.disks[l.shape_index], ref, u, &ps, &pdf);
      if (pdf == 0 || len2(sub(ps.p, ref.p)) == 0) { ls->Li = RGB(0); ls->wi = V3{}; ls->pdf = 0; return; }
      ls->wi = sub(ps.p, ref.p);  // un-normalised (diffuse.go:55)
      ls->p1 = ps;
      ls->pdf = pdf;
      V3 w = muls(ls->wi, -1);
      ls->Li = (l.two_sided || dot(ps.n, w) > 0) ? E : RGB(0);  // diffuse.go:36-41
      return;
    }
  }
}

// ---------------------------------------------------------------- integrator
struct Integ { int maxDepth; double rr; bool power = false; };

// EstimateDirect (integrator.go:79-195) for a SurfaceInteraction, handleMedia=false, specular=false
static inline RGB estimate_direct(const Scene& sc, const Hit& h, const BSDF& bsdf, P2 uScattering, const gopbrt_light& light,
                                  P2 uLight, RenderStats* st) {
  (void)uScattering;
  const int flags = BSDF_ALL & ~BSDF_SPECULAR;
  RGB Ld(0);
  Intr ref{h.p, h.perr, h.n};
  LightSample ls;
  light_sample_li(sc, light, ref, uLight, &ls);
  if (ls.pdf > 0 && !sblack(ls.Li)) {
    RGB f = bsdf_f(bsdf, h.wo, ls.wi, flags);
    f = smuls(f, absdot(ls.wi, h.ns));
    double scatteringPdf = bsdf_pdf(bsdf, h.wo, ls.wi, flags);
    if (g_trace)
      fprintf(g_trace, "  E pdf=%a Li=%a f=%a,%a,%a wi=%a,%a,%a ng=%a,%a,%a wo=%a,%a,%a ns=%a,%a,%a p=%a,%a,%a u=%a,%a shadow=%d\n", ls.pdf, ls.Li.c[0], f.c[0], f.c[1], f.c[2],
              ls.wi.x, ls.wi.y, ls.wi.z, bsdf.ng.x, bsdf.ng.y, bsdf.ng.z, h.wo.x, h.wo.y, h.wo.z, h.ns.x, h.ns.y, h.ns.z, h.p.x, h.p.y, h.p.z, uLight.x, uLight.y, (int)!sblack(f));
    if (!sblack(f)) {
      RGB Li = ls.Li;
      Ray sr = spawn_ray_to(ref, ls.p1, h.time);  // VisibilityTester.Unoccluded (light.go:46-48)
      st->shadow_rays++;
      TravStats ts;
      if (scene_intersect_p(sc, sr, &ts)) Li = RGB(0);
      st->snodes += ts.nodes; st->sprims += ts.prims;
      if (!sblack(Li)) {
        if (ls.delta) Ld = sadd(Ld, sdivs(smul(f, Li), ls.pdf));
        else {
          double weight = power_heuristic(ls.pdf, scatteringPdf);
          Ld = sadd(Ld, sdivs(smuls(smul(f, Li), weight), ls.pdf));
        }
      }
    }
  }
  // integrator.go:133-192: the BSDF-sampling MIS branch for non-delta lights never contributes (GetAreaLight() is
  // always nil, Light.Le is zero) and consumes no RNG.  Counted, not traced.
  if (!ls.delta) st->dead_mis_rays++;
  return Ld;
}

// Distribution1D.SampleDiscrete (sampling.go:42-55) via FindInterval (pkg/math/math.go:64-80); func[i] == 1 (uniform)
static inline void sample_discrete(const std::vector<double>& cdf, double func_int, double u, int* offset_out, double* pdf_out) {
  int size = (int)cdf.size(), first = 0, len = size;
  while (len > 0) {
    int half = len >> 1, middle = first + half;
    if (cdf[middle] <= u) { first = middle + 1; len -= half + 1; }
    else len = half;
  }
  *offset_out = (int)gm::Clamp((double)(first - 1), 0, (double)(size - 2));
  double pdf = 0;
  if (func_int > 0) pdf = 1.0 / (func_int / (double)(size - 1));
  *pdf_out = pdf;
}

// UniformSampleOneLight (integrator.go:48-77).  power_strategy: Path with LightSampleStrategy Power —
// ComputeLightPowerDistribution (lightdistribution.go:57-68) APPENDS the powers to a slice it already made n long, and
// every power is Spectrum.Y() == 0 (spectrum.go:227-229): 2n zeros, FuncInt == 0, so SampleDiscrete's pdf is 0 and the
// function returns black right after its Get1D.
static inline RGB uniform_sample_one_light(const Scene& sc, const Hit& h, const BSDF& bsdf, Sampler& smp, RenderStats* st, bool power_strategy = false) {
  int nLights = (int)sc.lights.size();
  if (nLights == 0) return RGB(0);
  double u = smp.get1d();
  if (power_strategy) return RGB(0);
  int offset;
  double lightPdf;
  sample_discrete(sc.light_cdf, sc.light_func_int, u, &offset, &lightPdf);
  if (lightPdf == 0.0) return RGB(0);
  P2 uLight = smp.get2d();
  P2 uScattering = smp.get2d();
  RGB s = estimate_direct(sc, h, bsdf, uScattering, sc.lights[offset], uLight, st);
  // spectrum.DivScalar(lightPdf) result is discarded (integrator.go:72)
  if (smax(s) > 10) st->radiance_gt10++;  // the reference panics here (integrator.go:73-75)
  return s;
}

// Path.Li (path.go:32-157)
static inline RGB path_li(const Scene& sc, Ray ray, Sampler& smp, const Integ& ig, RenderStats* st) {
  RGB L(0), beta(1.0);
  bool specularBounce = false;
  (void)specularBounce;
  int bounces = 0;
  double etaScale = 1.0;
  for (;;) {
    bounces++;
    Hit isect;
    TravStats ts;
    st->closest_rays++;
    bool found = scene_intersect(sc, ray, &isect, &ts);
    st->nodes += ts.nodes; st->prims += ts.prims;
    if (g_trace) fprintf(g_trace, "C sample=%d bounce=%d found=%d prim=%d shadow_so_far=%llu\n", smp.idx, bounces, (int)found, found ? isect.prim : -1, (unsigned long long)st->shadow_rays);
    // emission (path.go:48-63) is identically zero
    if (!found || bounces >= ig.maxDepth) break;
    BSDF bsdf;
    if (!compute_scattering(sc, isect, &bsdf, st)) break;
    if (bsdf_num_components(bsdf, BSDF_ALL & ~BSDF_SPECULAR) > 0) {
      RGB Ld = smul(beta, uniform_sample_one_light(sc, isect, bsdf, smp, st, ig.power));
      L = sadd(L, Ld);
    }
    V3 wo = ray.d;  // sic: not negated (path.go:91)
    P2 u = smp.get2d();
    RGB f; V3 wi; double pdf; int flags;
    bsdf_sample_f(bsdf, wo, u, BSDF_ALL, &f, &wi, &pdf, &flags);
    if (sblack(f) || pdf == 0.0) break;
    double wiAbsDotPdf = absdot(wi, isect.ns) / pdf;
    beta = smul(beta, smuls(f, wiAbsDotPdf));
    specularBounce = (flags & BSDF_SPECULAR) != 0;
    if ((flags & BSDF_SPECULAR) > 0 && (flags & BSDF_TRANSMISSION) > 0) {
      double eta = bsdf.eta;
      if (dot(wo, isect.n) > 0) etaScale *= eta * eta;
      else etaScale *= 1 / (eta * eta);
    }
    ray = spawn_ray(Intr{isect.p, isect.perr, isect.n}, wi, isect.time);  // wi is BSDF-local, used as world
    RGB rrBeta = smuls(beta, etaScale);
    if (smax(rrBeta) < ig.rr && bounces > 3) {
      double q = gm::Max(0.05, 1 - smax(rrBeta));
      if (smp.get1d() < q) break;
      beta = sdivs(beta, 1 - q);
    }
  }
  return L;
}

// ---------------------------------------------------------------- DirectLighting (pkg/integrator/directlighting.go)
// UniformSampleAllLights (integrator.go:23-46).  The sample arrays DirectLighting.Preprocess requests live on the
// prototype sampler only: PixelSampler.clone / RandomSampler.Clone build a fresh Sampler (pixel.go:34-42, random.go:33-37),
// so every worker's Get2DArray returns nil without consuming anything and each light takes the single-sample branch.
static inline RGB uniform_sample_all_lights(const Scene& sc, const Hit& h, const BSDF& bsdf, Sampler& smp, RenderStats* st) {
  RGB L(0);
  for (size_t j = 0; j < sc.lights.size(); j++) {
    P2 uLight = smp.get2d();
    P2 uScattering = smp.get2d();
    L = sadd(L, estimate_direct(sc, h, bsdf, uScattering, sc.lights[j], uLight, st));
  }
  return L;
}
struct DirectCfg { int maxDepth; int strategy; };  // strategy: 1 = UniformSampleAll, 2 = UniformSampleOne (directlighting.go:12-15)
static inline RGB direct_li(const Scene& sc, Ray ray, Sampler& smp, const DirectCfg& cfg, int depth, RenderStats* st);
// SamplerIntegratorSpecularReflect / SpecularTransmit (integrator.go:352-422); `type` = BSDF_REFLECTION or
// BSDF_TRANSMISSION | BSDF_SPECULAR.  The ray differentials they propagate feed nothing Li reads (the textures of the
// path ignore them), so they are not carried.  s.Li is entered with depth+1 on top of the depth+1 the caller passed.
static inline RGB specular_bounce(const Scene& sc, const Hit& isect, const BSDF& bsdf, Sampler& smp, const DirectCfg& cfg, int type,
                                  int depth, RenderStats* st) {
  P2 u = smp.get2d();
  RGB f; V3 wi; double pdf; int sampled;
  bsdf_sample_f(bsdf, isect.wo, u, type, &f, &wi, &pdf, &sampled);
  if (pdf > 0 && !sblack(f) && absdot(wi, isect.ns) != 0.0) {
    Ray rd = spawn_ray(Intr{isect.p, isect.perr, isect.n}, wi, isect.time);  // wi is BSDF-local, used as world
    return smuls(smul(f, direct_li(sc, rd, smp, cfg, depth + 1, st)), absdot(wi, isect.ns) / pdf);
  }
  return RGB(0);
}
// DirectLighting.Li (directlighting.go:62-104).  Light.Le and SurfaceInteraction.Le are identically zero.
static inline RGB direct_li(const Scene& sc, Ray ray, Sampler& smp, const DirectCfg& cfg, int depth, RenderStats* st) {
  RGB L(0);
  Hit isect;
  TravStats ts;
  st->closest_rays++;
  bool found = scene_intersect(sc, ray, &isect, &ts);
  st->nodes += ts.nodes; st->prims += ts.prims;
  if (!found) {
    for (size_t i = 0; i < sc.lights.size(); i++) L = sadd(L, RGB(0));
    return L;
  }
  BSDF bsdf;
  if (!compute_scattering(sc, isect, &bsdf, st, false)) return L;  // the reference panics (nil material / microfacet glass)
  L = sadd(L, RGB(0));  // si.Le(si.Wo)
  if (!sc.lights.empty()) {
    if (cfg.strategy == 1) L = sadd(L, uniform_sample_all_lights(sc, isect, bsdf, smp, st));
    else L = sadd(L, uniform_sample_one_light(sc, isect, bsdf, smp, st));
  }
  if (depth + 1 < cfg.maxDepth) {
    L = sadd(L, specular_bounce(sc, isect, bsdf, smp, cfg, BSDF_REFLECTION | BSDF_SPECULAR, depth + 1, st));
    L = sadd(L, specular_bounce(sc, isect, bsdf, smp, cfg, BSDF_TRANSMISSION | BSDF_SPECULAR, depth + 1, st));
  }
  return L;
}

// ---------------------------------------------------------------- camera (camera.go:192-242)
static inline Ray camera_ray(const gopbrt_camera& c, P2 pFilm, P2 pLens, double time) {
  Xf r2c, c2w;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { r2c.m.m[i][j] = c.raster_to_camera[i * 4 + j]; c2w.m.m[i][j] = c.camera_to_world[i * 4 + j]; }
  r2c.inv = r2c.m; c2w.inv = c2w.m;  // inverses are never read on this path
  V3 pCamera = xf_point(r2c, V3{pFilm.x, pFilm.y, 0}, V3{}, nullptr);
  Ray ray{V3{0, 0, 0}, normalized(pCamera), gm::Inf, 0};
  if (c.lens_radius > 0) {
    P2 pl = concentric_sample_disk(pLens);
    pl.x *= c.lens_radius; pl.y *= c.lens_radius;
    double ft = c.focal_distance / ray.d.z;
    V3 pFocus = add(muls(ray.d, ft), ray.o);
    ray.o = V3{pl.x, pl.y, 0};
    ray.d = normalized(sub(pFocus, ray.o));
  }
  ray = xf_ray(c2w, ray, nullptr, nullptr);  // AnimatedTransform.TransformRay, not animated (transform.go:592-596)
  ray.time = (1.0 - time) * c.shutter_open + time * c.shutter_close;  // math.Lerp (math.go:106-108)
  return ray;
}

// ---------------------------------------------------------------- film (film.go)
struct Bounds2i { int64_t x0, y0, x1, y1; };
struct FilmCfg {
  Bounds2i cropped;
  double rx, ry;
  double table[256];
};
static inline FilmCfg film_cfg(const gopbrt_film& f) {  // film.go:43-76
  FilmCfg c;
  c.cropped = Bounds2i{(int64_t)std::ceil((double)f.width * f.crop[0]), (int64_t)std::ceil((double)f.height * f.crop[1]),
                       (int64_t)std::ceil((double)f.width * f.crop[2]), (int64_t)std::ceil((double)f.height * f.crop[3])};
  c.rx = f.filter_radius[0]; c.ry = f.filter_radius[1];
  for (int i = 0; i < 256; i++) c.table[i] = 1.0;  // BoxFilter.Evaluate (filter.go:30-32)
  return c;
}
struct FilmTile {
  Bounds2i pb;
  std::vector<double> px;  // contribSum rgb + filterWeightSum
  int64_t tile_index = 0;
};
static inline FilmTile film_tile(const FilmCfg& f, Bounds2i sb) {  // film.go:106-113
  int64_t p0x = (int64_t)std::ceil((double)sb.x0 - 0.5 - f.rx), p0y = (int64_t)std::ceil((double)sb.y0 - 0.5 - f.ry);
  int64_t p1x = (int64_t)std::floor((double)sb.x1 - 0.5 + f.rx) + 1, p1y = (int64_t)std::floor((double)sb.y1 - 0.5 + f.ry) + 1;
  FilmTile t;
  // Bounds2i.Intersect (bounds.go:93-98)
  t.pb = Bounds2i{std::max(f.cropped.x0, p0x), std::max(f.cropped.y0, p0y), std::min(f.cropped.x1, p1x), std::min(f.cropped.y1, p1y)};
  int64_t area = std::max<int64_t>(0, (t.pb.x1 - t.pb.x0) * (t.pb.y1 - t.pb.y0));
  t.px.assign(area * 4, 0.0);
  return t;
}
static inline void film_add_sample(const FilmCfg& f, FilmTile& t, P2 pFilm, RGB L, double w) {  // film.go:211-248
  double dx = pFilm.x - 0.5, dy = pFilm.y - 0.5;
  double p0fx = std::ceil(dx - f.rx), p0fy = std::ceil(dy - f.ry);
  double p1fx = std::floor(dx + f.rx) + 1, p1fy = std::floor(dy + f.ry) + 1;
  int64_t p0x = (int64_t)gm::Max(p0fx, (double)t.pb.x0), p0y = (int64_t)gm::Max(p0fy, (double)t.pb.y0);
  int64_t p1x = (int64_t)gm::Min(p1fx, (double)t.pb.x1), p1y = (int64_t)gm::Min(p1fy, (double)t.pb.y1);
  double invx = 1.0 / f.rx, invy = 1.0 / f.ry;
  int64_t width = t.pb.x1 - t.pb.x0;
  for (int64_t y = p0y; y < p1y; y++) {
    double fy = std::fabs(((double)y - dy) * invy * 16.0);
    int iy = (int)gm::Min(std::floor(fy), 16.0 - 1);
    for (int64_t x = p0x; x < p1x; x++) {
      double fx = std::fabs(((double)x - dx) * invx * 16.0);
      int ix = (int)gm::Min(std::floor(fx), 16.0 - 1);
      double fw = f.table[iy * 16 + ix];
      double* p = &t.px[((x - t.pb.x0) + (y - t.pb.y0) * width) * 4];
      RGB c = smuls(L, w * fw);
      p[0] += c.c[0]; p[1] += c.c[1]; p[2] += c.c[2];
      p[3] += fw;
    }
  }
}
static inline void film_merge(const FilmCfg& f, const FilmTile& t, double* film) {  // film.go:115-132 + spectrum.go:35-41
  int64_t fw = f.cropped.x1 - f.cropped.x0, tw = t.pb.x1 - t.pb.x0;
  for (int64_t y = t.pb.y0; y < t.pb.y1; y++)
    for (int64_t x = t.pb.x0; x < t.pb.x1; x++) {
      const double* p = &t.px[((x - t.pb.x0) + (y - t.pb.y0) * tw) * 4];
      double* m = &film[((x - f.cropped.x0) + (y - f.cropped.y0) * fw) * 4];
      double X = 0.412453 * p[0] + 0.357580 * p[1] + 0.180423 * p[2];
      double Y = 0.212671 * p[0] + 0.715160 * p[1] + 0.072169 * p[2];
      double Z = 0.019334 * p[0] + 0.119193 * p[1] + 0.950227 * p[2];
      m[0] += X; m[1] += Y; m[2] += Z;
      m[3] += p[3];
    }
}

struct RenderOpts {
  int rank = 0, world = 1;
  int threads = 1;
  int deterministic = 1;              // 1: merge tiles in tile-index order after all workers finish
  int64_t tile_begin = 0, tile_end = -1;  // bounded sample of the workload (bench.py cpu_baseline)
};

// pbrt.Render + renderWorker (integrator.go:228-350)
static inline void render(const Scene& sc, const gopbrt_camera& cam, const gopbrt_sampler& scfg, const gopbrt_integrator& icfg,
                          const gopbrt_film& fcfg, const RenderOpts& opt, double* film, RenderStats* stats_out) {
  g_trace = nullptr;
  if (const char* tp = getenv("ORACLE_TRACE")) if (opt.threads <= 1) g_trace = fopen(tp, "w");
  FilmCfg fc = film_cfg(fcfg);
  Bounds2i sb = fc.cropped;  // GetSampleBounds returns CroppedPixelBounds (film.go:84-95)
  int64_t ts = icfg.tile_size;
  int64_t ntx = ((sb.x1 - sb.x0) + ts - 1) / ts, nty = ((sb.y1 - sb.y0) + ts - 1) / ts;
  int64_t ntiles = ntx * nty;
  int64_t fw = fc.cropped.x1 - fc.cropped.x0, fh = fc.cropped.y1 - fc.cropped.y0;
  std::fill(film, film + fw * fh * 4, 0.0);
  int64_t t_begin = opt.tile_begin, t_end = opt.tile_end < 0 ? ntiles : std::min(opt.tile_end, ntiles);
  std::atomic<int64_t> next{t_begin};
  std::mutex mu;
  RenderStats total;
  std::vector<std::vector<FilmTile>> done(std::max(1, opt.threads));
  Integ ig{icfg.max_depth, icfg.rr_threshold, icfg.kind == GOPBRT_INTEGRATOR_PATH && icfg.light_strategy == 2};
  DirectCfg dcfg{icfg.max_depth, icfg.light_strategy};
  auto worker = [&](int tid) {
    RenderStats st;
    Sampler smp;
    smp.init(scfg);
    for (;;) {
      int64_t t = next.fetch_add(1);
      if (t >= t_end) break;
      if (opt.world > 1 && (t % opt.world) != opt.rank && scfg.mode == GOPBRT_MODE_STRICT) continue;
      int64_t tx = t % ntx, ty = t / ntx;
      smp.clone_seed((uint64_t)(ty * ntx + tx));  // integrator.go:318,328
      int64_t x0 = sb.x0 + tx * ts, x1 = (int64_t)gm::Min((double)(x0 + ts), (double)sb.x1);
      int64_t y0 = sb.y0 + ty * ts, y1 = (int64_t)gm::Min((double)(y0 + ts), (double)sb.y1);
      FilmTile tile = film_tile(fc, Bounds2i{x0, y0, x1, y1});
      tile.tile_index = t;
      for (int64_t py = y0; py < y1; py++)
        for (int64_t px = x0; px < x1; px++) {
          smp.fast_pixel = (uint64_t)((py - sb.y0) * fw + (px - sb.x0));
          smp.start_pixel();
          while (smp.start_next_sample()) {
            if (scfg.mode == GOPBRT_MODE_FAST && opt.world > 1 && (smp.idx % opt.world) != opt.rank) continue;
            // GetCameraSample (sampler.go:75-80): Get2D pFilm, Get2D pLens, Get1D time
            P2 o = smp.get2d();
            P2 pFilm{(double)px + o.x, (double)py + o.y};
            P2 pLens = smp.get2d();
            double time = smp.get1d();
            Ray ray = camera_ray(cam, pFilm, pLens, time);
            st.camera_rays++;
            RGB L = icfg.kind == GOPBRT_INTEGRATOR_DIRECT_LIGHTING ? direct_li(sc, ray, smp, dcfg, 0, &st) : path_li(sc, ray, smp, ig, &st);
            if (snan(L)) { L = RGB(0.1); st.nan_samples++; }  // integrator.go:256-257 (the Y() guards are dead)
            film_add_sample(fc, tile, pFilm, L, 1.0);
          }
        }
      if (opt.deterministic) done[tid].push_back(std::move(tile));
      else { std::lock_guard<std::mutex> g(mu); film_merge(fc, tile, film); }
    }
    std::lock_guard<std::mutex> g(mu);
    total.add(st);
  };
  std::vector<std::thread> th;
  for (int i = 1; i < opt.threads; i++) th.emplace_back(worker, i);
  worker(0);
  for (auto& t : th) t.join();
  if (opt.deterministic) {
    std::vector<const FilmTile*> all;
    for (auto& v : done) for (auto& t : v) all.push_back(&t);
    std::sort(all.begin(), all.end(), [](const FilmTile* a, const FilmTile* b) { return a->tile_index < b->tile_index; });
    for (auto* t : all) film_merge(fc, *t, film);
  }
  if (stats_out) *stats_out = total;
  if (g_trace) { fclose(g_trace); g_trace = nullptr; }
}

}  // namespace oracle
