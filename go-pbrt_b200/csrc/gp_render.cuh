// gp_render.cuh — the wavefront path integrator: raygen+sampler, shade (materials, textures, lights), shadow
// resolve and film kernels.  Restates pbrt.Render / renderWorker (pkg/pbrt/integrator.go:228-350), Path.Li
// (pkg/integrator/path.go:32-157), UniformSampleOneLight / EstimateDirect (integrator.go:48-195), the samplers
// (pkg/sampler/*.go, pkg/pbrt/{rng,sampling}.go), the BSDFs (pkg/pbrt/reflection.go), the materials / textures
// (pkg/materials, pkg/textures, pkg/pbrt/texture.go), the lights (pkg/lights) and the film (pkg/pbrt/film.go),
// quirks included.
//
// A *lane* is one independent sampler stream of the reference: in STRICT mode one image tile (Clone(seed = tile
// index), integrator.go:318-328) whose pixels and samples are walked sequentially with the RNG state carried
// along; pbrt.Render(…, tileSize = 1) makes every pixel its own lane.  All per-lane state is structure-of-arrays
// with the lane index fastest, so a warp's accesses coalesce.  Stages hand lanes to each other through index
// queues filled with warp-ballot compaction and one aggregated atomicAdd per warp.
#pragma once
#include "gp_trace.cuh"

namespace gp {

struct Lanes {
  long long n;
  RayRec* ray;      // current path segment + closest-hit result
  ShadowRec* sray;  // pending visibility segment + gated light sample
  PathRec* path;    // throughput, sampler stream, film sum
  RadRec* rad;      // radiance sum + refraction scale (one sector per lane, an array of its own)
  FilmRec* fsum;    // pixel cursor, DirectLighting segment mask, uniform-footprint film sum (likewise)
  double* tables;   // [dim][k][lane] stratified 1-D tables
  double* tilepix;  // [lane][tile pixel][4] FilmTile accumulators (one contiguous record per lane)
  long long tile_stride;  // doubles per lane = tpw * tph * 4
  double* frames;   // DirectLighting only: [lane][level][8] = {D.rgb, f.rgb, c, -} of the Li recursion (see shade_lane_direct)
  unsigned char* occl;  // DirectLighting / UniformSampleAll only: [lane][segment] visibility results of the shadow stage
};

struct RenderParams {
  M4 raster_to_camera, camera_to_world;
  double lens_radius, focal_distance, shutter_open, shutter_close;
  int sampler_kind, xs, ys, jitter, ndims, mode, spp;
  int max_depth;
  int integrator;                // 0 = Path, 1 = DirectLighting (UniformSampleOne)
  int direct_levels;             // DirectLighting: frames per lane = max(1, maxDepth / 2)
  int n_seg;                     // shadow segments per lane: 1, or the number of lights for DirectLighting / UniformSampleAll
  int direct_all;                // DirectLighting strategy is UniformSampleAll (segments are collected by the shade stage)
  int uniform_fp;                // every sample of a lane has the same film footprint and weight: the lane's FilmTile is kept as one
                                 // RGB sum in PathRec.pad (see film_add_uniform) instead of a tile record in `tilepix`
  int light_power;               // Path with LightSampleStrategy Power (lightdistribution.go:44-68, bugs included: no light is ever sampled)
  double rr_threshold;
  long long tile_size, ntx, nty, ntiles;
  long long cx0, cy0, cx1, cy1;  // CroppedPixelBounds
  double frx, fry;               // filter radius
  int tpw, tph;                  // allocated tile-pixel-bounds extent
  int rank, world;               // tile partition (STRICT): this rank owns tiles t with t % world == rank
  int s_rank, s_world;           // sample partition (FAST): this rank owns samples s with s % s_world == s_rank
  int groups;                    // FAST: lanes per tile; lane-slot k works on tile k / groups and, of this rank's samples,
                                 // on those with s % (s_world * groups) == s_rank * groups + k % groups  (1 in STRICT mode)
  long long lane_base;           // first lane-slot (tile = ((lane_base + lane) / groups) * world + rank) of this pass
  long long lanes_active;        // lanes in use this pass
  int groups_merged;             // film merge only: k_group_sums has folded every tile's lane groups into group 0's record
  int last_in_place;             // FAST + uniform footprint + Path: a lane's LAST sample is not retired through the regeneration
                                 // queue — its radiance stays in PathRec.L and the film fold adds it (see lane_on_last_sample)
};

// FAST mode with the uniform footprint (tileSize 1: the lane's tile is one pixel) and P.last_in_place: is the sample the lane
// is working on (index sidx) the last one of its share?  The lane's samples are s = sidx, sidx + s_mod, ... < spp.
// Such a sample is never sent through the regeneration queue: nothing would be generated after it, and the only thing its
// retirement does — pad += L (film_add_uniform) — is done by the film fold (k_group_sums / k_fold_last), which adds pad + L
// in the same order.  With one sample per lane (config 2 on one B200: 63 lane groups) raygen then runs once per frame, and
// 130 M PathRecs are not read and rewritten once more just to move 24 bytes inside them.
GP_D bool lane_on_last_sample(const RenderParams& P, int sidx) {
  const int s_mod = P.s_world * P.groups;
  return P.last_in_place && sidx + (s_mod > 1 ? s_mod : 1) >= P.spp;
}

struct RenderCounters {
  unsigned long long camera_rays, closest_rays, shadow_rays, dead_mis_rays, radiance_gt10, nan_samples, unsupported, efloat_panics;
  unsigned long long root_culled;  // scene.Intersect queries answered by the root-bound test inside raygen
  unsigned long long shaded;       // lanes handed to the shade stage (hits)
};

struct Queues {
  int *extend, *extend_next, *shadow, *regen, *regen_next;
  int* shade[4];  // hit lanes binned by shade class (counts in cnt[8..11])
  int* cnt;  // [0]=extend [1]=extend_next [2]=shadow [3]=regen [4]=regen_next [5]=shade (hits) [6]=extend work counter [7]=any-hit work counter
};

GP_D void queue_push(int* q, int* cnt, bool pred, int v) {
  unsigned m = __ballot_sync(0xffffffffu, pred);
  if (m == 0) return;
  int lane_id = threadIdx.x & 31;
  int leader = __ffs(m) - 1;
  int base = 0;
  if (lane_id == leader) base = atomicAdd(cnt, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (pred) q[base + __popc(m & ((1u << lane_id) - 1u))] = v;
}

// CTA-wide pushes (every thread of the CTA must call them, inside CTA-uniform control flow): one atomic per CTA and
// queue, and the CTA's entries land in the queue as one contiguous run in thread order, which keeps neighbouring lanes
// (neighbouring pixels / sample groups) next to each other for the stage that consumes the queue.
template <int NQ>
GP_D void block_push(int* const (&q)[NQ], int* const (&cnt)[NQ], const bool (&pred)[NQ], int v, int (*s_cnt)[NQ], int* s_base) {
  const unsigned FULL = 0xffffffffu;
  const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  unsigned m[NQ];
#pragma unroll
  for (int k = 0; k < NQ; k++) { m[k] = __ballot_sync(FULL, pred[k]); if (lane_id == 0) s_cnt[warp][k] = __popc(m[k]); }
  __syncthreads();
  if (threadIdx.x < NQ) {  // one thread per queue: the warps' counts become their offsets in the CTA's run, then the queue's one atomic
    int tot = 0;
    for (int w = 0; w < nwarps; w++) { const int c = s_cnt[w][threadIdx.x]; s_cnt[w][threadIdx.x] = tot; tot += c; }
    s_base[threadIdx.x] = tot ? atomicAdd(cnt[threadIdx.x], tot) : 0;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NQ; k++)
    if (pred[k]) q[k][s_base[k] + s_cnt[warp][k] + __popc(m[k] & ((1u << lane_id) - 1u))] = v;
  __syncthreads();
}

// The same for U entries per thread at once.  A CTA-wide push costs three barriers, and a barrier costs the spread of its
// warps' arrival times — in the raygen and shade kernels, whose warps run hundreds to thousands of instructions between two
// pushes at their own pace, a quarter / an eighth of all stall samples (ncu).  So a raygen thread works through U queue entries
// (the CTA through U * blockDim consecutive ones), keeps what it wants pushed — values in shared memory, the predicates as a
// bit mask, bit u * NQ + k for entry u and queue k — and the CTA pushes once per batch: the barriers are paid once per U lanes,
// and the CTA's run in the output queue is U times longer, in input order (entry-major, then thread order).
constexpr int kPushBatch = 4;
template <int NQ, int U>
struct PushBatch {
  int cnt[U][8][NQ];   // [entry][warp][queue] number of pushes
  int off[U][8][NQ];   // exclusive prefix of cnt in (entry, warp) order
  int base[NQ];
  int val[U][256];     // [entry][thread] the value to push
};
template <int NQ, int U>
GP_D void block_push_batch(int* const (&q)[NQ], int* const (&cnt)[NQ], unsigned flags, PushBatch<NQ, U>& S) {
  const unsigned FULL = 0xffffffffu;
  const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  unsigned m[U][NQ];
#pragma unroll
  for (int u = 0; u < U; u++)
#pragma unroll
    for (int k = 0; k < NQ; k++) {
      m[u][k] = __ballot_sync(FULL, (flags >> (u * NQ + k)) & 1u);
      if (lane_id == 0) S.cnt[u][warp][k] = __popc(m[u][k]);
    }
  __syncthreads();
  if (threadIdx.x < NQ) {  // one thread per queue: the prefix over (entry, warp) and the queue's one atomic
    int tot = 0;
    for (int u = 0; u < U; u++)
      for (int w2 = 0; w2 < nwarps; w2++) { S.off[u][w2][threadIdx.x] = tot; tot += S.cnt[u][w2][threadIdx.x]; }
    S.base[threadIdx.x] = tot ? atomicAdd(cnt[threadIdx.x], tot) : 0;
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < U; u++)
#pragma unroll
    for (int k = 0; k < NQ; k++)
      if ((flags >> (u * NQ + k)) & 1u) q[k][S.base[k] + S.off[u][warp][k] + __popc(m[u][k] & ((1u << lane_id) - 1u))] = S.val[u][threadIdx.x];
  __syncthreads();
}

// ---------------------------------------------------------------- RNG (pkg/pbrt/rng.go — a PCG32 *variant*)
struct Smp {
  unsigned long long state, inc;
  int cur1, cur2, sidx;
  long long lane;
};
GP_D uint32_t rng_u32(Smp& s) {
  unsigned long long old = s.state;
  s.state = old * 0x5851f42d4c957f2dULL + s.inc;
  uint32_t xs = (uint32_t)(((old >> 18) ^ old) >> 27);
  uint32_t rot = (uint32_t)(old >> 59);
  return (xs >> rot) | (xs << ((rot + 1u) & 31u));  // sic: not a rotate (rng.go:41)
}
GP_D void rng_set_sequence(Smp& s, unsigned long long seed) {
  s.state = 0;
  s.inc = (seed << 1) | 1ULL;
  rng_u32(s);
  s.state += 0x853c49e6748fea9bULL;
  rng_u32(s);
}
GP_D uint32_t rng_u32b(Smp& s, uint32_t b) {
  uint32_t threshold = (~b + 1u) % b;
  for (;;) {
    uint32_t r = rng_u32(s);
    if (r >= threshold) return r % b;
  }
}
GP_D double rng_uniform(Smp& s) { return go_min(one_minus_epsilon(), (double)rng_u32(s) * 2.3283064365386963e-10); }

GP_D uint32_t kensler_permute(uint32_t i, uint32_t l, uint32_t p) {
  uint32_t w = l - 1;
  w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
  do {
    i ^= p; i *= 0xe170893d; i ^= p >> 16; i ^= (i & w) >> 4; i ^= p >> 8; i *= 0x0929eb3f; i ^= p >> 23;
    i ^= (i & w) >> 1; i *= 1 | p >> 27; i *= 0x6935fa69; i ^= (i & w) >> 11; i *= 0x74dcb303; i ^= (i & w) >> 2;
    i *= 0x9e501cc3; i ^= (i & w) >> 2; i *= 0xc860a3df; i &= w; i ^= i >> 5;
  } while (i >= l);
  return (i + p) % l;
}
GP_D uint32_t hash_u32(unsigned long long a, uint32_t b) {
  unsigned long long x = a * 0x9E3779B97F4A7C15ULL + (unsigned long long)b * 0xD1B54A32D192ED03ULL + 0x632BE59BD9B4E019ULL;
  x ^= x >> 32; x *= 0xD6E8FEB86659FD93ULL; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ULL; x ^= x >> 32;
  return (uint32_t)x;
}

// PixelSampler.Get1D / Get2D (pkg/sampler/pixel.go:60-80), RandomSampler (random.go:21-27)
GP_D double get1d(Smp& s, const Lanes& L, const RenderParams& P, unsigned long long fast_pixel) {
  int nd = P.sampler_kind == 0 ? P.ndims : 0;
  if (P.mode == 1) {
    if (s.cur1 < nd) {
      uint32_t j = kensler_permute((uint32_t)s.sidx, (uint32_t)P.spp, hash_u32(fast_pixel, (uint32_t)s.cur1));
      s.cur1++;
      double delta = 0.5;
      if (P.jitter) delta = rng_uniform(s);
      return go_min(((double)j + delta) * (1.0 / (double)P.spp), one_minus_epsilon());
    }
    return rng_uniform(s);
  }
  if (s.cur1 < nd) {
    double v = L.tables[((size_t)s.cur1 * P.spp + s.sidx) * L.n + s.lane];
    s.cur1++;
    return v;
  }
  return rng_uniform(s);
}
GP_D void get2d(Smp& s, const RenderParams& P, double* x, double* y) {
  int nd = P.sampler_kind == 0 ? P.ndims : 0;
  if (s.cur2 < nd) { s.cur2++; *x = 0; *y = 0; return; }  // StratifiedSample2D never writes its table
  *x = rng_uniform(s);
  *y = rng_uniform(s);
}

// Stratified.StartPixel (stratified.go:21-48)
GP_D void start_pixel(Smp& s, const Lanes& L, const RenderParams& P) {
  if (P.sampler_kind == 0 && P.mode == 0) {
    int n = P.spp;
    double inv = 1.0 / (double)n;
    for (int d = 0; d < P.ndims; d++) {
      double* tab = L.tables + (size_t)d * P.spp * L.n + s.lane;
      for (int k = 0; k < n; k++) {  // StratifiedSample1D (sampling.go:101-110)
        double delta = 0.5;
        if (P.jitter) delta = rng_uniform(s);
        tab[(size_t)k * L.n] = go_min(((double)k + delta) * inv, one_minus_epsilon());
      }
      for (int k = 0; k < n; k++) {  // ShuffleSamples1D (sampling.go:129-136)
        int other = k + (int)rng_u32b(s, (uint32_t)(n - k));
        double a = tab[(size_t)k * L.n], b = tab[(size_t)other * L.n];
        tab[(size_t)k * L.n] = b;
        tab[(size_t)other * L.n] = a;
      }
    }
    for (int d = 0; d < P.ndims; d++) {
      if (P.jitter)
        for (int k = 0; k < n; k++) { rng_uniform(s); rng_uniform(s); }  // StratifiedSample2D draws, writes nothing
      for (int k = 0; k < n; k++) rng_u32b(s, (uint32_t)(n - k));        // ShuffleSamples2D permutes zeros
    }
  }
  s.sidx = 0;
}

// Quotient / remainder of non-negative 64-bit integers.  Tile, pixel and lane indices nearly always fit 32 bits, where the
// hardware-assisted unsigned division is a few instructions; the generic 64-bit division is an ~100-instruction
// subroutine, and the per-lane index arithmetic of raygen and shade used eight of them per sample.
GP_D long long div_nn(long long a, long long b) {
  if ((((unsigned long long)a | (unsigned long long)b) >> 32) == 0) return (long long)((unsigned)a / (unsigned)b);
  return a / b;
}
GP_D long long mod_nn(long long a, long long b) {
  if ((((unsigned long long)a | (unsigned long long)b) >> 32) == 0) return (long long)((unsigned)a % (unsigned)b);
  return a % b;
}

// ---------------------------------------------------------------- sampling warps (pkg/pbrt/sampling.go)
GP_D void concentric_sample_disk(double ux, double uy, double* ox, double* oy) {
  double x = ux * 2.0 - 1, y = uy * 2.0 - 1;
  if (x == 0 && y == 0) { *ox = 0; *oy = 0; return; }
  // sampling.go:183-189: r = x, theta = Pi/4 * (y / x) where |x| > |y|, else r = y, theta = Pi/2 - Pi/4 * (x / y) — ONE
  // division behind selects (each case performs its own operations; the two branches split every warp down the middle)
  const bool wide = fabs(x) > fabs(y);
  const double r = wide ? x : y;
  const double q = kPiOver4 * ((wide ? y : x) / r);
  const double theta = wide ? q : kPiOver2 - q;
  const SinCos sc_ = go_sincos(theta);
  *ox = sc_.cs * r;
  *oy = sc_.sn * r;
}
GP_D V3 cosine_sample_hemisphere(double ux, double uy) {
  double dx, dy;
  concentric_sample_disk(ux, uy, &dx, &dy);
  double z = sqrt(go_max(0.0, 1.0 - dx * dx - dy * dy));
  return mk3(dx, dy, z);
}
GP_D V3 uniform_sample_sphere(double ux, double uy) {
  double z = 1.0 - 2.0 * ux;
  double r = sqrt(go_max(0, 1 - z * z));
  double phi = 2 * kPi * uy;
  const SinCos sc_ = go_sincos(phi);
  return mk3(r * sc_.cs, r * sc_.sn, z);
}

// ---------------------------------------------------------------- spectrum helpers (pkg/pbrt/spectrum.go)
struct RGB { double r, g, b; };
GP_D RGB rgb(double r, double g, double b) { RGB c; c.r = r; c.g = g; c.b = b; return c; }
GP_D RGB operator*(RGB a, RGB b) { return rgb(a.r * b.r, a.g * b.g, a.b * b.b); }
GP_D RGB operator*(RGB a, double s) { return rgb(a.r * s, a.g * s, a.b * s); }
GP_D RGB operator/(RGB a, double s) { return rgb(a.r / s, a.g / s, a.b / s); }
GP_D RGB operator+(RGB a, RGB b) { return rgb(a.r + b.r, a.g + b.g, a.b + b.b); }
GP_D bool is_black(RGB a) { return !(a.r != 0.0) && !(a.g != 0.0) && !(a.b != 0.0); }
GP_D double max_comp(RGB a) { return go_max(go_max(a.r, a.g), a.b); }
GP_D RGB clamp_rgb(RGB a, double lo, double hi) { return rgb(go_clamp(a.r, lo, hi), go_clamp(a.g, lo, hi), go_clamp(a.b, lo, hi)); }

// ---------------------------------------------------------------- textures (texture.go, checkerboard.go)
GP_D RGB tex_eval(const DevScene& sc, int id, const Hit& h) {
  for (int guard = 0; guard < 64; guard++) {
    const TextureDev& t = sc.textures[id];
    if (t.kind == 0) return rgb(t.rgb[0], t.rgb[1], t.rgb[2]);
    double s, tt;
    if (t.mapping == 1) {  // PlanarMapping2D.Map (texture.go:42-46)
      s = t.ds + dot(h.p, mk3(t.vs[0], t.vs[1], t.vs[2]));
      tt = t.dt + dot(h.p, mk3(t.vt[0], t.vt[1], t.vt[2]));
    } else {  // UVMapping2D.Map (texture.go:22-26)
      s = t.su * h.u + t.du;
      tt = t.sv * h.v + t.dv;
    }
    long long k = (long long)(floor(s) + floor(tt));  // checkerboard.go:32
    id = (k % 2 == 0) ? t.tex1 : t.tex2;
  }
  return rgb(0, 0, 0);
}

// ---------------------------------------------------------------- BSDF (reflection.go)
enum { BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16, BSDF_ALL = 31 };
enum { BX_NONE = -1, BX_LAMBERT = 0, BX_OREN_NAYAR = 1, BX_SPEC_REFL_NOOP = 2, BX_FRESNEL_SPECULAR = 3, BX_GLASS_SPLIT = 4 };
// every material of the reference's working subset builds at most ONE BxDF (matte.go, mirror.go, glass.go:45-46)
struct BSDF {
  V3 ns, ng, ss, ts;
  double eta;
  int kind, type;
  RGB r, t;
  double a, b, etaB;
};
GP_D bool matches(int t, int flags) { return (t & flags) == t; }
GP_D double fr_dielectric(double cosThetaI, double etaI, double etaT) {  // reflection.go:21-42
  cosThetaI = go_clamp(cosThetaI, -1, 1);
  bool entering = cosThetaI > 0;
  if (!entering) { double t = etaI; etaI = etaT; etaT = t; cosThetaI = fabs(cosThetaI); }
  double sinThetaI = sqrt(go_max(0, 1 - cosThetaI * cosThetaI));
  double sinThetaT = etaI / etaT * sinThetaI;
  if (sinThetaT >= 1) return 1;
  double cosThetaT = sqrt(go_max(0, 1 - sinThetaT * sinThetaT));
  double Rparl = ((etaT * cosThetaI) - (etaI * cosThetaT)) / ((etaT * cosThetaI) + (etaI * cosThetaT));
  double Rperp = ((etaI * cosThetaI) - (etaT * cosThetaT)) / ((etaI * cosThetaI) + (etaT * cosThetaT));
  return (Rparl * Rparl + Rperp * Rperp) / 2;
}
GP_D double sin2theta(V3 w) { return go_max(0, 1 - w.z * w.z); }
GP_D double sintheta(V3 w) { return sqrt(sin2theta(w)); }
GP_D double cosphi(V3 w) { double s = sintheta(w); return s == 0 ? 1 : go_clamp(w.x / s, -1, 1); }
GP_D double sinphi(V3 w) { double s = sintheta(w); return s == 0 ? 0 : go_clamp(w.y / s, -1, 1); }

GP_D RGB bxdf_f(const BSDF& b, V3 wo, V3 wi, bool lambert_only = false) {
  if (lambert_only || b.kind == BX_LAMBERT) return b.r * kInvPi;  // reflection.go:589-591
  if (b.kind == BX_OREN_NAYAR) {                   // reflection.go:628-652
    double sinThetaI = sintheta(wi), sinThetaO = sintheta(wo);
    double maxCos = 0.0;
    if (sinThetaI > 1e-4 && sinThetaO > 1e-4) {
      double sinPhiI = sinphi(wi), cosPhiI = cosphi(wi), sinPhiO = sinphi(wo), cosPhiO = cosphi(wo);
      double dCos = cosPhiI * cosPhiO + sinPhiI * sinPhiO;
      maxCos = go_max(0.0, dCos);
    }
    double sinAlpha, tanBeta;
    if (fabs(wi.z) > fabs(wo.z)) { sinAlpha = sinThetaO; tanBeta = sinThetaO / fabs(wo.z); }
    else { sinAlpha = sinThetaI; tanBeta = sinThetaO / fabs(wo.z); }
    return b.r * (kInvPi * (b.a + b.b * maxCos * sinAlpha * tanBeta));
  }
  return rgb(0, 0, 0);
}
GP_D double bxdf_pdf(const BSDF& b, V3 wo, V3 wi, bool lambert_only = false) {
  if (lambert_only || b.kind == BX_LAMBERT || b.kind == BX_OREN_NAYAR) {  // reflection.go:343-348
    if (wo.z * wi.z > 0) return fabs(wi.z) * kInvPi;
    return 0;
  }
  return 0;
}
GP_D V3 to_local(const BSDF& b, V3 v) { return mk3(dot(v, b.ss), dot(v, b.ts), dot(v, b.ns)); }

// BSDF.F (reflection.go:164-181) / BSDF.Pdf (:255-278) for a single-BxDF BSDF
GP_D RGB bsdf_f(const BSDF& b, V3 woW, V3 wiW, int flags, bool lambert_only = false) {
  V3 wi = to_local(b, wiW), wo = to_local(b, woW);
  if (wo.z == 0.0) return rgb(0, 0, 0);
  bool reflect = dot(wiW, b.ng) * dot(woW, b.ng) > 0;
  RGB f = rgb(0, 0, 0);
  if (b.kind != BX_NONE && matches(b.type, flags) &&
      ((reflect && (b.type & BSDF_REFLECTION) > 0) || (!reflect && (b.type & BSDF_TRANSMISSION) > 0)))
    f = f + bxdf_f(b, wo, wi, lambert_only);
  return f;
}
GP_D double bsdf_pdf(const BSDF& b, V3 woW, V3 wiW, int flags, bool lambert_only = false) {
  if (b.kind == BX_NONE) return 0;
  V3 wo = to_local(b, woW), wi = to_local(b, wiW);
  if (wo.z == 0) return 0;
  double pdf = 0;
  int m = 0;
  if (matches(b.type, flags)) { m++; pdf += bxdf_pdf(b, wo, wi, lambert_only); }
  if (m <= 0) return 0;
  return pdf / (double)m;
}
// BSDF.SampleF (reflection.go:183-253) with matchingComps in {0,1}: returns the LOCAL wi
GP_D void bsdf_sample_f(const BSDF& b, V3 woWorld, double ux, double uy, int type, RGB* f, V3* wi, double* pdf, int* sampled, bool lambert_only = false) {
  *f = rgb(0, 0, 0); *wi = mk3(0, 0, 0); *pdf = 0; *sampled = 0;
  if (b.kind == BX_NONE || !matches(b.type, type)) return;
  double comp = go_min(floor(ux * 1.0), 1.0 - 1);
  double urx = go_min(ux * 1.0 - comp, one_minus_epsilon());
  V3 wo = to_local(b, woWorld);
  if (wo.z == 0.0) return;
  RGB ff; V3 w; double p; int st = 0;
  if (lambert_only || b.kind == BX_LAMBERT || b.kind == BX_OREN_NAYAR) {  // sampleF (reflection.go:305-314): sampledType 0
    w = cosine_sample_hemisphere(urx, uy);
    if (wo.z < 0) w.z *= -1;
    p = bxdf_pdf(b, wo, w, lambert_only);
    ff = bxdf_f(b, wo, w, lambert_only);
  } else if (b.kind == BX_SPEC_REFL_NOOP) {  // reflection.go:557-562 + FresnelNoOp
    w = mk3(-wo.x, -wo.y, wo.z);
    p = 1.0;
    ff = (rgb(1.0, 1.0, 1.0) * b.r) / fabs(w.z);
  } else {  // FresnelSpecular.SampleF (reflection.go:482-523)
    double F = fr_dielectric(wo.z, 1.0, b.etaB);
    if (urx < F) {
      w = mk3(-wo.x, -wo.y, wo.z);
      ff = (b.r * F) / fabs(w.z);
      p = F;
      st = BSDF_SPECULAR | BSDF_REFLECTION;
    } else {
      bool entering = wo.z > 0;
      double etaI = entering ? 1.0 : b.etaB, etaT = entering ? b.etaB : 1.0;
      V3 n = faceforward(mk3(0, 0, 1), wo);  // Refract (reflection.go:106-118)
      double eta = etaI / etaT;
      double cosThetaI = dot(n, wo);
      double sin2ThetaI = go_max(0, 1 - cosThetaI * cosThetaI);
      double sin2ThetaT = eta * eta * sin2ThetaI;
      if (sin2ThetaT >= 1) return;
      double cosThetaT = sqrt(1 - sin2ThetaT);
      w = wo * -eta + n * (eta * cosThetaI - cosThetaT);
      RGB ft = b.t * (1 - F);
      ft = ft * ((etaI * etaI) / (etaT / etaT));  // sic
      ff = ft / fabs(w.z);
      p = 1 - F;
      st = BSDF_SPECULAR | BSDF_TRANSMISSION;
    }
  }
  if (p == 0.0) return;
  *f = ff; *wi = w; *pdf = p; *sampled = st;
}

// NewOrenNayar (reflection.go:616-626), sigma in degrees; `0.45 * s2 / (s2 * 0.09)` as written
GP_D void oren_nayar_init(BSDF* b, double sigma_deg) {
  b->kind = BX_OREN_NAYAR;
  double s = kPi / 180.0 * sigma_deg;
  double s2 = s * s;
  b->a = 1.0 - (s2 / (2.0 * (s2 + 0.33)));
  b->b = 0.45 * s2 / (s2 * 0.09);
}

// Material.ComputeScatteringFunctions (matte.go:21-37, mirror.go:21-32, glass.go:27-75) + NewBSDF (reflection.go:128-140)
// lambert_only (a compile-time constant at the call site): the hit's shade class says its material is MatteMaterial with
// sigma == 0 (scene build, gopbrt.cu), so only that branch is compiled into the caller
GP_D bool compute_scattering(const DevScene& sc, int prim, const Hit& h, BSDF* b, bool allowMultipleLobes = true, bool lambert_only = false) {
  int mi = sc.prims[prim].z;
  if (mi < 0) return false;  // primitive.go:73-75 panics
  const MaterialDev& m = sc.materials[mi];
  b->ns = h.ns; b->ng = h.n;
  b->ss = normalized(h.sdpdu);
  b->ts = cross(b->ns, b->ss);
  b->kind = BX_NONE; b->type = 0; b->eta = 1.0; b->a = 0; b->b = 0; b->etaB = 1.0;
  b->r = rgb(0, 0, 0); b->t = rgb(0, 0, 0);
  if (lambert_only) {
    RGB r = clamp_rgb(tex_eval(sc, m.tex_a, h), 0, d_inf());
    if (!is_black(r)) { b->type = BSDF_REFLECTION | BSDF_DIFFUSE; b->r = r; b->kind = BX_LAMBERT; }
    return true;
  }
  if (m.kind == 0) {
    RGB r = clamp_rgb(tex_eval(sc, m.tex_a, h), 0, d_inf());
    double sig = go_clamp(m.sigma, 0, 90);
    if (!is_black(r)) {
      b->type = BSDF_REFLECTION | BSDF_DIFFUSE;
      b->r = r;
      if (sig == 0) b->kind = BX_LAMBERT;
      else oren_nayar_init(b, sig);
    }
    return true;
  }
  if (m.kind == 1) {
    RGB r = clamp_rgb(tex_eval(sc, m.tex_a, h), 0.0, d_inf());
    if (!is_black(r)) { b->kind = BX_SPEC_REFL_NOOP; b->type = BSDF_REFLECTION | BSDF_DIFFUSE; b->r = r; }  // sic
    return true;
  }
  if (m.kind == 2) {
    b->eta = m.eta;
    RGB R = clamp_rgb(tex_eval(sc, m.tex_a, h), 0, 1), T = clamp_rgb(tex_eval(sc, m.tex_b, h), 0, 1);
    if (is_black(R) && is_black(T)) return true;
    if (!(m.u_rough == 0 && m.v_rough == 0)) return false;  // microfacet branch: the reference panics
    if (!allowMultipleLobes) {
      // glass.go:57-72: SpecularReflection(R, FresnelDielectric) — typed Reflection|Diffuse (reflection.go:540),
      // F = 0, Pdf = 0 — plus SpecularTransmission(T) typed Transmission|Specular.  `type` carries the reflection lobe
      // (the only one a non-specular query can match); the transmission lobe is present iff T is not black (b->t).
      b->kind = BX_GLASS_SPLIT;
      b->type = is_black(R) ? 0 : (BSDF_REFLECTION | BSDF_DIFFUSE);
      b->r = R; b->t = T; b->etaB = m.eta;
      if (is_black(R)) b->kind = is_black(T) ? BX_NONE : BX_GLASS_SPLIT;
      return true;
    }
    b->kind = BX_FRESNEL_SPECULAR;
    b->type = BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_SPECULAR;
    b->r = R; b->t = T; b->etaB = m.eta;
    return true;
  }
  return false;
}

// Distribution1D.SampleDiscrete (sampling.go:42-55) via FindInterval (pkg/math/math.go:64-80) over the uniform light
// distribution (func[i] == 1, lightdistribution.go:25-34): cdf has n + 1 entries
GP_D void sample_discrete(const double* cdf, int n, double func_int, double u, int* offset_out, double* pdf_out) {
  int size = n + 1, first = 0, len = size;
  while (len > 0) {
    int half = len >> 1, middle = first + half;
    if (cdf[middle] <= u) { first = middle + 1; len -= half + 1; }
    else len = half;
  }
  *offset_out = (int)go_clamp((double)(first - 1), 0, (double)(size - 2));
  double pdf = 0;
  if (func_int > 0) pdf = 1.0 / (func_int / (double)n);
  *pdf_out = pdf;
}

// ---------------------------------------------------------------- lights (pkg/lights, sphere.go:270-344, disk.go:160-170, shape.go:50-65)
struct Intr { V3 p, perr, n; };
GP_D void sphere_sample(const DevScene& sc, const SphereDev& s, double ux, double uy, Intr* it, double* pdf) {
  M4 m = load_m4_plain(sc, s.xf, false), inv = load_m4_plain(sc, s.xf, true);
  V3 pObj = uniform_sample_sphere(ux, uy) * s.radius;
  V3 n = normalized(xf_normal_inv(inv, pObj));
  if (s.flags & RF_REVERSE) n = n * -1.0;
  pObj = pObj * (s.radius / sqrt(dist2(pObj, mk3(0, 0, 0))));
  V3 pObjError = vabs(pObj) * gamma_n(5);
  it->p = xf_point(m, pObj, pObjError, &it->perr);
  it->n = n;
  *pdf = 1.0 / (s.phiMax * s.radius * (s.zMax - s.zMin));
}
// (inlined: as out-of-line functions their reference / pointer arguments — the shape record, the reference point, the sampled
// point — went through the local-memory stack, ~60 local loads and stores per shaded lane: ncu counted more L1<->L2 traffic
// from local memory than from the lane records)
GP_D void sphere_sample_at(const DevScene& sc, const SphereDev& s, const Intr& ref, double ux, double uy, Intr* it, double* pdf) {
  M4 m = load_m4_plain(sc, s.xf, false);
  V3 pCenter = xf_point(m, mk3(0, 0, 0), mk3(0, 0, 0), nullptr);
  V3 pOrigin = offset_ray_origin(ref.p, ref.perr, ref.n, pCenter - ref.p);
  if (dist2(pOrigin, pCenter) <= s.radius * s.radius) {
    double p;
    sphere_sample(sc, s, ux, uy, it, &p);
    V3 wi = it->p - ref.p;
    if (len2(wi) == 0) p = 0;
    else {
      wi = normalized(wi);
      p *= dist2(ref.p, it->p) / fabs(dot(it->n, wi * -1.0));
    }
    if (is_inf(p)) p = 0.0;
    *pdf = p;
    return;
  }
  V3 wc = normalized(pCenter - ref.p);
  V3 wcX, wcY;
  coordinate_system(wc, &wcX, &wcY);
  double radius2 = s.radius * s.radius;
  double sinThetaMax2 = radius2 / dist2(ref.p, pCenter);
  double cosThetaMax = sqrt(go_max(0, 1.0 - sinThetaMax2));
  double cosTheta = (1.0 - ux) + ux * cosThetaMax;
  double sinTheta = sqrt(go_max(0, 1 - cosTheta * cosTheta));
  double phi = uy * 2 * kPi;
  double dc = sqrt(dist2(ref.p, pCenter));
  double ds = dc * cosTheta - sqrt(go_max(0, radius2 - (dc * dc) * (sinTheta * sinTheta)));
  double cosAlpha = (dc * dc + radius2 - ds * ds) / (2.0 * dc * s.radius);
  double sinAlpha = sqrt(go_max(0, 1.0 - cosAlpha * cosAlpha));
  V3 x = wcX * -1.0, y = wcY * -1.0, z = wc * -1.0;
  const SinCos scPhi = go_sincos(phi);
  V3 nWorld = x * (sinAlpha * scPhi.cs) + y * (sinAlpha * scPhi.sn) + z * cosAlpha;  // geometry.go:66-70
  V3 pWorld = pCenter + nWorld * s.radius;
  it->p = pWorld;
  it->perr = vabs(pWorld) * gamma_n(5.0);
  it->n = nWorld;
  if (s.flags & RF_REVERSE) it->n = it->n * -1.0;
  *pdf = 1.0 / (2.0 * kPi * (1.0 - cosThetaMax));  // UniformConePdf (sampling.go:169-171)
}
GP_D void disk_sample_at(const DevScene& sc, const DiskDev& d, const Intr& ref, double ux, double uy, Intr* it, double* pdf) {
  M4 m = load_m4_plain(sc, d.xf, false), inv = load_m4_plain(sc, d.xf, true);
  double px, py;
  concentric_sample_disk(ux, uy, &px, &py);
  V3 pObj = mk3(px * d.radius, py * d.radius, d.height);
  V3 n = xf_normal_inv(inv, mk3(0, 0, 1));
  if (d.flags & RF_REVERSE) n = n * -1.0;
  it->n = n;
  it->p = xf_point(m, pObj, mk3(0, 0, 0), &it->perr);
  double p = 1 / (d.phiMax * 0.5 * (d.radius * d.radius - d.innerRadius * d.innerRadius));
  V3 wi = it->p - ref.p;
  if (len2(wi) == 0.0) { *pdf = 0; return; }
  wi = normalized(wi);
  p *= dist2(ref.p, it->p) / fabs(dot(it->n, wi * -1.0));
  if (is_inf(p)) p = 0;
  *pdf = p;
}

// SpawnRayToInteraction (interaction.go:91-102): the direction of the visibility segment; its origin stays the
// UN-offset point and tMax = 1 - ShadowEpsilon
GP_D V3 spawn_ray_to(const Intr& from, const Intr& to) {
  V3 origin = offset_ray_origin(from.p, from.perr, from.n, to.p - from.p);
  V3 target = offset_ray_origin(to.p, to.perr, to.n, origin - to.p);
  return target - origin;
}

struct LightSample { RGB Li; V3 wi; double pdf; Intr p1; bool delta; };
// Point.SampleLi (point.go:44-49), Distant.SampleLi (distant.go:40-44), DiffuseAreaLight.SampleLi (diffuse.go:47-59)
GP_D void light_sample_li(const DevScene& sc, const LightDev& l, const Intr& ref, double ux, double uy, LightSample* ls) {
  RGB E = rgb(l.rgb[0], l.rgb[1], l.rgb[2]);
  V3 v = mk3(l.v[0], l.v[1], l.v[2]);
  V3 zero = mk3(0, 0, 0);
  if (l.kind == 1) {
    ls->wi = normalized(v - ref.p);
    ls->pdf = 1.0;
    ls->p1.p = v; ls->p1.perr = zero; ls->p1.n = zero;
    ls->Li = E / dist2(v, ref.p);
    ls->delta = true;
  } else if (l.kind == 0) {
    ls->p1.p = v * (2 * sc.world_radius); ls->p1.perr = zero; ls->p1.n = zero;  // sic: a fixed point
    ls->Li = E; ls->wi = v; ls->pdf = 1; ls->delta = true;
  } else {
    ls->delta = false;
    Intr ps;
    double pdf;
    if (l.shape_kind == RK_SPHERE) sphere_sample_at(sc, sc.spheres[l.shape_index], ref, ux, uy, &ps, &pdf);
    else disk_sample_at(sc, sc.disks[l.shape_index], ref, ux, uy, &ps, &pdf);
    if (pdf == 0 || len2(ps.p - ref.p) == 0) { ls->Li = rgb(0, 0, 0); ls->wi = zero; ls->pdf = 0; ls->p1 = ps; return; }
    ls->wi = ps.p - ref.p;  // un-normalised (diffuse.go:55)
    ls->p1 = ps;
    ls->pdf = pdf;
    V3 w = ls->wi * -1.0;
    ls->Li = (l.two_sided || dot(ps.n, w) > 0) ? E : rgb(0, 0, 0);
  }
}

// RGBToXYZ (spectrum.go:35-41)
GP_D void rgb_to_xyz(double r, double g, double b, double* X, double* Y, double* Z) {
  *X = 0.412453 * r + 0.357580 * g + 0.180423 * b;
  *Y = 0.212671 * r + 0.715160 * g + 0.072169 * b;
  *Z = 0.019334 * r + 0.119193 * g + 0.950227 * b;
}

// ---------------------------------------------------------------- film geometry (film.go:106-113)
GP_D void tile_bounds(const RenderParams& P, long long tile, long long* x0, long long* y0, long long* x1, long long* y1) {
  long long ty = div_nn(tile, P.ntx), tx = tile - ty * P.ntx;
  *x0 = P.cx0 + tx * P.tile_size;
  *x1 = (long long)go_min((double)(*x0 + P.tile_size), (double)P.cx1);  // integrator.go:322-325
  *y0 = P.cy0 + ty * P.tile_size;
  *y1 = (long long)go_min((double)(*y0 + P.tile_size), (double)P.cy1);
}
GP_D void tile_pixel_bounds(const RenderParams& P, long long x0, long long y0, long long x1, long long y1, long long* px0, long long* py0,
                            long long* px1, long long* py1) {
  long long a = (long long)ceil((double)x0 - 0.5 - P.frx), b = (long long)ceil((double)y0 - 0.5 - P.fry);
  long long c = (long long)floor((double)x1 - 0.5 + P.frx) + 1, d = (long long)floor((double)y1 - 0.5 + P.fry) + 1;
  *px0 = a > P.cx0 ? a : P.cx0; *py0 = b > P.cy0 ? b : P.cy0;
  *px1 = c < P.cx1 ? c : P.cx1; *py1 = d < P.cy1 ? d : P.cy1;
}

// FilmTile.AddSample (film.go:211-248) with the box filter table (all ones, filter.go:30-32), sampleWeight 1
GP_D void film_add_sample(const Lanes& L, const RenderParams& P, long long lane, long long tile, double fx, double fy, RGB Lc) {
  long long x0, y0, x1, y1, bx0, by0, bx1, by1;
  tile_bounds(P, tile, &x0, &y0, &x1, &y1);
  tile_pixel_bounds(P, x0, y0, x1, y1, &bx0, &by0, &bx1, &by1);
  double dx = fx - 0.5, dy = fy - 0.5;
  double p0fx = ceil(dx - P.frx), p0fy = ceil(dy - P.fry);
  double p1fx = floor(dx + P.frx) + 1, p1fy = floor(dy + P.fry) + 1;
  long long p0x = (long long)go_max(p0fx, (double)bx0), p0y = (long long)go_max(p0fy, (double)by0);
  long long p1x = (long long)go_min(p1fx, (double)bx1), p1y = (long long)go_min(p1fy, (double)by1);
  // contribSum += L*w*f with L == +0 leaves the (never negative-zero) sums untouched: only the weight moves
  bool zero = Lc.r == 0 && Lc.g == 0 && Lc.b == 0 && !sign_bit(Lc.r) && !sign_bit(Lc.g) && !sign_bit(Lc.b);
  double* tile_px = L.tilepix + (size_t)lane * L.tile_stride;
  if (p1x - p0x == 2 && p1y - p0y == 2 && P.tpw == 2 && p0x == bx0 && p0y == by0) {
    // The usual footprint: the lane's whole 2 x 2 FilmTile = one 128-byte line.  All of it is read before any of it is
    // written: a store into the line would otherwise push the next pixel's load back out to L2, one serialised round
    // trip per pixel.  The per-pixel sums are independent, so the order of the accesses changes no bit.
    double2* qp = (double2*)tile_px;
    double2 v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = qp[i];
    const double fw = 1.0;
    if (!zero) {
      RGB c = Lc * (1.0 * fw);
#pragma unroll
      for (int i = 0; i < 4; i++) { v[2 * i].x += c.r; v[2 * i].y += c.g; v[2 * i + 1].x += c.b; }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) v[2 * i + 1].y += fw;
#pragma unroll
    for (int i = 0; i < 8; i++)
      if (!zero || (i & 1)) qp[i] = v[i];
    return;
  }
  for (long long y = p0y; y < p1y; y++)
    for (long long x = p0x; x < p1x; x++) {
      double fw = 1.0;
      size_t k = (size_t)((y - by0) * P.tpw + (x - bx0)) * 4;
      double* q = tile_px + k;
      if (!zero) {
        RGB c = Lc * (1.0 * fw);
        q[0] += c.r;
        q[1] += c.g;
        q[2] += c.b;
      }
      q[3] += fw;
    }
}

// Uniform-footprint renders (tileSize 1, Stratified sampler with at least one sampled dimension): every sample of a lane
// goes through the integer corner of the lane's ONE pixel (StratifiedSample2D never writes its table) with
// filter weight 1 (box filter), so all pixels of its FilmTile that a sample touches receive the very same sequence of
// additions `contribSum += L` and `filterWeightSum += 1` (film.go:241-243).  The tile is then ONE running RGB sum, kept
// in the PathRec's spare 24 bytes (it is in registers whenever a sample retires), and the weight is the number of samples
// the lane has retired — known in closed form.  Saves the 256-byte tile read-modify-write per sample.
GP_D void film_add_uniform(FilmRec& pt, RGB Lc) {
  pt.pad[0] += Lc.r * (1.0 * 1.0);  // L.MulScalar(sampleWeight * filterWeight)
  pt.pad[1] += Lc.g * (1.0 * 1.0);
  pt.pad[2] += Lc.b * (1.0 * 1.0);
}
// the pixels [p0, p1) a sample at the integer corner (px, py) touches inside its tile's pixel bounds (film.go:217-221)
GP_D void uniform_footprint(const RenderParams& P, long long px, long long py, long long bx0, long long by0, long long bx1, long long by1,
                            long long* p0x, long long* p0y, long long* p1x, long long* p1y) {
  double dx = (double)px - 0.5, dy = (double)py - 0.5;
  *p0x = (long long)go_max(ceil(dx - P.frx), (double)bx0); *p0y = (long long)go_max(ceil(dy - P.fry), (double)by0);
  *p1x = (long long)go_min(floor(dx + P.frx) + 1, (double)bx1); *p1y = (long long)go_min(floor(dy + P.fry) + 1, (double)by1);
}

// ---------------------------------------------------------------- raygen + sampler
// PerspectiveCamera.GenerateRayDifferential (camera.go:192-242), the ray itself (Path.Li drops the differentials)
GP_D Ray camera_ray(const RenderParams& P, double fx, double fy, double lx, double ly) {
  V3 pCamera = xf_point(P.raster_to_camera, mk3(fx, fy, 0), mk3(0, 0, 0), nullptr);
  Ray ray;
  ray.o = mk3(0, 0, 0);
  ray.d = normalized(pCamera);
  ray.tmax = d_inf();
  if (P.lens_radius > 0) {
    double plx, ply;
    concentric_sample_disk(lx, ly, &plx, &ply);
    plx *= P.lens_radius; ply *= P.lens_radius;
    double ft = P.focal_distance / ray.d.z;
    V3 pFocus = ray.d * ft + ray.o;
    ray.o = mk3(plx, ply, 0);
    ray.d = normalized(pFocus - ray.o);
  }
  return xf_ray(P.camera_to_world, ray, nullptr, nullptr);
}

// One lane's raygen step: retires the lane's finished sample into its film tile (renderWorker, integrator.go:252-265),
// advances the sampler (StartNextSample / next pixel + StartPixel) and generates the next camera ray
// (GenerateRayDifferential, camera.go:192-242; the differentials are dropped by Path.Li).  Returns false when the
// lane's tile is exhausted.
struct PathRec;
GP_D PathRec initial_path(const RenderParams& P, long long lane);  // defined below
GP_D RGB direct_unwind(const Lanes& L, const RenderParams& P, long long lane, const PathRec& pt, const RadRec& rd, int seg_mask);  // DirectLighting, defined below
GP_D bool generate_lane(const DevScene& sc, const Lanes& L, const RenderParams& P, long long lane, bool have_sample,
                        unsigned long long& cam, unsigned long long& nans, unsigned long long& culled) {
  bool go = false;
  const long long slot = P.lane_base + lane, slot_tile = P.groups == 1 ? slot : div_nn(slot, P.groups);
  long long tile = slot_tile * P.world + P.rank;
  const int s_mod = P.s_world * P.groups, s_res = P.s_rank * P.groups + (int)(slot - slot_tile * P.groups);
  PathRec pt = have_sample ? L.path[lane] : initial_path(P, lane);  // the pass's first launch starts every lane from scratch
  RadRec rd;
  rd.Lr = 0; rd.Lg = 0; rd.Lb = 0; rd.eta_scale = 1.0;
  FilmRec fm;
  fm.pix = -1; fm.has_sample = 0; fm.pad[0] = fm.pad[1] = fm.pad[2] = 0;
  if (have_sample) {  // every lane of the regeneration queue carries a finished sample
    rd = L.rad[lane];
    fm = L.fsum[lane];
    RGB Lc = P.integrator == 1 ? direct_unwind(L, P, lane, pt, rd, fm.has_sample) : rgb(rd.Lr, rd.Lg, rd.Lb);
    if (is_nan(Lc.r) || is_nan(Lc.g) || is_nan(Lc.b)) { Lc = rgb(0.1, 0.1, 0.1); nans++; }  // integrator.go:256-257
    if (P.uniform_fp) film_add_uniform(fm, Lc);
    else film_add_sample(L, P, lane, tile, pt.fx, pt.fy, Lc);
  }
  Smp s;
  s.state = pt.rng_state; s.inc = pt.rng_inc; s.sidx = pt.sidx; s.cur1 = 0; s.cur2 = 0; s.lane = lane;
  int pix = fm.pix;
  long long x0, y0, x1, y1;
  tile_bounds(P, tile, &x0, &y0, &x1, &y1);
  long long tw = x1 - x0, area = tw * (y1 - y0);
  // root bound of the BVH (node 0), exactly what the extend stage would test first
  float4 rn0 = make_float4(0, 0, 0, 0), rn1 = make_float4(0, 0, 0, 0);
  if (sc.n_nodes > 0) { rn0 = __ldg(sc.nodes); rn1 = __ldg(sc.nodes + 1); }
  for (;;) {  // samples whose camera ray misses the root bound are finished on the spot (see below)
    bool have = false;
    for (;;) {
      if (pix >= 0) {
        // PixelSampler.StartNextSample (pixel.go:48-52) + Sampler.StartNextSample (sampler.go:29-34): increments first
        s.cur1 = 0; s.cur2 = 0;
        s.sidx += 1;
        if (s_mod > 1) {
          // FAST: samples split by index (over ranks, then lane groups) — step straight to this lane's next sample.
          // (Searching for it one index at a time let the lanes of a warp leave the search at different iterations and
          // run the whole camera-ray body one lane group at a time: 4 of 32 lanes active in a pass's first launch.)
          int d = s_res - s.sidx % s_mod;
          if (d < 0) d += s_mod;
          s.sidx += d;
        }
        if (s.sidx < P.spp) {
          have = true;
          break;
        }
      }
      pix++;
      if (pix >= area) break;
      start_pixel(s, L, P);
    }
    if (!have) break;
    long long prow = tw == 1 ? pix : div_nn(pix, tw);
    long long px = x0 + (pix - prow * tw), py = y0 + prow;
    unsigned long long fast_pixel = (unsigned long long)((py - P.cy0) * (P.cx1 - P.cx0) + (px - P.cx0));
    if (P.mode == 1) rng_set_sequence(s, fast_pixel * (unsigned long long)P.spp + (unsigned long long)s.sidx);
    // GetCameraSample (sampler.go:75-80): Get2D pFilm, Get2D pLens, Get1D time
    double ox, oy, lx, ly;
    get2d(s, P, &ox, &oy);
    double fx = (double)px + ox, fy = (double)py + oy;
    get2d(s, P, &lx, &ly);
    double time = get1d(s, L, P, fast_pixel);
    Ray ray = camera_ray(P, fx, fy, lx, ly);
    (void)time;  // ray.Time = Lerp(time, open, open): unused without animated transforms
    cam++;
    // BVH.Intersect's first step (bvh.go:673-675): the root node's slab test.  A camera ray that fails it hits
    // nothing, Path.Li returns L = 0 after that one scene.Intersect query (path.go:45,66) and draws no further
    // samples, so the sample is added to the film right here instead of travelling through extend and back.
    V3 invd = mk3(1 / ray.d.x, 1 / ray.d.y, 1 / ray.d.z);
    bool enters = sc.n_nodes > 0 && slab_test((double)rn0.x, (double)rn0.y, (double)rn0.z, (double)rn1.x, (double)rn1.y, (double)rn1.z,
                                               ray.o, invd, invd.x < 0, invd.y < 0, invd.z < 0, ray.tmax);
    if (!enters) {
      culled++;
      if (!P.uniform_fp) film_add_sample(L, P, lane, tile, fx, fy, rgb(0, 0, 0));  // (uniform footprint: L = +0 adds nothing, the weight is counted)
      continue;
    }
    RayRec rr;
    rr.ox = ray.o.x; rr.oy = ray.o.y; rr.oz = ray.o.z; rr.dx = ray.d.x; rr.dy = ray.d.y; rr.dz = ray.d.z;
    rr.tmax = d_inf(); rr.hit_rec = -1; rr.pad = lane_on_last_sample(P, s.sidx) ? 1 : 0;  // see k_split_hits
    L.ray[lane] = rr;
    pt.fx = fx; pt.fy = fy;
    rd.Lr = 0; rd.Lg = 0; rd.Lb = 0;
    pt.br = 1.0; pt.bg = 1.0; pt.bb = 1.0;
    rd.eta_scale = 1.0;
    pt.bounces = (s.cur1 << 8) | (s.cur2 << 16);  // bounces in bits 0-7, sampler dimensions above
    fm.has_sample = 0;  // DirectLighting / UniformSampleAll: no shadow segments pending
    go = true;
    break;
  }
  pt.rng_state = s.state; pt.rng_inc = s.inc; pt.sidx = s.sidx; fm.pix = pix;
  if (P.last_in_place && !go) { rd.Lr = 0; rd.Lg = 0; rd.Lb = 0; }  // no sample in flight: the film fold adds pad + L
  L.path[lane] = pt;
  L.rad[lane] = rd;
  L.fsum[lane] = fm;
  return go;
}


// Retires the lane's finished sample into its film tile (renderWorker, integrator.go:252-265), advances the sampler
// (StartNextSample / next pixel + StartPixel) and generates the next camera ray (GenerateRayDifferential,
// camera.go:192-242; the differentials are dropped by Path.Li).  Lanes whose tile is exhausted leave the wavefront.
// MODE / INTEG: the sampler mode and integrator kind as compile-time constants (the launch picks the instantiation): the
// FAST-mode raygen then carries none of Stratified.StartPixel's table code and the Path raygen none of the DirectLighting
// unwind, which is what sets the register budget — and with it the number of resident warps that hide the three dependent
// HBM round trips (queue -> PathRec -> FilmTile) of this stage.
#ifndef GP_GEN_BLOCKS
#define GP_GEN_BLOCKS 4
#endif
template <int MODE, int INTEG, int UFP>
__global__ void __launch_bounds__(128, GP_GEN_BLOCKS) k_generate(DevScene sc, Lanes L, RenderParams P_in, Queues Q, const int* __restrict__ in_queue,
                                                  const int* __restrict__ in_count, RenderCounters* ctr) {
  RenderParams P = P_in;
  P.mode = MODE;
  P.integrator = INTEG;
  P.uniform_fp = UFP;
  long long n = in_queue ? (long long)*in_count : P.lanes_active;
  int lane_id = threadIdx.x & 31;
  unsigned long long cam = 0, nans = 0, culled = 0;
  __shared__ PushBatch<1, kPushBatch> s_push;
  for (long long cbase = (long long)blockIdx.x * blockDim.x * kPushBatch; cbase < n; cbase += (long long)gridDim.x * blockDim.x * kPushBatch) {
    unsigned flags = 0;
#pragma unroll 1
    for (int u = 0; u < kPushBatch; u++) {
      const long long i = cbase + (long long)u * blockDim.x + threadIdx.x;
      if (i < n) {
        const long long lane = in_queue ? in_queue[i] : i;
        if (generate_lane(sc, L, P, lane, in_queue != nullptr, cam, nans, culled)) flags |= 1u << u;
        s_push.val[u][threadIdx.x] = (int)lane;
      }
    }
    int* const qs[1] = {Q.extend};
    int* const cs[1] = {Q.cnt + 0};
    block_push_batch<1, kPushBatch>(qs, cs, flags, s_push);
  }
  cam = warp_sum(cam); nans = warp_sum(nans); culled = warp_sum(culled);
  if (lane_id == 0) {
    if (cam) atomicAdd(&ctr->camera_rays, cam);
    if (nans) atomicAdd(&ctr->nan_samples, nans);
    if (culled) atomicAdd(&ctr->root_culled, culled);
  }
}

// ---------------------------------------------------------------- split the extend queue by outcome
// A ray that escaped the scene ends its path (path.go:66): its lane goes straight to the regeneration queue; only real
// hits reach the shade stage.  Order-preserving warp-ballot compaction (the queue order is roughly pixel order, which
// keeps the rays of a warp coherent in the next extend).
__global__ void __launch_bounds__(256) k_split_hits(Lanes L, Queues Q, const unsigned char* __restrict__ codes) {
  long long n = Q.cnt[0];
  int lane_id = threadIdx.x & 31;
  // one atomicAdd per CTA and output queue instead of one per warp: the five counters are single addresses, and at
  // several million lanes per launch the per-warp atomics on them serialise
  __shared__ int s_cnt[8][5];
  __shared__ int s_base[5];
  const int warp = threadIdx.x >> 5;
  for (long long cbase = (long long)blockIdx.x * blockDim.x; cbase < n; cbase += (long long)gridDim.x * blockDim.x) {
    long long i = cbase + threadIdx.x;
    bool valid = i < n;
    int lane = 0, rec = -1, cls = 0;
    if (valid) {
      lane = Q.extend[i];
      if (codes) { int c = codes[i]; rec = (c & 7) < 4 ? 0 : -1; cls = c; }  // filed by the extend kernel under the queue position
      else {
        int2 rc = *(const int2*)&L.ray[lane].hit_rec;  // {hit_rec, shade class}
        rec = rc.x; cls = rc.y;
      }
    }
    // an escaped ray ends its sample; bit 3 = it was the lane's last sample, which stays in place (lane_on_last_sample)
    int bin = !valid ? -1 : (rec >= 0 ? (cls & 3) : ((cls & 8) ? -1 : 4));
    unsigned m[5];
#pragma unroll
    for (int k = 0; k < 5; k++) { m[k] = __ballot_sync(0xffffffffu, bin == k); if (lane_id == 0) s_cnt[warp][k] = __popc(m[k]); }
    __syncthreads();
    if (threadIdx.x < 5) {
      int tot = 0;
      for (int w = 0; w < 8; w++) { const int c = s_cnt[w][threadIdx.x]; s_cnt[w][threadIdx.x] = tot; tot += c; }  // counts -> offsets
      int* cnt = threadIdx.x < 4 ? Q.cnt + 8 + threadIdx.x : Q.cnt + 4;
      s_base[threadIdx.x] = tot ? atomicAdd(cnt, tot) : 0;
    }
    __syncthreads();
    if (bin >= 0) {
      const int off = s_base[bin] + s_cnt[warp][bin];
      unsigned mm = bin == 0 ? m[0] : (bin == 1 ? m[1] : (bin == 2 ? m[2] : (bin == 3 ? m[3] : m[4])));
      int* q = bin < 4 ? Q.shade[bin] : Q.regen_next;
      q[off + __popc(mm & ((1u << lane_id) - 1u))] = lane;
    }
    __syncthreads();
  }
}

// FAST mode: the index of the pixel a lane is working on (row-major in the cropped window) — what keys the sample's counter-
// based streams.  With tileSize 1 the tile IS the pixel (ntx = the window's width), so the index is the tile number and the
// shade stage is spared tile_bounds and two of its three integer divisions per lane.
GP_D unsigned long long lane_fast_pixel(const RenderParams& P, long long lane, int pix) {
  const long long tile = (P.groups == 1 ? P.lane_base + lane : div_nn(P.lane_base + lane, P.groups)) * P.world + P.rank;
  if (P.tile_size == 1) return (unsigned long long)tile;
  long long x0, y0, x1, y1;
  tile_bounds(P, tile, &x0, &y0, &x1, &y1);
  const long long prow = div_nn(pix, x1 - x0);
  const long long px = x0 + (pix - prow * (x1 - x0)), py = y0 + prow;
  return (unsigned long long)((py - P.cy0) * (P.cx1 - P.cx0) + (px - P.cx0));
}

// ---------------------------------------------------------------- shade
// One Path.Li loop body for one lane (path.go:40-155) after its closest-hit query.  Sets cont (the path continues with a
// new ray in L.ray[lane]), finished (the sample is complete) and shadow (a visibility segment is pending in L.sray[lane]).
// cls: the hit's shade class (bit 0 = material is not plain Lambert, bit 1 = sphere / disk hit) when the caller shades one
// class only (a compile-time constant after inlining: the other classes' code is not compiled in), or -1 = any
GP_D void shade_lane(const DevScene& sc, const Lanes& L, const RenderParams& P, long long lane, bool& cont, bool& finished, bool& shadow,
                     unsigned& n_counts, int& bad, const int cls = -1) {
  const bool lambert_only = cls >= 0 && !(cls & 1), tri_only = cls >= 0 && !(cls & 2);
  PathRec pt = L.path[lane];
  RayRec rr = L.ray[lane];
  int packed = pt.bounces;
  int bounces = (packed & 255) + 1;  // bounces++ (path.go:41)
  int rec = rr.hit_rec;
  bool w_eta = false, w_beta = false, w_rng = false;
  double eta_new = 0;
  finished = true;
  if (rec >= 0 && bounces < P.max_depth) {  // path.go:66
    Ray ray;
    ray.o = mk3(rr.ox, rr.oy, rr.oz);
    ray.d = mk3(rr.dx, rr.dy, rr.dz);
    ray.tmax = rr.tmax;
    Hit h;
    int prim;
    hit_record(sc, rec, ray, ray.tmax, &h, &prim, bad, tri_only);
    BSDF bsdf;
    if (!compute_scattering(sc, prim, h, &bsdf, true, lambert_only)) {
      n_counts += 0x10000u;
    } else {
      Smp s;
      s.state = pt.rng_state; s.inc = pt.rng_inc; s.sidx = pt.sidx;
      s.cur1 = (packed >> 8) & 255; s.cur2 = (packed >> 16) & 255; s.lane = lane;
      const unsigned long long fast_pixel = P.mode == 1 ? lane_fast_pixel(P, lane, P.tile_size == 1 ? 0 : L.fsum[lane].pix) : 0ULL;
      RGB beta = rgb(pt.br, pt.bg, pt.bb);
      Intr ref; ref.p = h.p; ref.perr = h.perr; ref.n = h.n;
      // --- UniformSampleOneLight (integrator.go:48-77), skipped for perfectly specular BSDFs (path.go:84)
      if (bsdf.kind != BX_NONE && matches(bsdf.type, BSDF_ALL & ~BSDF_SPECULAR)) {
        if (sc.n_lights > 0) {
          double u = get1d(s, L, P, fast_pixel);
          int offset;
          double lightPdf;
          sample_discrete(sc.light_cdf, sc.n_lights, sc.light_func_int, u, &offset, &lightPdf);
          // LightSampleStrategy Power: ComputeLightPowerDistribution (lightdistribution.go:57-68) APPENDS the powers to a
          // slice it already made n long, and every power is Spectrum.Y() == 0 (spectrum.go:227-229): 2n zeros, FuncInt
          // == 0, so SampleDiscrete's pdf is 0 and UniformSampleOneLight returns black right after this Get1D
          if (P.light_power) lightPdf = 0;
          if (lightPdf != 0.0) {
            double ulx, uly, usx, usy;
            get2d(s, P, &ulx, &uly);
            get2d(s, P, &usx, &usy);  // uScattering: drawn, used only by the dead MIS branch
            // --- EstimateDirect (integrator.go:79-195), handleMedia = false, specular = false
            const int flags = BSDF_ALL & ~BSDF_SPECULAR;
            LightSample ls;
            light_sample_li(sc, sc.lights[offset], ref, ulx, uly, &ls);
            if (!ls.delta) n_counts += 1u;
            if (ls.pdf > 0 && !is_black(ls.Li)) {
              RGB f = bsdf_f(bsdf, h.wo, ls.wi, flags, lambert_only);
              f = f * fabs(dot(ls.wi, h.ns));
              double scatteringPdf = bsdf_pdf(bsdf, h.wo, ls.wi, flags, lambert_only);
              if (!is_black(f)) {
                RGB Ld;
                if (ls.delta) Ld = (f * ls.Li) / ls.pdf;
                else {
                  double ff = 1.0 * ls.pdf, gg = 1.0 * scatteringPdf;  // PowerHeuristic (sampling.go:208-212)
                  double weight = (ff * ff) / (ff * ff + gg * gg);
                  Ld = ((f * ls.Li) * weight) / ls.pdf;
                }
                Ld = rgb(0, 0, 0) + Ld;  // Ld.AddAssign on a zero spectrum (integrator.go:123-126)
                // VisibilityTester.Unoccluded -> SpawnRayToInteraction (interaction.go:91-102)
                V3 d = spawn_ray_to(ref, ls.p1);
                RGB c = beta * Ld;  // Ld := beta.Mul(...) (path.go:85)
                ShadowRec sr;
                sr.ox = ref.p.x; sr.oy = ref.p.y; sr.oz = ref.p.z; sr.dx = d.x; sr.dy = d.y; sr.dz = d.z;  // tMax = 1 - ShadowEpsilon
                sr.pr = c.r; sr.pg = c.g; sr.pb = c.b;
                sr.gt10 = max_comp(Ld) > 10 ? 1 : 0;  // integrator.go:73-75 panics; counted when unoccluded
                sr.pad = 0; sr.pad2[0] = 0; sr.pad2[1] = 0;
                L.sray[lane] = sr;
                shadow = true;
              }
            }
          }
        }
        if (!shadow) {
          // L.AddAssign(beta.Mul(0)) (path.go:85-86): a no-op unless beta is not finite (0*Inf = NaN), kept for parity
          // Adding a zero of either sign changes no radiance sum (the sum is never -0: it starts at +0), so the record's
          // radiance is only touched — in place — when the product is NOT a zero, i.e. NaN.
          RGB z = beta * rgb(0, 0, 0);
          if (!(z.r == 0 && z.g == 0 && z.b == 0)) { RadRec* q = L.rad + lane; q->Lr += z.r; q->Lg += z.g; q->Lb += z.b; }
        }
      }
      // --- sample the BSDF for the next direction (path.go:90-117); wo = ray.Direction, sic
      double ux, uy;
      get2d(s, P, &ux, &uy);
      RGB f; V3 wi; double pdf; int sflags;
      bsdf_sample_f(bsdf, ray.d, ux, uy, BSDF_ALL, &f, &wi, &pdf, &sflags, lambert_only);
      if (!(is_black(f) || pdf == 0.0)) {
        double wiAbsDotPdf = fabs(dot(wi, h.ns)) / pdf;
        beta = beta * (f * wiAbsDotPdf);
        double etaScale = L.rad[lane].eta_scale;
        if ((sflags & BSDF_SPECULAR) > 0 && (sflags & BSDF_TRANSMISSION) > 0) {
          double eta = bsdf.eta;
          if (dot(ray.d, h.n) > 0) etaScale *= eta * eta;
          else etaScale *= 1 / (eta * eta);
          eta_new = etaScale;
          w_eta = true;
        }
        V3 o = offset_ray_origin(h.p, h.perr, h.n, wi);  // SpawnRay (interaction.go:68-77); wi is BSDF-local
        bool alive = true;
        RGB rrBeta = beta * etaScale;
        if (max_comp(rrBeta) < P.rr_threshold && bounces > 3) {  // path.go:145-153
          double q = go_max(0.05, 1 - max_comp(rrBeta));
          if (get1d(s, L, P, fast_pixel) < q) alive = false;
          else beta = beta / (1 - q);
        }
        if (alive) {
          RayRec nr;
          nr.ox = o.x; nr.oy = o.y; nr.oz = o.z; nr.dx = wi.x; nr.dy = wi.y; nr.dz = wi.z;
          nr.tmax = d_inf(); nr.hit_rec = -1; nr.pad = lane_on_last_sample(P, pt.sidx) ? 1 : 0;
          L.ray[lane] = nr;
          pt.br = beta.r; pt.bg = beta.g; pt.bb = beta.b;
          w_beta = true;
          cont = true;
          finished = false;
        }
      }
      pt.rng_state = s.state;  // (the stream's increment does not change inside a sample)
      w_rng = true;
      packed = (s.cur1 << 8) | (s.cur2 << 16);
    }
  }
  // Only what this loop body changed goes back: the bounce count and sampler dimensions always, the sampler state and
  // (if the path goes on) the throughput, rarely the refraction scale (the radiance sum is the shadow stage's to add to).
  // pFilm, the stream increment, the pixel and the film sum are the raygen stage's; not keeping them for a whole-record
  // store frees registers for the float64 chain (and the record's fourth sector is not written at all).
  PathRec* const pp = L.path + lane;
  pp->bounces = (packed & ~255) | bounces;
  if (w_rng) pp->rng_state = pt.rng_state;
  if (w_beta) { pp->br = pt.br; pp->bg = pt.bg; pp->bb = pt.bb; }
  if (w_eta) L.rad[lane].eta_scale = eta_new;
  if (finished && lane_on_last_sample(P, pt.sidx)) finished = false;  // stays in place: no trip through the regeneration queue
}

// ---------------------------------------------------------------- DirectLighting (pkg/integrator/directlighting.go)
// DirectLighting.Li (directlighting.go:62-104) is a recursion: L = direct(si) + SpecularReflect + SpecularTransmit, each
// specular term being f * Li(child) * |wi.ns| / pdf (integrator.go:352-422) with the child entered two depth units down.
// With the reference's BxDF typing the Reflection|Specular query never matches anything (SpecularReflection is typed
// Reflection|Diffuse) and the Transmission|Specular query matches only smooth glass, so the recursion is a
// CHAIN.  The wavefront walks it forwards, one level per iteration, and keeps per level the frame {D, f, c}: D = the
// level's direct light (it arrives through the shadow stage, so it is filed when the next level starts or when the
// sample retires), f and c = the specular term's factors.  The sample's radiance is the chain unwound backwards,
// L_j = D_j + (f_j * L_{j+1}) * c_j — the very additions and multiplications of the recursion, in its order.
// The ray differentials the reference threads through feed nothing these materials and textures read.
constexpr int kDirectHitBit = 1 << 24;  // PathRec.bounces: the chain ended AT a hit (its D is still in pt.L at retire)

// UniformSampleAll: the level's direct light is the sum over the lights, in light order, of the segments the shadow
// stage found unoccluded (L.AddAssign(EstimateDirect(...)) per light, integrator.go:23-46); pt.has_sample holds the
// mask of the segments that were emitted.  UniformSampleOne: the shadow stage has already added its single segment.
GP_D void direct_collect(const Lanes& L, const RenderParams& P, long long lane, int& seg_mask_io, RadRec& rd) {
  if (!P.direct_all) return;
  unsigned mask = (unsigned)seg_mask_io;
  for (int j = 0; j < P.n_seg; j++) {
    if (!((mask >> j) & 1u)) continue;
    size_t e = (size_t)lane * P.n_seg + j;
    if (L.occl[e]) continue;
    const ShadowRec* sr = L.sray + e;
    rd.Lr += sr->pr; rd.Lg += sr->pg; rd.Lb += sr->pb;
  }
  seg_mask_io = 0;
}

// ALL = true: returns in seg_mask the segments (one per light) written to L.sray[lane * n_seg + j]
template <bool ALL>
GP_D void shade_lane_direct(const DevScene& sc, const Lanes& L, const RenderParams& P, long long lane, bool& cont, bool& finished, bool& shadow,
                            unsigned& seg_mask, unsigned& n_counts, int& bad) {
  PathRec pt = L.path[lane];
  RadRec rd = L.rad[lane];
  RayRec rr = L.ray[lane];
  int has_sample = 0;  // the pending-segment mask (UniformSampleAll only)
  if (ALL) { has_sample = L.fsum[lane].has_sample; direct_collect(L, P, lane, has_sample, rd); }
  int packed = pt.bounces;
  const int level = packed & 255;  // frames filed so far == specular bounces taken; Li's depth argument is 2 * level
  double* fr = L.frames + ((size_t)lane * P.direct_levels) * 8;
  finished = true;
  packed |= kDirectHitBit;
  if (level > 0) {  // the previous level's direct light is complete now: file it, start this level's sum at zero
    fr[(level - 1) * 8 + 0] = rd.Lr; fr[(level - 1) * 8 + 1] = rd.Lg; fr[(level - 1) * 8 + 2] = rd.Lb;
    rd.Lr = 0; rd.Lg = 0; rd.Lb = 0;
  }
  Ray ray;
  ray.o = mk3(rr.ox, rr.oy, rr.oz);
  ray.d = mk3(rr.dx, rr.dy, rr.dz);
  ray.tmax = rr.tmax;
  Hit h;
  int prim;
  hit_record(sc, rr.hit_rec, ray, ray.tmax, &h, &prim, bad);
  BSDF bsdf;
  if (!compute_scattering(sc, prim, h, &bsdf, false)) {
    n_counts += 0x10000u;
  } else {
    Smp s;
    s.state = pt.rng_state; s.inc = pt.rng_inc; s.sidx = pt.sidx;
    s.cur1 = (packed >> 8) & 255; s.cur2 = (packed >> 16) & 255; s.lane = lane;
    const unsigned long long fast_pixel = P.mode == 1 ? lane_fast_pixel(P, lane, P.tile_size == 1 ? 0 : L.fsum[lane].pix) : 0ULL;
    Intr ref; ref.p = h.p; ref.perr = h.perr; ref.n = h.n;
    // --- direct light (directlighting.go:85-96): UniformSampleOneLight (integrator.go:48-77) or, ALL, every light once
    //     (UniformSampleAllLights' single-sample branch, integrator.go:31-36: uLight then uScattering per light)
    if (sc.n_lights > 0) {
      int offset = 0;
      bool sample = true;
      if (!ALL) {
        double u = get1d(s, L, P, fast_pixel);
        double lightPdf;
        sample_discrete(sc.light_cdf, sc.n_lights, sc.light_func_int, u, &offset, &lightPdf);
        sample = lightPdf != 0.0;
      }
      const int n_loop = ALL ? sc.n_lights : 1;
      for (int j = 0; j < n_loop && sample; j++) {
        if (ALL) offset = j;
        double ulx, uly, usx, usy;
        get2d(s, P, &ulx, &uly);
        get2d(s, P, &usx, &usy);
        const int flags = BSDF_ALL & ~BSDF_SPECULAR;  // EstimateDirect (integrator.go:79-195), specular = false
        LightSample ls;
        light_sample_li(sc, sc.lights[offset], ref, ulx, uly, &ls);
        if (!ls.delta) n_counts += 1u;
        if (ls.pdf > 0 && !is_black(ls.Li)) {
          RGB f = bsdf_f(bsdf, h.wo, ls.wi, flags);
          f = f * fabs(dot(ls.wi, h.ns));
          double scatteringPdf = bsdf_pdf(bsdf, h.wo, ls.wi, flags);
          if (!is_black(f)) {
            RGB Ld;
            if (ls.delta) Ld = (f * ls.Li) / ls.pdf;
            else {
              double ff = 1.0 * ls.pdf, gg = 1.0 * scatteringPdf;  // PowerHeuristic (sampling.go:208-212)
              double weight = (ff * ff) / (ff * ff + gg * gg);
              Ld = ((f * ls.Li) * weight) / ls.pdf;
            }
            Ld = rgb(0, 0, 0) + Ld;
            V3 d = spawn_ray_to(ref, ls.p1);
            ShadowRec sr;
            sr.ox = ref.p.x; sr.oy = ref.p.y; sr.oz = ref.p.z; sr.dx = d.x; sr.dy = d.y; sr.dz = d.z;
            sr.pr = Ld.r; sr.pg = Ld.g; sr.pb = Ld.b;   // L.AddAssign(Ld) on this level's sum
            sr.gt10 = (!ALL && max_comp(Ld) > 10) ? 1 : 0;  // integrator.go:73-75 (UniformSampleOneLight only) panics; counted when unoccluded
            sr.pad = 1;                                 // an occluded segment contributes nothing at all (Li set to 0, :117-121)
            sr.pad2[0] = 0; sr.pad2[1] = 0;
            if (ALL) { L.sray[(size_t)lane * P.n_seg + j] = sr; seg_mask |= 1u << j; }
            else { L.sray[lane] = sr; shadow = true; }
          }
        }
      }
      if (ALL) has_sample = (int)seg_mask;
    }
    // --- specular recursion (directlighting.go:98-102): both terms draw their Get2D before anything else
    if (2 * level + 1 < P.max_depth) {
      double u1x, u1y, u2x, u2y;
      get2d(s, P, &u1x, &u1y);  // SpecularReflect: BSDF.SampleF(wo, u, Reflection|Specular) matches no BxDF -> Spectrum(0)
      get2d(s, P, &u2x, &u2y);  // SpecularTransmit: BSDF.SampleF(wo, u, Transmission|Specular)
      if (bsdf.kind == BX_GLASS_SPLIT && !is_black(bsdf.t)) {
        V3 wo = to_local(bsdf, h.wo);  // si.Wo (integrator.go:391), NOT the un-negated ray direction Path uses
        if (wo.z != 0.0) {
          // SpecularTransmission.SampleF (reflection.go:428-451), mode == Radiance; one matching component, so the
          // remapped sample is unused and pdf stays 1
          bool entering = wo.z > 0;
          double etaI = entering ? 1.0 : bsdf.etaB, etaT = entering ? bsdf.etaB : 1.0;
          V3 n = faceforward(mk3(0, 0, 1), wo);
          double eta = etaI / etaT;
          double cosThetaI = dot(n, wo);
          double sin2ThetaI = go_max(0, 1 - cosThetaI * cosThetaI);
          double sin2ThetaT = eta * eta * sin2ThetaI;
          if (!(sin2ThetaT >= 1)) {
            double cosThetaT = sqrt(1 - sin2ThetaT);
            V3 wi = wo * -eta + n * (eta * cosThetaI - cosThetaT);
            double F = fr_dielectric(wi.z, 1.0, bsdf.etaB);
            RGB ft = bsdf.t * rgb(1.0 - F, 1.0 - F, 1.0 - F);
            ft = ft * ((etaI * etaI) / (etaT * etaT));
            RGB f = ft / fabs(wi.z);
            const double pdf = 1;
            double ad = fabs(dot(wi, h.ns));  // wi is BSDF-local, used as world
            if (pdf > 0 && !is_black(f) && ad != 0.0) {
              V3 o = offset_ray_origin(h.p, h.perr, h.n, wi);  // si.SpawnRay(wi)
              RayRec nr;
              nr.ox = o.x; nr.oy = o.y; nr.oz = o.z; nr.dx = wi.x; nr.dy = wi.y; nr.dz = wi.z;
              nr.tmax = d_inf(); nr.hit_rec = -1; nr.pad = 0;
              L.ray[lane] = nr;
              fr[level * 8 + 3] = f.r; fr[level * 8 + 4] = f.g; fr[level * 8 + 5] = f.b; fr[level * 8 + 6] = ad / pdf;
              cont = true;
              finished = false;
              packed = (packed & ~(255 | kDirectHitBit)) | (level + 1);
            }
          }
        }
      }
    }
    pt.rng_state = s.state;
    packed = (packed & ~((255 << 8) | (255 << 16))) | (s.cur1 << 8) | (s.cur2 << 16);
  }
  // only what a DirectLighting level changes goes back (see shade_lane): the level's radiance sum, the pending-segment mask,
  // the sampler state, the level / dimension counters
  PathRec* const pp = L.path + lane;
  RadRec* const rp = L.rad + lane;
  rp->Lr = rd.Lr; rp->Lg = rd.Lg; rp->Lb = rd.Lb;
  if (ALL) L.fsum[lane].has_sample = has_sample;
  pp->rng_state = pt.rng_state;
  pp->bounces = packed;
}

// the radiance of a finished DirectLighting sample: the chain of frames unwound (see shade_lane_direct)
GP_D RGB direct_unwind(const Lanes& L, const RenderParams& P, long long lane, const PathRec& pt_in, const RadRec& rd_in, int seg_mask) {
  PathRec pt = pt_in;
  RadRec rd = rd_in;
  direct_collect(L, P, lane, seg_mask, rd);
  const int level = pt.bounces & 255;
  const bool at_hit = (pt.bounces & kDirectHitBit) != 0;
  const double* fr = L.frames + ((size_t)lane * P.direct_levels) * 8;
  // at a hit: pt.L is that level's own direct light.  Otherwise the last ray escaped: its Li is the sum of the lights'
  // Le = 0, and pt.L is still the previous level's direct light.
  RGB Ln = at_hit ? rgb(rd.Lr, rd.Lg, rd.Lb) : rgb(0, 0, 0);
  for (int j = level - 1; j >= 0; j--) {
    RGB D = (j == level - 1 && !at_hit) ? rgb(rd.Lr, rd.Lg, rd.Lb) : rgb(fr[j * 8], fr[j * 8 + 1], fr[j * 8 + 2]);
    RGB f = rgb(fr[j * 8 + 3], fr[j * 8 + 4], fr[j * 8 + 5]);
    Ln = D + (f * Ln) * fr[j * 8 + 6];  // L.AddAssign(f.Mul(s.Li(...)).MulScalar(wi.AbsDot(ns) / pdf))
  }
  return Ln;
}

// One Path.Li loop body per lane (path.go:40-155) after the closest-hit query: scattering functions, one light
// sample (UniformSampleOneLight / EstimateDirect) whose visibility test is deferred to the shadow queue, BSDF
// sampling, throughput update, SpawnRay, Russian roulette.
// INTEG: 0 = Path, 1 = DirectLighting / UniformSampleOne, 2 = DirectLighting / UniformSampleAll.  MODE: sampler mode.
// CLS (Path only): the ONE shade class this launch works on — 0 triangle + Lambert, 1 triangle + other material,
// 2 sphere / disk + Lambert, 3 sphere / disk + other — or -1: all four bins back to back.  One instantiation per class
// keeps the sphere / disk hit reconstruction (Go trig, two matrix pairs) and the Oren-Nayar / mirror / glass lobes out
// of the kernel that shades the common case, whose register budget then allows more resident warps for what is one
// long dependent float64 chain per lane.
#ifndef GP_SHADE_BLOCKS0
#define GP_SHADE_BLOCKS0 4
#endif
#ifndef GP_SHADE_BLOCKS_OTHER
#define GP_SHADE_BLOCKS_OTHER 4
#endif
template <int CLS> struct ShadeBlocks { static constexpr int value = GP_SHADE_BLOCKS_OTHER; };
template <> struct ShadeBlocks<0> { static constexpr int value = GP_SHADE_BLOCKS0; };
template <int INTEG, int MODE, int CLS>
__global__ void __launch_bounds__(128, ShadeBlocks<CLS>::value) k_shade(DevScene sc, Lanes L, RenderParams P_in, Queues Q, RenderCounters* ctr) {
  RenderParams P = P_in;
  P.mode = MODE;
  P.integrator = INTEG == 0 ? 0 : 1;
  // lanes whose ray hit something, binned by shade class; CLS < 0: the kernel walks the four bins back to back, each
  // starting on a warp boundary, so that (almost) every warp shades one kind of hit
  long long n0 = Q.cnt[8], n1 = Q.cnt[9], n2 = Q.cnt[10], n3 = Q.cnt[11];
  long long o1 = (n0 + 31) & ~31LL, o2 = o1 + ((n1 + 31) & ~31LL), o3 = o2 + ((n2 + 31) & ~31LL);
  long long n = o3 + n3;
  if (CLS >= 0) n = Q.cnt[8 + (CLS >= 0 ? CLS : 0)];
  int lane_id = threadIdx.x & 31;
  unsigned n_counts = 0;  // this thread's dead MIS rays (low 16 bits) and unsupported materials (high 16): a thread sees < 2^15 lanes per launch
  int bad = 0;
  // (per-lane pushes: batching them as raygen does bought nothing here on config 2 and cost 4-15 % on the sphere shade classes)
  __shared__ int s_cnt[4][3];
  __shared__ int s_base[3];
  for (long long cbase = (long long)blockIdx.x * blockDim.x; cbase < n; cbase += (long long)gridDim.x * blockDim.x) {
    long long i = cbase + threadIdx.x;
    const int* bin_q;
    long long bi;
    if (CLS >= 0) { bin_q = Q.shade[CLS >= 0 ? CLS : 0]; bi = i < n ? i : -1; }
    else if (i >= o3) { bin_q = Q.shade[3]; bi = i - o3; if (bi >= n3) bi = -1; }
    else if (i >= o2) { bin_q = Q.shade[2]; bi = i - o2; if (bi >= n2) bi = -1; }
    else if (i >= o1) { bin_q = Q.shade[1]; bi = i - o1; if (bi >= n1) bi = -1; }
    else { bin_q = Q.shade[0]; bi = i; if (bi >= n0) bi = -1; }
    bool valid = i < n && bi >= 0;
    bool cont = false, finished = false, shadow = false;
    unsigned seg_mask = 0;
    long long lane = 0;
    if (valid) {
      lane = bin_q[bi];
      if (INTEG == 0) shade_lane(sc, L, P, lane, cont, finished, shadow, n_counts, bad, CLS);
      else if (INTEG == 1) shade_lane_direct<false>(sc, L, P, lane, cont, finished, shadow, seg_mask, n_counts, bad);
      else shade_lane_direct<true>(sc, L, P, lane, cont, finished, shadow, seg_mask, n_counts, bad);
    }
    if (INTEG == 2) {  // UniformSampleAll: one shadow-queue entry per emitted segment (entry = lane * n_seg + light)
      for (int j = 0; j < P.n_seg; j++) queue_push(Q.shadow, Q.cnt + 2, (seg_mask >> j) & 1u, (int)(lane * P.n_seg + j));
    }
    int* const qs[3] = {Q.shadow, Q.extend_next, Q.regen_next};
    int* const cs[3] = {Q.cnt + 2, Q.cnt + 1, Q.cnt + 4};
    const bool ps[3] = {shadow, cont, finished && valid};
    block_push<3>(qs, cs, ps, (int)lane, s_cnt, s_base);
  }
  unsigned long long n_unsupported = warp_sum((unsigned long long)(n_counts >> 16)), n_dead = warp_sum((unsigned long long)(n_counts & 0xffffu));
  if (lane_id == 0) {
    if (n_unsupported) atomicAdd(&ctr->unsupported, n_unsupported);
    if (n_dead) atomicAdd(&ctr->dead_mis_rays, n_dead);
  }
  if (bad) atomicAdd(&ctr->efloat_panics, 1ULL);
}

// end of a wavefront iteration: rotate the queues on the device and publish the number of lanes still in flight
__global__ void k_advance(Queues Q, RenderCounters* ctr, int* host_visible_remaining) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    ctr->closest_rays += (unsigned long long)Q.cnt[0];
    ctr->shadow_rays += (unsigned long long)Q.cnt[2];
    ctr->shaded += (unsigned long long)Q.cnt[8] + (unsigned long long)Q.cnt[9] + (unsigned long long)Q.cnt[10] + (unsigned long long)Q.cnt[11];
    Q.cnt[0] = Q.cnt[1];  // extend <- extend_next (the host swaps the pointers)
    Q.cnt[1] = 0;
    Q.cnt[2] = 0;
    Q.cnt[3] = Q.cnt[4];  // regen <- regen_next
    Q.cnt[4] = 0;
    Q.cnt[8] = 0; Q.cnt[9] = 0; Q.cnt[10] = 0; Q.cnt[11] = 0;
    Q.cnt[6] = 0;  // work counters of the persistent traversal warps
    Q.cnt[7] = 0;
    *host_visible_remaining = Q.cnt[0] + Q.cnt[3];
  }
}

// FAST mode with the uniform footprint: a tile's lane groups each hold a partial RGB sum of the tile's one pixel (PathRec.pad).
// They are added up in ascending group order into group 0's record before the merge — a tile is ONE FilmTile whose samples
// were accumulated by several partial sums, and MergeFilmTile converts it to XYZ once (film.go:115-132) — so that the merge
// reads one record per tile instead of `groups` (32 groups: 5.8 -> about 1 ms per 1080p frame).
// LAST (RenderParams.last_in_place): a lane's last sample was never retired by raygen; its radiance is still in PathRec.L and
// is added here: group sum = pad + L', L' = L after renderWorker's NaN replacement (integrator.go:256-257) — the addition
// film_add_uniform would have made, in the same place of the same order.  SINGLE: no lane has more than one sample, so
// every pad is still +0 and only the L sector of each record is read.
GP_D void last_sample_fix(double& r, double& g, double& b, unsigned long long& nans) {
  if (is_nan(r) || is_nan(g) || is_nan(b)) { r = g = b = 0.1; nans++; }
  r = r * (1.0 * 1.0); g = g * (1.0 * 1.0); b = b * (1.0 * 1.0);  // L.MulScalar(sampleWeight * filterWeight)
}
GP_D void last_sample_radiance(const RadRec* q, double& r, double& g, double& b, unsigned long long& nans) {
  r = __ldcs(&q->Lr); g = __ldcs(&q->Lg); b = __ldcs(&q->Lb);
  last_sample_fix(r, g, b, nans);
}
// One WARP per tile: the tile's `groups` records are contiguous (groups x 128 bytes), so the lanes fetch them side by side — 32
// sectors of one or two DRAM pages per load instruction instead of 32 sectors of 32 different tiles 8 KB apart —, two rounds
// of loads in flight, and leave each record's value (the NaN replacement applied) in shared memory; lanes 0-2 then add up one
// colour channel each, in ascending group order: the same chain of additions one thread made before.
// Dynamic shared memory: 3 * groups doubles per warp.
template <bool LAST, bool SINGLE>
__global__ void __launch_bounds__(128) k_group_sums(Lanes L, RenderParams P, RenderCounters* ctr) {
  extern __shared__ double s_gs[];
  const int G = P.groups;
  const long long n_tiles = P.lanes_active / G;
  const int lane_id = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* const v = s_gs + (size_t)w * 3 * G;  // [channel][group]
  const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
  unsigned long long nans = 0;
  for (long long t = (long long)blockIdx.x * (blockDim.x >> 5) + w; t < n_tiles; t += n_warps) {
    FilmRec* base = L.fsum + t * G;
    const RadRec* rbase = L.rad + t * G;
    for (int k0 = 0; k0 < G; k0 += 64) {
      const int ka = k0 + lane_id, kb = ka + 32;
      double2 la0 = make_double2(0, 0), la1 = la0, lb0 = la0, lb1 = la0, pa0 = la0, pa1 = la0, pb0 = la0, pb1 = la0;
      if (ka < G) {
        if (LAST) { la0 = __ldcs((const double2*)&rbase[ka].Lr); la1 = __ldcs((const double2*)&rbase[ka].Lb); }      // {Lr, Lg} {Lb, eta}
        if (!SINGLE) { pa0 = make_double2(__ldcs(&base[ka].pad[0]), __ldcs(&base[ka].pad[1])); pa1.x = __ldcs(&base[ka].pad[2]); }
      }
      if (kb < G) {
        if (LAST) { lb0 = __ldcs((const double2*)&rbase[kb].Lr); lb1 = __ldcs((const double2*)&rbase[kb].Lb); }
        if (!SINGLE) { pb0 = make_double2(__ldcs(&base[kb].pad[0]), __ldcs(&base[kb].pad[1])); pb1.x = __ldcs(&base[kb].pad[2]); }
      }
      if (ka < G) {
        double r = pa0.x, g = pa0.y, b = pa1.x;
        if (LAST) { double lr = la0.x, lg = la0.y, lb = la1.x; last_sample_fix(lr, lg, lb, nans); r += lr; g += lg; b += lb; }
        v[ka] = r; v[G + ka] = g; v[2 * G + ka] = b;
      }
      if (kb < G) {
        double r = pb0.x, g = pb0.y, b = pb1.x;
        if (LAST) { double lr = lb0.x, lg = lb0.y, lb = lb1.x; last_sample_fix(lr, lg, lb, nans); r += lr; g += lg; b += lb; }
        v[kb] = r; v[G + kb] = g; v[2 * G + kb] = b;
      }
    }
    __syncwarp();
    if (lane_id < 3) {
      const double* c = v + lane_id * G;
      double acc = c[0];
      int k = 1;
      for (; k + 4 <= G; k += 4) { const double x0 = c[k], x1 = c[k + 1], x2 = c[k + 2], x3 = c[k + 3]; acc += x0; acc += x1; acc += x2; acc += x3; }
      for (; k < G; k++) acc += c[k];
      base->pad[lane_id] = acc;
    }
    __syncwarp();
  }
  if (LAST) {
    nans = warp_sum(nans);
    if ((threadIdx.x & 31) == 0 && nans) atomicAdd(&ctr->nan_samples, nans);
  }
}
// The same fold for FEW lane groups (a rank's share of a multi-GPU frame: 8 groups at N = 8): one THREAD per tile with all of the
// tile's records in flight at once — a warp per tile would run at a quarter of its lanes and one dependent chain per tile.
template <bool LAST, bool SINGLE, int GMAX>
__global__ void __launch_bounds__(128) k_group_sums_small(Lanes L, RenderParams P, RenderCounters* ctr) {
  const int G = P.groups;
  const long long n_tiles = P.lanes_active / G;
  unsigned long long nans = 0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += (long long)gridDim.x * blockDim.x) {
    FilmRec* base = L.fsum + t * G;
    const RadRec* rbase = L.rad + t * G;
    double2 l0[GMAX], l1[GMAX];
    double p0[GMAX], p1[GMAX], p2[GMAX];
#pragma unroll
    for (int k = 0; k < GMAX; k++) {
      if (k < G) {
        if (LAST) { l0[k] = __ldcs((const double2*)&rbase[k].Lr); l1[k] = __ldcs((const double2*)&rbase[k].Lb); }
        if (!SINGLE) { p0[k] = __ldcs(&base[k].pad[0]); p1[k] = __ldcs(&base[k].pad[1]); p2[k] = __ldcs(&base[k].pad[2]); }
      }
    }
    double r = 0, g = 0, b = 0;
#pragma unroll
    for (int k = 0; k < GMAX; k++) {
      if (k < G) {
        double xr = 0, xg = 0, xb = 0;
        if (!SINGLE) { xr = p0[k]; xg = p1[k]; xb = p2[k]; }
        if (LAST) { double lr = l0[k].x, lg = l0[k].y, lb = l1[k].x; last_sample_fix(lr, lg, lb, nans); xr += lr; xg += lg; xb += lb; }
        if (k == 0) { r = xr; g = xg; b = xb; } else { r += xr; g += xg; b += xb; }
      }
    }
    base->pad[0] = r; base->pad[1] = g; base->pad[2] = b;
  }
  if (LAST) {
    nans = warp_sum(nans);
    if ((threadIdx.x & 31) == 0 && nans) atomicAdd(&ctr->nan_samples, nans);
  }
}
// last_in_place when the lane groups of a pass cannot be folded per tile (a pass that does not hold whole tiles): every lane
// retires its last sample into its own pad
__global__ void k_fold_last(Lanes L, RenderParams P, RenderCounters* ctr) {
  unsigned long long nans = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.lanes_active; i += (long long)gridDim.x * blockDim.x) {
    FilmRec* q = L.fsum + i;
    double lr, lg, lb;
    last_sample_radiance(L.rad + i, lr, lg, lb, nans);
    q->pad[0] += lr; q->pad[1] += lg; q->pad[2] += lb;
  }
  nans = warp_sum(nans);
  if ((threadIdx.x & 31) == 0 && nans) atomicAdd(&ctr->nan_samples, nans);
}

// Film.MergeFilmTile (film.go:115-132): every film pixel gathers the tiles that cover it in ascending tile order
// (the order a single renderWorker would merge them), converting each tile's RGB sum to XYZ first (spectrum.go:35-41).
__global__ void k_film_merge(Lanes L, RenderParams P, double* __restrict__ film) {
  long long fw = P.cx1 - P.cx0, fh = P.cy1 - P.cy0;
  long long npx = fw * fh;
  long long ext = (long long)ceil(P.frx > P.fry ? P.frx : P.fry) + 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
    long long frow = div_nn(i, fw);
    long long x = P.cx0 + (i - frow * fw), y = P.cy0 + frow;
    double X = film[i * 4], Y = film[i * 4 + 1], Z = film[i * 4 + 2], W = film[i * 4 + 3];
    long long ty0 = (y - ext - P.cy0) / P.tile_size, ty1 = (y + ext - P.cy0) / P.tile_size;
    long long tx0 = (x - ext - P.cx0) / P.tile_size, tx1 = (x + ext - P.cx0) / P.tile_size;
    if (P.uniform_fp) {
      // tileSize 1, every sample at its pixel's integer corner: tile (= pixel) t reaches pixel x only if
      // ceil(t - 0.5 - r) <= x <= floor(t - 0.5 + r) (uniform_footprint), i.e. x + 0.5 - r <= t <= x + 0.5 + r — with the box
      // filter's r = 0.5 that is 2 x 2 tiles instead of the 5 x 5 of the general search window (the membership tests below are
      // unchanged, and so is the order of the additions)
      tx0 = (long long)ceil((double)x + 0.5 - P.frx) - P.cx0; tx1 = (long long)floor((double)x + 0.5 + P.frx) - P.cx0;
      ty0 = (long long)ceil((double)y + 0.5 - P.fry) - P.cy0; ty1 = (long long)floor((double)y + 0.5 + P.fry) - P.cy0;
    }
    if (ty0 < 0) ty0 = 0; if (tx0 < 0) tx0 = 0;
    if (ty1 >= P.nty) ty1 = P.nty - 1; if (tx1 >= P.ntx) tx1 = P.ntx - 1;
    for (long long ty = ty0; ty <= ty1; ty++)
      for (long long tx = tx0; tx <= tx1; tx++) {
        long long tile = ty * P.ntx + tx;
        if (tile % P.world != P.rank) continue;
        long long x0, y0, x1, y1, bx0, by0, bx1, by1;
        tile_bounds(P, tile, &x0, &y0, &x1, &y1);
        tile_pixel_bounds(P, x0, y0, x1, y1, &bx0, &by0, &bx1, &by1);
        if (x < bx0 || x >= bx1 || y < by0 || y >= by1) continue;
        if (x - bx0 >= P.tpw || y - by0 >= P.tph) continue;  // the bound's last column / row: never touched by a sample, not stored
        size_t k = (size_t)((y - by0) * P.tpw + (x - bx0)) * 4;
        if (P.uniform_fp) {  // tileSize 1: the tile is the pixel (x0, y0); is (x, y) inside its samples' footprint?
          long long p0x, p0y, p1x, p1y;
          uniform_footprint(P, x0, y0, bx0, by0, bx1, by1, &p0x, &p0y, &p1x, &p1y);
          if (x < p0x || x >= p1x || y < p0y || y >= p1y) continue;
        }
        // the tile's lane groups in ascending order (one group in STRICT mode; one pre-reduced sum after k_group_sums)
        const int n_grp = P.groups_merged ? 1 : P.groups;
        for (int grp = 0; grp < n_grp; grp++) {
          long long lane = (tile / P.world) * P.groups + grp - P.lane_base;
          if (lane < 0 || lane >= P.lanes_active) continue;
          double r, g, b, w;
          if (P.uniform_fp) {
            const FilmRec* pt = L.fsum + lane;
            r = pt->pad[0]; g = pt->pad[1]; b = pt->pad[2];
            // filterWeightSum = the samples this lane retired: indices 1 .. spp-1 (sampler.go:29-34) with
            // s % (s_world * groups) == s_rank * groups + grp (pre-reduced groups: s % s_world == s_rank)
            const int s_mod = P.groups_merged ? P.s_world : P.s_world * P.groups, s_res = P.groups_merged ? P.s_rank : P.s_rank * P.groups + grp;
            int cnt = 0;
            if (P.spp > 1) {
              int first = s_res == 0 ? s_mod : s_res;  // the first index >= 1 in the residue class
              if (first <= P.spp - 1) cnt = (P.spp - 1 - first) / s_mod + 1;
            }
            w = (double)cnt;
          } else {
            const double* q = L.tilepix + (size_t)lane * L.tile_stride + k;
            r = q[0]; g = q[1]; b = q[2]; w = q[3];
          }
          double tx_, ty_, tz_;
          rgb_to_xyz(r, g, b, &tx_, &ty_, &tz_);
          X += tx_; Y += ty_; Z += tz_;
          W += w;
        }
      }
    film[i * 4] = X; film[i * 4 + 1] = Y; film[i * 4 + 2] = Z; film[i * 4 + 3] = W;
  }
}

// pass start: the lane's initial sampler / path state (Sampler.Clone(seed = tile index), pixel.go:34-42); the first raygen
// launch of a pass builds it in registers instead of reading it back from memory
GP_D PathRec initial_path(const RenderParams& P, long long lane) {
  long long tile = (P.groups == 1 ? P.lane_base + lane : div_nn(P.lane_base + lane, P.groups)) * P.world + P.rank;
  Smp s;
  s.state = 0x853c49e6748fea9bULL; s.inc = 0xda3e39cb94b95bdbULL;
  rng_set_sequence(s, (unsigned long long)tile);
  PathRec pt;
  pt.br = pt.bg = pt.bb = 1.0; pt.fx = pt.fy = 0;
  pt.rng_state = s.state; pt.rng_inc = s.inc;
  pt.sidx = 0; pt.bounces = 0;
  return pt;
}

}  // namespace gp
