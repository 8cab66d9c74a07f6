// ORACLE — test infrastructure only.  Nothing under go-pbrt_b200/ may include, link or call this.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load liboracle.so.
//
// C entry points (ctypes) over oracle_core.h / oracle_render.h.  Built by oracle/Makefile:
//   g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -pthread
#include <chrono>
#include <cstring>
#include <string>

#include "oracle_render.h"

using namespace oracle;

extern "C" {

// accel_mode: 0 = reference RecursiveBuild(SplitSAH)+[64]-stack traversal, 1 = oracle's own median tree, 2 = brute force
void* oracle_scene_create(const gopbrt_scene_desc* d, int accel_mode) { return scene_from_desc(d, accel_mode); }
void oracle_scene_destroy(void* s) { delete (Scene*)s; }
void oracle_scene_world_bound(void* s, double* out6) {
  Scene* sc = (Scene*)s;
  out6[0] = sc->world.mn.x; out6[1] = sc->world.mn.y; out6[2] = sc->world.mn.z;
  out6[3] = sc->world.mx.x; out6[4] = sc->world.mx.y; out6[5] = sc->world.mx.z;
}
int64_t oracle_scene_bvh_nodes(void* s) { return (int64_t)((Scene*)s)->nodes.size(); }
void oracle_prim_bound(void* s, int64_t i, double* out6) {
  const B3& b = ((Scene*)s)->prim_bounds[i];
  out6[0] = b.mn.x; out6[1] = b.mn.y; out6[2] = b.mn.z; out6[3] = b.mx.x; out6[4] = b.mx.y; out6[5] = b.mx.z;
}
uint64_t oracle_efloat_panics() { return g_counters.efloat_panics.load(); }
uint64_t oracle_stack_overflows() { return g_counters.stack_overflows.load(); }
void oracle_reset_counters() { g_counters.efloat_panics = 0; g_counters.stack_overflows = 0; }

// Aggregate.Intersect over n rays.  counts (may be NULL): [nodes, prims] totals.
void oracle_trace_closest(void* s, int64_t n, const double* ox, const double* oy, const double* oz, const double* dx,
                          const double* dy, const double* dz, const double* tmax, int32_t* prim, double* t, double* p,
                          double* nrm, uint64_t* counts, int threads) {
  Scene* sc = (Scene*)s;
  threads = std::max(1, threads);
  std::vector<TravStats> tsv(threads);
  auto work = [&](int tid) {
    for (int64_t i = tid; i < n; i += threads) {
      Ray r{V3{ox[i], oy[i], oz[i]}, V3{dx[i], dy[i], dz[i]}, tmax[i], 0};
      Hit h;
      bool hit = scene_intersect(*sc, r, &h, &tsv[tid]);
      prim[i] = hit ? h.prim : -1;
      t[i] = hit ? r.tmax : tmax[i];
      if (p) { p[3 * i] = hit ? h.p.x : 0; p[3 * i + 1] = hit ? h.p.y : 0; p[3 * i + 2] = hit ? h.p.z : 0; }
      if (nrm) { nrm[3 * i] = hit ? h.n.x : 0; nrm[3 * i + 1] = hit ? h.n.y : 0; nrm[3 * i + 2] = hit ? h.n.z : 0; }
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < threads; i++) th.emplace_back(work, i);
  work(0);
  for (auto& x : th) x.join();
  if (counts) { counts[0] = counts[1] = 0; for (auto& ts : tsv) { counts[0] += ts.nodes; counts[1] += ts.prims; } }
}

void oracle_trace_any(void* s, int64_t n, const double* ox, const double* oy, const double* oz, const double* dx,
                      const double* dy, const double* dz, const double* tmax, uint8_t* hit, uint64_t* counts, int threads) {
  Scene* sc = (Scene*)s;
  threads = std::max(1, threads);
  std::vector<TravStats> tsv(threads);
  auto work = [&](int tid) {
    for (int64_t i = tid; i < n; i += threads) {
      Ray r{V3{ox[i], oy[i], oz[i]}, V3{dx[i], dy[i], dz[i]}, tmax[i], 0};
      hit[i] = scene_intersect_p(*sc, r, &tsv[tid]) ? 1 : 0;
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < threads; i++) th.emplace_back(work, i);
  work(0);
  for (auto& x : th) x.join();
  if (counts) { counts[0] = counts[1] = 0; for (auto& ts : tsv) { counts[0] += ts.nodes; counts[1] += ts.prims; } }
}

// full shading-relevant hit record for one ray: out[0..] = p(3) perr(3) n(3) wo(3) ns(3) sdpdu(3) u v t prim
int oracle_hit_record(void* s, const double* o, const double* d, double tmax, double* out) {
  Scene* sc = (Scene*)s;
  Ray r{V3{o[0], o[1], o[2]}, V3{d[0], d[1], d[2]}, tmax, 0};
  Hit h;
  if (!scene_intersect(*sc, r, &h)) return 0;
  const V3* vs[6] = {&h.p, &h.perr, &h.n, &h.wo, &h.ns, &h.sdpdu};
  for (int k = 0; k < 6; k++) { out[3 * k] = vs[k]->x; out[3 * k + 1] = vs[k]->y; out[3 * k + 2] = vs[k]->z; }
  out[18] = h.u; out[19] = h.v; out[20] = r.tmax; out[21] = (double)h.prim;
  return 1;
}

struct oracle_render_opts {
  int32_t rank, world, threads, deterministic;
  int64_t tile_begin, tile_end;
};
// stats_out[0..10] = camera, closest, shadow, dead_mis, nodes, prims, snodes, sprims, gt10, nan, unsupported; returns seconds
double oracle_render(void* s, const gopbrt_camera* cam, const gopbrt_sampler* smp, const gopbrt_integrator* ig,
                     const gopbrt_film* film, const oracle_render_opts* o, double* film_out, uint64_t* stats_out) {
  Scene* sc = (Scene*)s;
  RenderOpts ro;
  ro.rank = o->rank; ro.world = o->world; ro.threads = o->threads; ro.deterministic = o->deterministic;
  ro.tile_begin = o->tile_begin; ro.tile_end = o->tile_end;
  RenderStats st;
  auto t0 = std::chrono::steady_clock::now();
  render(*sc, *cam, *smp, *ig, *film, ro, film_out, &st);
  auto t1 = std::chrono::steady_clock::now();
  if (stats_out) {
    uint64_t v[11] = {st.camera_rays, st.closest_rays, st.shadow_rays, st.dead_mis_rays, st.nodes, st.prims, st.snodes, st.sprims,
                      st.radiance_gt10, st.nan_samples, st.unsupported_material};
    std::memcpy(stats_out, v, sizeof(v));
  }
  return std::chrono::duration<double>(t1 - t0).count();
}

// one camera ray (GenerateRayDifferential): out = o(3) d(3)
void oracle_camera_ray(const gopbrt_camera* cam, double fx, double fy, double lx, double ly, double time, double* out) {
  Ray r = camera_ray(*cam, P2{fx, fy}, P2{lx, ly}, time);
  out[0] = r.o.x; out[1] = r.o.y; out[2] = r.o.z; out[3] = r.d.x; out[4] = r.d.y; out[5] = r.d.z;
}

// ---- known-answer helpers for tests/test_oracle_golden.py ----
void oracle_kat_efloat_add(double a, double aerr, double b, double berr, double* out3) {
  EF r = ef_add(ef_new(a, aerr), ef_new(b, berr));
  out3[0] = r.v; out3[1] = r.lo; out3[2] = r.hi;
}
void oracle_kat_offset_ray_origin(const double* p, const double* perr, const double* n, const double* w, double* out3) {
  V3 r = offset_ray_origin(V3{p[0], p[1], p[2]}, V3{perr[0], perr[1], perr[2]}, V3{n[0], n[1], n[2]}, V3{w[0], w[1], w[2]});
  out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
double oracle_kat_machine_epsilon() { return gm::MachineEpsilon(); }
double oracle_kat_gamma(double n) { return gm::Gamma(n); }
void oracle_kat_transform_ray(const gopbrt_transform* t, const double* o, const double* d, double* out6) {
  Ray r = xf_ray(xf_from(*t), Ray{V3{o[0], o[1], o[2]}, V3{d[0], d[1], d[2]}, gm::Inf, 0}, nullptr, nullptr);
  out6[0] = r.o.x; out6[1] = r.o.y; out6[2] = r.o.z; out6[3] = r.d.x; out6[4] = r.d.y; out6[5] = r.d.z;
}
void oracle_kat_transform_point(const gopbrt_transform* t, const double* p, const double* pe, double* out6) {
  V3 e;
  V3 r = xf_point(xf_from(*t), V3{p[0], p[1], p[2]}, V3{pe[0], pe[1], pe[2]}, &e);
  out6[0] = r.x; out6[1] = r.y; out6[2] = r.z; out6[3] = e.x; out6[4] = e.y; out6[5] = e.z;
}
void oracle_kat_spawn_ray_to(const double* p0, const double* p1, double* out7) {
  Ray r = spawn_ray_to(Intr{V3{p0[0], p0[1], p0[2]}, V3{}, V3{}}, Intr{V3{p1[0], p1[1], p1[2]}, V3{}, V3{}}, 0);
  out7[0] = r.o.x; out7[1] = r.o.y; out7[2] = r.o.z; out7[3] = r.d.x; out7[4] = r.d.y; out7[5] = r.d.z; out7[6] = r.tmax;
}
void oracle_kat_rng(uint64_t seed, int use_seed, int n, uint32_t* out) {
  Rng r;
  if (use_seed) r.set_sequence(seed);
  for (int i = 0; i < n; i++) out[i] = r.u32();
}
double oracle_kat_trig(int which, double x, double y) {
  switch (which) {
    case 0: return gm::Sin(x);
    case 1: return gm::Cos(x);
    case 2: return gm::Tan(x);
    case 3: return gm::Atan2(x, y);
    case 4: return gm::Acos(x);
    case 5: return gm::Asin(x);
  }
  return 0;
}
// sampler stream: the first `n` Get1D values and Get2D values of sample `idx` of a freshly StartPixel'ed pixel
void oracle_kat_sampler(const gopbrt_sampler* cfg, uint64_t seed, int n_pixels_skip, int sample_idx, int n, double* out1d, double* out2d) {
  Sampler s;
  s.init(*cfg);
  s.clone_seed(seed);
  for (int k = 0; k <= n_pixels_skip; k++) s.start_pixel();
  for (int k = 0; k < sample_idx; k++) s.start_next_sample();
  for (int i = 0; i < n; i++) out1d[i] = s.get1d();
  for (int i = 0; i < n; i++) { P2 p = s.get2d(); out2d[2 * i] = p.x; out2d[2 * i + 1] = p.y; }
}

// ---- generic known-answer hook: evaluates ONE oracle function on flat float64 arguments (tests/test_shading_kats.py
// checks it against tests/golden/shading_kats.json, the independent plain-Python restatement of the cited Go lines).
// Returns the number of doubles written, or -1 for an unknown function / bad argument count.
int oracle_kat_eval(void* scene, const char* fn_c, const double* in, int n_in, double* out, int n_out) {
  std::string fn(fn_c);
  auto v3 = [&](int i) { return V3{in[i], in[i + 1], in[i + 2]}; };
  auto put3 = [&](int i, V3 v) { out[i] = v.x; out[i + 1] = v.y; out[i + 2] = v.z; };
  auto putc = [&](int i, RGB c) { out[i] = c.c[0]; out[i + 1] = c.c[1]; out[i + 2] = c.c[2]; };
  if (fn == "fr_dielectric" && n_in == 3 && n_out >= 1) { out[0] = fr_dielectric(in[0], in[1], in[2]); return 1; }
  if (fn == "oren_nayar_f" && n_in == 10 && n_out >= 3) {
    BxDF x = make_oren_nayar(RGB(in[1], in[2], in[3]), in[0]);
    putc(0, bxdf_f(x, v3(4), v3(7)));
    return 3;
  }
  if (fn == "fresnel_specular_sample_f" && n_in == 12 && n_out >= 8) {
    BxDF x;
    x.kind = BX_FRESNEL_SPECULAR; x.type = BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_SPECULAR;
    x.r = RGB(in[0], in[1], in[2]); x.t = RGB(in[3], in[4], in[5]); x.etaA = 1.0; x.etaB = in[6];
    RGB f; V3 wi; double pdf; int st;
    bxdf_sample_f(x, v3(7), P2{in[10], in[11]}, &f, &wi, &pdf, &st);
    putc(0, f); put3(3, wi); out[6] = pdf; out[7] = (double)st;
    return 8;
  }
  if (fn == "concentric_sample_disk" && n_in == 2 && n_out >= 2) { P2 d = concentric_sample_disk(P2{in[0], in[1]}); out[0] = d.x; out[1] = d.y; return 2; }
  if (fn == "cosine_sample_hemisphere" && n_in == 2 && n_out >= 3) { put3(0, cosine_sample_hemisphere(P2{in[0], in[1]})); return 3; }
  if (fn == "lambert_sample_f" && n_in == 8 && n_out >= 7) {
    // BSDF.SampleF (reflection.go:183-253) over one Lambertian lobe in the identity shading frame: returns the LOCAL wi
    BSDF b;
    b.ns = V3{0, 0, 1}; b.ng = V3{0, 0, 1}; b.ss = V3{1, 0, 0}; b.ts = V3{0, 1, 0};
    BxDF x; x.kind = BX_LAMBERT; x.type = BSDF_REFLECTION | BSDF_DIFFUSE; x.r = RGB(in[0], in[1], in[2]);
    b.bx[b.n++] = x;
    RGB f; V3 wi; double pdf; int st;
    bsdf_sample_f(b, v3(3), P2{in[6], in[7]}, BSDF_ALL, &f, &wi, &pdf, &st);
    putc(0, f); put3(3, wi); out[6] = pdf;
    return 7;
  }
  if (fn == "offset_ray_origin" && n_in == 12 && n_out >= 3) { put3(0, offset_ray_origin(v3(0), v3(3), v3(6), v3(9))); return 3; }
  if (fn == "coordinate_system" && n_in == 3 && n_out >= 6) { V3 a, b; coordinate_system(v3(0), &a, &b); put3(0, a); put3(3, b); return 6; }
  if (fn == "sample_discrete_uniform" && n_in == 2 && n_out >= 2) {
    int n = (int)in[0];
    std::vector<double> cdf;
    double func_int;
    uniform_light_distribution(n, &cdf, &func_int);
    int off; double pdf;
    sample_discrete(cdf, func_int, in[1], &off, &pdf);
    out[0] = off; out[1] = pdf;
    return 2;
  }
  if (fn == "rgb_to_xyz" && n_in == 3 && n_out >= 3) {
    // through the film: one tile holding the value, merged into a one-pixel film (film.go:115-132 + spectrum.go:35-41)
    gopbrt_film fc{1, 1, {0, 0, 1, 1}, {0.5, 0.5}};
    FilmCfg f = film_cfg(fc);
    FilmTile t = film_tile(f, Bounds2i{0, 0, 1, 1});
    t.px[0] = in[0]; t.px[1] = in[1]; t.px[2] = in[2];
    double film[4] = {0, 0, 0, 0};
    film_merge(f, t, film);
    out[0] = film[0]; out[1] = film[1]; out[2] = film[2];
    return 3;
  }
  if (fn == "film_add_sample" && n_in == 11) {
    gopbrt_film fc{(int)in[0], (int)in[1], {0, 0, 1, 1}, {in[3], in[4]}};
    FilmCfg f = film_cfg(fc);
    int64_t ts = (int64_t)in[2], tile = (int64_t)in[5];
    int64_t ntx = ((int64_t)in[0] + ts - 1) / ts, tx = tile % ntx, ty = tile / ntx;
    int64_t x0 = tx * ts, y0 = ty * ts;
    int64_t x1 = (int64_t)gm::Min((double)(x0 + ts), (double)f.cropped.x1), y1 = (int64_t)gm::Min((double)(y0 + ts), (double)f.cropped.y1);
    FilmTile t = film_tile(f, Bounds2i{x0, y0, x1, y1});
    film_add_sample(f, t, P2{in[6], in[7]}, RGB(in[8], in[9], in[10]), 1.0);
    int k = 0;
    if (n_out < 4) return -1;
    out[k++] = (double)t.pb.x0; out[k++] = (double)t.pb.y0; out[k++] = (double)t.pb.x1; out[k++] = (double)t.pb.y1;
    int64_t w = t.pb.x1 - t.pb.x0;
    for (int64_t y = t.pb.y0; y < t.pb.y1; y++)
      for (int64_t x = t.pb.x0; x < t.pb.x1; x++) {
        const double* p = &t.px[((x - t.pb.x0) + (y - t.pb.y0) * w) * 4];
        if (p[3] == 0) continue;  // pixels of the tile the sample did not touch
        if (k + 6 > n_out) return -1;
        out[k++] = (double)x; out[k++] = (double)y; out[k++] = p[0]; out[k++] = p[1]; out[k++] = p[2]; out[k++] = p[3];
      }
    return k;
  }
  if (fn == "rng_u32" && n_in == 3) {
    Rng r;
    if (in[0] != 0) r.set_sequence((uint64_t)in[1]);
    int n = (int)in[2];
    if (n > n_out) return -1;
    for (int i = 0; i < n; i++) out[i] = (double)r.u32();
    return n;
  }
  if (fn == "rng_uniform" && n_in == 2) {
    Rng r; r.set_sequence((uint64_t)in[0]);
    int n = (int)in[1];
    if (n > n_out) return -1;
    for (int i = 0; i < n; i++) out[i] = r.uniform();
    return n;
  }
  if (fn == "rng_u32b" && n_in == 3) {
    Rng r; r.set_sequence((uint64_t)in[0]);
    int b = (int)in[1], n = (int)in[2];
    if (n > n_out) return -1;
    for (int i = 0; i < n; i++) out[i] = (double)r.u32b((uint32_t)(b - i));
    return n;
  }
  if (fn == "stratified_start_pixel" && n_in == 5) {
    gopbrt_sampler cfg{GOPBRT_SAMPLER_STRATIFIED, (int)in[1], (int)in[2], (int)in[3], (int)in[4], GOPBRT_MODE_STRICT};
    Sampler s;
    s.init(cfg);
    s.clone_seed((uint64_t)in[0]);
    s.start_pixel();
    int k = 0;
    for (auto& t : s.s1d) for (double v : t) { if (k >= n_out) return -1; out[k++] = v; }
    for (auto& t : s.s2d) for (const P2& v : t) if (v.x != 0 || v.y != 0) return -1;  // the 2-D tables stay all zeros
    if (k + 2 > n_out) return -1;
    out[k++] = s.rng.uniform(); out[k++] = s.rng.uniform();
    return k;
  }
  if (fn == "light_sample_li" && n_in == 12 && n_out >= 17 && scene) {
    const Scene* sc = (const Scene*)scene;
    int li = (int)in[0];
    if (li < 0 || li >= (int)sc->lights.size()) return -1;
    LightSample ls;
    light_sample_li(*sc, sc->lights[li], Intr{v3(1), v3(4), v3(7)}, P2{in[10], in[11]}, &ls);
    putc(0, ls.Li); put3(3, ls.wi); out[6] = ls.pdf; put3(7, ls.p1.p); put3(10, ls.p1.perr); put3(13, ls.p1.n); out[16] = ls.delta ? 1.0 : 0.0;
    return 17;
  }
  if (fn == "spawn_ray_to" && n_in == 18 && n_out >= 7) {
    Ray r = spawn_ray_to(Intr{v3(0), v3(3), v3(6)}, Intr{v3(9), v3(12), v3(15)}, 0);
    put3(0, r.o); put3(3, r.d); out[6] = r.tmax;
    return 7;
  }
  if (fn == "go_math" && n_in == 2 && n_out >= 6) {
    out[0] = gomath::Max(in[0], in[1]); out[1] = gomath::Min(in[0], in[1]); out[2] = gomath::Sin(in[0]); out[3] = gomath::Cos(in[0]); out[4] = gomath::Sin(in[0]); out[5] = gomath::Cos(in[0]);
    return 6;
  }
  if (fn == "camera_ray" && n_in == 38 && n_out >= 6) {
    gopbrt_camera c;
    for (int i = 0; i < 16; i++) { c.raster_to_camera[i] = in[i]; c.camera_to_world[i] = in[16 + i]; }
    c.lens_radius = in[32]; c.focal_distance = in[33]; c.shutter_open = 0; c.shutter_close = 0;
    Ray r = camera_ray(c, P2{in[34], in[35]}, P2{in[36], in[37]}, 0.0);
    put3(0, r.o); put3(3, r.d);
    return 6;
  }
  // ---- the reference's own unit-test vectors for the vector / spectrum / flag / partition helpers the path is made of
  // (pkg/geometry/xyz_test.go, pkg/pbrt/spectrum_test.go, pkg/pbrt/reflection_test.go:9-15, pkg/accelerator/bvh_test.go:143-264;
  // replayed by tests/test_oracle_golden.py)
  if (fn == "xyz_op" && n_in == 8 && n_out >= 3) {  // in: op, a.xyz, b.xyz, scalar
    int op = (int)in[0];
    V3 a = v3(1), b = v3(4);
    double s = in[7];
    switch (op) {
      case 0: put3(0, vabs(a)); return 3;                                  // Abs            xyz.go
      case 1: out[0] = absdot(a, b); return 1;                             // AbsDot
      case 2: put3(0, add(a, b)); return 3;                                // Add / AddAssign
      case 3: put3(0, V3{a.x + s, a.y + s, a.z + s}); return 3;            // AddConst
      case 4: put3(0, cross(a, b)); return 3;                              // Cross
      case 5: out[0] = dist(a, b); return 1;                               // Distance
      case 6: out[0] = dist2(a, b); return 1;                              // DistanceSquared
      case 7: put3(0, V3{a.x / b.x, a.y / b.y, a.z / b.z}); return 3;      // Div / DivAssign
      case 8: put3(0, divs(a, s)); return 3;                               // DivScalar
      case 9: out[0] = dot(a, b); return 1;                                // Dot
      case 10: out[0] = length(a); return 1;                               // Length
      case 11: out[0] = len2(a); return 1;                                 // LengthSquared
      case 12: put3(0, mul(a, b)); return 3;                               // Mul / MulAssign
      case 13: put3(0, muls(a, s)); return 3;                              // MulScalar
      case 14: put3(0, normalized(a)); return 3;                           // Normalize / Normalized
      case 15: put3(0, sub(a, b)); return 3;                               // Sub / SubAssign
      case 16: out[0] = (a.x == b.x && a.y == b.y && a.z == b.z) ? 1 : 0; return 1;  // Equals (NotEquals = !)
      case 17: out[0] = a[(int)s]; return 1;                               // Index
    }
    return -1;
  }
  if (fn == "spectrum_op" && n_in == 8 && n_out >= 3) {  // in: op, a.rgb, b.rgb, scalar
    int op = (int)in[0];
    RGB a(in[1], in[2], in[3]), b(in[4], in[5], in[6]);
    double s = in[7];
    switch (op) {
      case 0: putc(0, sadd(a, b)); return 3;                                           // Add / AddAssign   spectrum.go
      case 1: putc(0, RGB(a.c[0] + s, a.c[1] + s, a.c[2] + s)); return 3;              // AddScalar
      case 2: putc(0, sdivs(a, s)); return 3;                                          // DivScalar
      case 3: putc(0, smul(a, b)); return 3;                                           // Mul
      case 4: out[0] = sblack(a) ? 1 : 0; return 1;                                    // IsBlack
    }
    return -1;
  }
  if (fn == "matches_flags" && n_in == 2 && n_out >= 1) { out[0] = matches((int)in[0], (int)in[1]) ? 1 : 0; return 1; }  // reflection.go:301-303
  if (fn == "partition_at" && n_in >= 4) {  // in: start, end, pivot, centroid.x of every element (its primitive number too); bvh.go:163-175
    int n = n_in - 3;
    if (n_out < n + 1) return -1;
    std::vector<BuildInfo> info(n);
    for (int i = 0; i < n; i++) { info[i].prim = (int)in[3 + i]; info[i].c = V3{in[3 + i], 0, 0}; }
    int64_t r = partition_at(info, (int64_t)in[0], (int64_t)in[1], (int64_t)in[2],
                             [](const BuildInfo& a, const BuildInfo& b) { return a.c[0] < b.c[0]; });
    for (int i = 0; i < n; i++) out[i] = (double)info[i].prim;
    out[n] = (double)r;
    return n + 1;
  }
  return -1;
}
}
