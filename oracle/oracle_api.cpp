// ORACLE — test infrastructure only.  Nothing under go-pbrt_b200/ may include, link or call this.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load liboracle.so.
//
// C entry points (ctypes) over oracle_core.h / oracle_render.h.  Built by oracle/Makefile:
//   g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -pthread
#include <chrono>
#include <cstring>

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
}
