// gp_trace.cuh — BVH traversal kernels: `extend` (closest hit == BVH.Intersect, pkg/accelerator/bvh.go:659-712) and
// `shadow` (any hit == BVH.IntersectP, bvh.go:713-765) over index queues of per-lane ray records.
//
// One thread per ray; persistent grid (a multiple of the SM count) striding over the queue.  The traversal stack
// lives in shared memory, interleaved [entry][thread] so a warp's pushes/pops hit 32 distinct banks.
// Node fetch = float4 loads of 32-byte records; an inner node's record points at a 128-byte group holding the records
// of its (up to four) grandchildren, so one fetch brings two tree levels.  Node boxes are float32 rounded OUTWARD from the
// float64 union of the primitives' bounds and tested with a conservative float32 slab test (directed rounding), so a node
// is entered whenever the reference's own (float64) box would be; at the leaf each primitive is admitted only if ITS OWN float64 world bound
// passes Bounds3.IntersectP with the running tMax — the topology-independent parity spec of SURVEY §8a — and then
// gets the float64/EFloat shape test.  Near child first by split axis and ray sign, as the reference.
#pragma once
#include "gp_scene.cuh"

namespace gp {

constexpr int kStackDepth = 64;
constexpr int kTraceThreads = 128;

struct RaySoA { double *ox, *oy, *oz, *dx, *dy, *dz, *tmax; };

// Per-lane state is array-of-records, grouped by the stage that touches it.  The wavefront's active set is SPARSE
// (a stage sees a few percent to a few tens of percent of the lanes, picked through an index queue), so a
// plane-per-field layout would pull a whole 32-byte sector for every 8-byte field; with one aligned record per lane
// every sector that is fetched is fully used, and a record moves with 16-byte vector loads/stores.
struct __align__(32) RayRec {     // extend: in o,d,tmax; out tmax (= tHit), hit_rec.  64 B = 2 sectors
  double ox, oy, oz, dx, dy, dz, tmax;
  int hit_rec, pad;
};
struct __align__(32) ShadowRec {  // shade -> shadow: the visibility segment + the light sample it gates.  96 B = 3 sectors
  double ox, oy, oz, dx, dy, dz;
  double pr, pg, pb;              // beta * Ld, added to the path radiance if the segment is unoccluded
  int gt10, pad;
  double pad2[2];
};
struct __align__(32) RadRec {     // the lane's radiance sum (+ the refraction scale), one sector, in an array of ITS OWN: it is what the
  double Lr, Lg, Lb, eta_scale;   // shadow stage reads and writes and what the film fold streams over — densely, 32 of every 32 bytes
};                                // (as a PathRec sector the fold fetched one sector of every 128-byte record: 2.6 ms per 1080p frame)
struct __align__(32) FilmRec {    // what only the raygen stage and the film fold / merge touch (and the DirectLighting segment mask),
  int pix, has_sample;            // likewise an array of its own: the shade stage then reads and writes EVERY byte of the PathRecs it
  double pad[3];                  // fetches.  pad: uniform footprint, the lane's FilmTile as one RGB sum (film_add_uniform)
};
struct __align__(32) PathRec {    // Path.Li loop state + sampler stream of the lane.  64 B = 2 sectors
  double br, bg, bb, fx;          // sector 0: throughput
  double fy;                      // sector 1: sampler stream
  unsigned long long rng_state, rng_inc;
  int bounces, sidx;
};
static_assert(sizeof(RayRec) == 64 && sizeof(ShadowRec) == 96 && sizeof(PathRec) == 64 && sizeof(RadRec) == 32 && sizeof(FilmRec) == 32, "lane record layout");

struct TraceCounters {
  unsigned long long nodes, prims, snodes, sprims, efloat_panics, stack_overflows, t_tri, t_sph, t_gen, st_tri, st_sph, st_gen;
};

struct TravCnt { unsigned long long nodes, prims, tri, sph, gen; };

GP_D unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// ---- closest hit, order-independent (DESIGN §2 "parity spec") ----
// The closest hit of a ray is the minimum tHit over all primitives whose own float64 bound AND shape test pass with the
// ray's ORIGINAL tMax; two candidates with bit-identical tHit go to the lower primitive index.  The reference mutates
// r.TMax as it goes (primitive.go:51), which gives the same hit except where two candidates' distances differ by less
// than the rounding of the slab / shape arithmetic (coplanar faces, grazing hits within the EFloat bound): there the
// reference's answer depends on its own tree's visit order, and so would this backend's on its tree, its flat table or
// its traversal schedule.  Stated this way the result is a function of the ray and the scene alone.
// The running best distance is used only to SKIP work that cannot win: nodes and candidates are culled against
// tBest * (1 + 2^-20) — a candidate whose own bound starts beyond that is farther than the best hit by far more than
// any rounding — never against tBest itself.
constexpr double kCullSlack = 1.0 + 9.5367431640625e-07;  // 1 + 2^-20
GP_D double cull_tmax(bool hit_any, double t_best, double t_orig) {
  double c = t_best * kCullSlack;
  return (hit_any && c < t_orig) ? c : t_orig;
}
GP_D bool closer_hit(const DevScene& sc, double t, unsigned ri, double t_best, int rec) {
  return t < t_best || (t == t_best && rec >= 0 && sc.recs[ri].prim < sc.recs[rec].prim);
}

constexpr int kChunkMax = 256;     // most ray indices a warp claims with one atomicAdd (fewer when the queue is short)
#ifndef GP_REFILL_IDLE
#define GP_REFILL_IDLE 16
#endif
constexpr int kRefillIdle = GP_REFILL_IDLE;     // idle lanes that trigger a refill from the warp's chunk
#ifndef GP_DESCEND_STEPS
#define GP_DESCEND_STEPS 16
#endif
constexpr int kDescendSteps = GP_DESCEND_STEPS;
#ifndef GP_QUADRIC_BATCH
#define GP_QUADRIC_BATCH 8
#endif
constexpr int kQuadricBatch = GP_QUADRIC_BATCH;   // parked sphere/disk tests that trigger their batched execution

// Node words.  A record's `a` field (and every stack entry, and the traversal's current node) is ONE packed word:
//   leaf:     first primitive record << 3 | (nPrims - 1) << 1 | 1          (nPrims <= 4, first < 2^29)
//   interior: child-group index << 7 | axisL << 5 | axisR << 3 | axis0 << 1   (group g = records 4g .. 4g+3, 128 bytes)
//   kEmptyWord marks the unused slot of a group whose left / right child is a leaf.
// One word per stacked node halves the traversal stack's shared memory (44 -> 22 KB per CTA at depth 24), which is what
// bounded the resident CTAs per SM next to the register file.
constexpr unsigned kEmptyWord = 0xffffffffu;
GP_D bool word_is_leaf(unsigned w) { return (w & 1u) != 0; }
GP_D unsigned leaf_first(unsigned w) { return w >> 3; }
GP_D unsigned leaf_count(unsigned w) { return ((w >> 1) & 3u) + 1u; }

// One traversal step from an inner node whose box has passed: fetch its 4-record child group (128 contiguous bytes:
// slots 0,1 = children of the left child, or the left child itself + an empty slot when it is a leaf; slots 2,3
// likewise for the right child), test the (up to four) float32 boxes, continue with the first passing child in the
// binary tree's near-first order — near side of the node's own split axis first, inside each side the near child of
// that side's split axis first — and stack the other passing ones so that they pop in that same order.
// `cur` is the word of the current node on entry and of the next node on exit (have_cur = false when the traversal is
// over).  Returns the number of child records tested.
GP_D int quad_step(const DevScene& sc, unsigned& cur, bool& have_cur, int& sp, unsigned* stack, int stride, int stack_cap,
                   const RayF32& rf, int nx, int ny, int nz, float tub, int& ovf) {
  const float4* pp = sc.nodes + 8 * (size_t)(cur >> 7);
  float4 c0a = __ldg(pp), c0b = __ldg(pp + 1), c1a = __ldg(pp + 2), c1b = __ldg(pp + 3);
  float4 c2a = __ldg(pp + 4), c2b = __ldg(pp + 5), c3a = __ldg(pp + 6), c3b = __ldg(pp + 7);
  const unsigned w0 = __float_as_uint(c0a.w), w1 = __float_as_uint(c1a.w), w2 = __float_as_uint(c2a.w), w3 = __float_as_uint(c3a.w);
  const bool le = w1 != kEmptyWord, re = w3 != kEmptyWord;
  bool p0 = slab_test_f32_maybe(c0a, c0b, rf, nx, ny, nz, tub);
  bool p1 = le && slab_test_f32_maybe(c1a, c1b, rf, nx, ny, nz, tub);
  bool p2 = slab_test_f32_maybe(c2a, c2b, rf, nx, ny, nz, tub);
  bool p3 = re && slab_test_f32_maybe(c3a, c3b, rf, nx, ny, nz, tub);
  const int ax0 = (cur >> 1) & 3, ax1 = (cur >> 5) & 3, ax2 = (cur >> 3) & 3;
  const bool n0 = (ax0 == 0 ? nx : (ax0 == 1 ? ny : nz)) != 0;
  const bool n1 = (ax1 == 0 ? nx : (ax1 == 1 ? ny : nz)) != 0;
  const bool n2 = (ax2 == 0 ? nx : (ax2 == 1 ? ny : nz)) != 0;
  // left side in visit order (lf first, ls second), right side likewise
  unsigned lf = n1 ? w1 : w0, ls = n1 ? w0 : w1;
  bool plf = n1 ? p1 : p0, pls = n1 ? p0 : p1;
  unsigned rf_ = n2 ? w3 : w2, rs = n2 ? w2 : w3;
  bool prf = n2 ? p3 : p2, prs = n2 ? p2 : p3;
  // o0..o3: the four children in visit order
  unsigned o0 = n0 ? rf_ : lf, o1 = n0 ? rs : ls, o2 = n0 ? lf : rf_, o3 = n0 ? ls : rs;
  bool q0 = n0 ? prf : plf, q1 = n0 ? prs : pls, q2 = n0 ? plf : prf, q3 = n0 ? pls : prs;
  const int tested = 2 + (le ? 1 : 0) + (re ? 1 : 0);
  if (sp + 3 > stack_cap) { ovf = 1; have_cur = false; sp = 0; return tested; }
  if (q3) { stack[sp * stride] = o3; ++sp; }
  if (q2) { stack[sp * stride] = o2; ++sp; }
  if (q1) { stack[sp * stride] = o1; ++sp; }
  if (q0) cur = o0;
  else if (sp > 0) { --sp; cur = stack[sp * stride]; }
  else have_cur = false;
  return tested;
}

// Persistent warps with dynamic ray replacement ("while-while" traversal):
//   refill  idle lanes take the next ray of the warp's chunk (one atomicAdd per 256 rays), so a warp is never held
//           hostage by its single longest ray;
//   phase 1 every lane descends inner nodes until it holds a leaf or its traversal ends;
//   phase 2 the lanes holding a leaf run the primitive candidates (own-bound test, then the float64/EFloat shape test)
//           together.
// Visit order per ray is the reference's: near child first by split axis and ray sign, far child on the stack;
// a leaf's primitives in order with the running tMax (closest hit) or first hit wins (any hit).
// ANY=false: rays[lane].tmax = tHit (unchanged on a miss), rays[lane].hit_rec = leaf-record index or -1; if `occluded` is
//            set (render loop) it also receives, per QUEUE POSITION, the hit's shade class or 4 for a miss (k_split_hits).
// ANY=true : occluded[lane] = 1/0.
// queue == nullptr: ray i is lane i (batched API); otherwise lane = queue[i] for i < *count.
// dynamic shared memory: stack_cap * blockDim.x unsigned (one word per stacked node), [entry][thread] (bank-conflict free).
// MODE 0: closest hit over RayRec.  MODE 1: any hit over RayRec, occluded[lane] = 1/0 (batched API).
// MODE 3: any hit over ShadowRec segments (queue entries index srays and occluded directly), occluded[e] = 1/0:
//         DirectLighting / UniformSampleAll, whose per-light segments are summed in light order by the shade stage.
// MODE 2: any hit over the render's ShadowRec queue, resolved in place: an unoccluded segment adds its deferred light
//         sample to the lane's radiance (L.AddAssign(Ld), path.go:86).
// QUADRICS = false: the scene holds triangles only (BASELINE configs 4 and 5) — the kernel carries no sphere / disk test,
//         no parked-candidate state and no EFloat code, which is what sets the register budget of the general variant.
#ifndef GP_TRACE_BLOCKS
#define GP_TRACE_BLOCKS 4
#endif
#ifndef GP_TRACE_BLOCKS_TRI
#define GP_TRACE_BLOCKS_TRI 5
#endif
template <int MODE, bool COUNT, bool QUADRICS>
__global__ void __launch_bounds__(kTraceThreads, QUADRICS ? GP_TRACE_BLOCKS : GP_TRACE_BLOCKS_TRI) k_trace(DevScene sc, RayRec* __restrict__ rays, const ShadowRec* __restrict__ srays,
                                                            RadRec* __restrict__ rads, unsigned char* __restrict__ occluded,
                                                            const int* __restrict__ queue, const int* __restrict__ count, long long n_direct,
                                                            int stack_cap, int* work_counter, TraceCounters* ctr, unsigned long long* gt10_counter) {
  constexpr bool ANY = MODE != 0;
  extern __shared__ unsigned s_stack[];
  unsigned* stack = s_stack + threadIdx.x;
  const int stride = kTraceThreads;
  const unsigned FULL = 0xffffffffu;
  const int lane_id = threadIdx.x & 31;
  const long long n = queue ? (long long)*count : n_direct;
  TravCnt c = {0, 0, 0, 0, 0};
  int bad = 0, ovf = 0;
  unsigned long long gt10 = 0;
  // chunk size: large enough to amortise the atomic, small enough that a short queue still spreads over every warp
  long long per_warp = n / ((long long)gridDim.x * (kTraceThreads / 32) * 2);
  const int kChunk = per_warp >= kChunkMax ? kChunkMax : (per_warp <= 32 ? 32 : (int)(per_warp & ~31LL));

  bool has_ray = false;
  long long lane = 0;
  int qpos = 0;  // the ray's position in the queue (MODE 0: the hit code is filed under it for k_split_hits)
  Ray ray;
  V3 invd;
  int nx = 0, ny = 0, nz = 0, sp = 0, rec = -1, pending = -1, rec_cls = 0;
  unsigned cur = 0, leaf_a = 0, leaf_n = 0, leaf_i = 0;
  unsigned leaf2 = 0;  // word of a second leaf, found while other lanes of the warp were still looking for their first (0 = none)
  bool have_cur = false;  // cur = word of a node whose box has already passed
  TriRay tray;
  tray.kx = tray.ky = tray.kz = 0; tray.Sx = tray.Sy = tray.Sz = 0;
  RayF32 rf;
  float tmax_ub = 0.f;
  bool hit_any = false;
  double t_best = 0;  // closest tHit so far (ray.tmax keeps the ray's original tMax)
  long long w_next = 0, w_end = 0;  // warp-uniform chunk [w_next, w_end)
  bool exhausted = false;

  for (;;) {
    // ---- refill
    unsigned idle = __ballot_sync(FULL, !has_ray);
    if (idle != 0 && (__popc(idle) >= kRefillIdle || idle == FULL) && !(exhausted && w_next >= w_end)) {
      if (w_next >= w_end) {
        int base = 0;
        if (lane_id == 0) base = atomicAdd(work_counter, kChunk);
        base = __shfl_sync(FULL, base, 0);
        if ((long long)base >= n) { exhausted = true; w_next = w_end = 0; }
        else { w_next = base; w_end = (long long)base + kChunk < n ? (long long)base + kChunk : n; }
      }
      long long avail = w_end - w_next;
      int need = __popc(idle);
      int take = (long long)need < avail ? need : (int)avail;
      if (!has_ray) {
        int r = __popc(idle & ((1u << lane_id) - 1u));
        if (r < take) {
          long long i = w_next + r;
          lane = queue ? queue[i] : i;
          qpos = (int)i;
          if (MODE >= 2) {
            const double2* q = (const double2*)(srays + lane);
            double2 a = q[0], b = q[1], c2 = q[2];
            ray.o = mk3(a.x, a.y, b.x);
            ray.d = mk3(b.y, c2.x, c2.y);
            ray.tmax = 1 - 0.0001;  // 1 - ShadowEpsilon (interaction.go:99)
          } else {
            const double2* q = (const double2*)(rays + lane);
            double2 a = q[0], b = q[1], c2 = q[2], d2 = q[3];
            ray.o = mk3(a.x, a.y, b.x);
            ray.d = mk3(b.y, c2.x, c2.y);
            ray.tmax = d2.x;
            // RayRec.pad bit 0 ("the lane's last sample", gp_render.cuh lane_on_last_sample) travels as bit 3 of the shade
            // class and comes out with a MISS (a hit overwrites it: the shade stage works it out from the PathRec)
            if (MODE == 0) rec_cls = (int)((unsigned long long)__double_as_longlong(d2.y) >> 32 & 1ULL) << 3;
          }
          invd = mk3(1 / ray.d.x, 1 / ray.d.y, 1 / ray.d.z);  // bvh.go:665-666
          nx = invd.x < 0; ny = invd.y < 0; nz = invd.z < 0;
          tray = tri_ray_setup(ray.d);
          rf = ray_f32(ray.o, invd);
          tmax_ub = __double2float_ru(ray.tmax);
          t_best = ray.tmax;
          sp = 0; rec = -1; hit_any = false; pending = -1; leaf_n = 0; leaf_i = 0;
          have_cur = false;
          leaf2 = 0;
          // r.TMax <= 0 (or NaN): no shape test can return a hit (every one of them rejects t <= 0 and t >= tMax) and the
          // reference's root test already fails tMin < r.TMax unless the origin is inside — the ray is a miss without traversal
          if (sc.n_nodes > 0 && ray.tmax > 0) {  // the root's own box (bvh.go:673-675)
            float4 r0 = __ldg(sc.nodes), r1 = __ldg(sc.nodes + 1);
            if (COUNT) c.nodes++;
            if (slab_test_f32_maybe(r0, r1, rf, nx, ny, nz, tmax_ub)) { cur = __float_as_uint(r0.w); have_cur = true; }
          }
          has_ray = true;
        }
      }
      w_next += take;
    }
    if (__ballot_sync(FULL, has_ray) == 0) {
      if (exhausted && w_next >= w_end) break;
      continue;
    }
    // ---- phase 1: descend to the next leaf (lanes that still have leaf candidates or a deferred test skip this).
    //      One step (quad_step) fetches the 128-byte group of the current node's grandchildren and tests their boxes.
    //      The warp leaves the phase once no lane is still LOOKING for a leaf; until then a lane that has found its leaf
    //      keeps descending to its next one (leaf2) instead of idling.  That is speculative only in the sense that the
    //      first leaf may shorten tMax: every candidate of the second leaf is still tested against the running tMax,
    //      and the closest hit over all candidates does not depend on the order the leaves were reached in (exact t
    //      ties aside, which no two tree layouts agree on anyway).
    {
      const bool in_leaf = leaf_i < leaf_n || pending >= 0;
      if (!in_leaf) {
        leaf_i = 0;
        if (leaf2) { leaf_a = leaf_first(leaf2); leaf_n = leaf_count(leaf2); leaf2 = 0; }  // (left over when a parked sphere/disk test cut phase 2 short)
        else leaf_n = 0;
      }
      for (int step = 0; step < kDescendSteps; step++) {
        const bool looking = has_ray && !in_leaf && leaf_n == 0 && have_cur;
        if (__ballot_sync(FULL, looking) == 0) break;
#ifdef GP_NO_SPEC
        if (looking) {
#else
        if (has_ray && !in_leaf && have_cur && leaf2 == 0) {
#endif
          if (word_is_leaf(cur)) {  // a leaf: keep it for phase 2, continue from the stack
            if (leaf_n == 0) { leaf_a = leaf_first(cur); leaf_n = leaf_count(cur); }
            else leaf2 = cur;
            if (sp > 0) { --sp; cur = stack[sp * stride]; }
            else have_cur = false;
          } else {
            int tested = quad_step(sc, cur, have_cur, sp, stack, stride, stack_cap, rf, nx, ny, nz, tmax_ub, ovf);
            if (COUNT) c.nodes += tested;
          }
        }
      }
    }
    // ---- phase 2: the leaf's candidates in order; triangles are tested here, a sphere/disk candidate is parked in
    //      `pending` (the lane stops at it, so the per-ray test order and running tMax stay the reference's)
    for (;;) {
      if (QUADRICS && pending >= 0) break;
      if (leaf_i >= leaf_n) {
        if (leaf2 == 0) break;
        leaf_a = leaf_first(leaf2); leaf_n = leaf_count(leaf2); leaf_i = 0; leaf2 = 0;
      }
      unsigned ri = leaf_a + leaf_i;
      leaf_i++;
      const PrimRec* prec = sc.recs + ri;
      uint32_t flags = prec->flags;
      if (!QUADRICS || (flags & RK_KIND_MASK) == RK_TRIANGLE) {
        // own float64 world bound = min/max of the (finite) vertices: a compare-select equals Go's Min/Max up to the
        // sign of a zero, which the slab test cannot observe
        const double2* q = (const double2*)prec;
        double2 v0 = q[0], v1 = q[1], v2 = q[2], v3 = q[3], v4 = q[4];  // {flags|prim, d0} {d1,d2} {d3,d4} {d5,d6} {d7,d8}
        V3 p0 = mk3(v0.y, v1.x, v1.y), p1 = mk3(v2.x, v2.y, v3.x), p2 = mk3(v3.y, v4.x, v4.y);
        double x0 = p0.x < p1.x ? p0.x : p1.x, x1 = p0.x < p1.x ? p1.x : p0.x; x0 = p2.x < x0 ? p2.x : x0; x1 = p2.x > x1 ? p2.x : x1;
        double y0 = p0.y < p1.y ? p0.y : p1.y, y1 = p0.y < p1.y ? p1.y : p0.y; y0 = p2.y < y0 ? p2.y : y0; y1 = p2.y > y1 ? p2.y : y1;
        double z0 = p0.z < p1.z ? p0.z : p1.z, z1 = p0.z < p1.z ? p1.z : p0.z; z0 = p2.z < z0 ? p2.z : z0; z1 = p2.z > z1 ? p2.z : z1;
        if (!slab_test(x0, y0, z0, x1, y1, z1, ray.o, invd, nx, ny, nz, cull_tmax(hit_any, t_best, ray.tmax))) continue;
        if (COUNT) { c.prims++; c.tri++; }
        double t;
        if (tri_test_pre(p0, p1, p2, ray, tray, &t, nullptr)) {
          if (ANY) { hit_any = true; have_cur = false; leaf_i = leaf_n; leaf2 = 0; break; }
          if (closer_hit(sc, t, ri, t_best, rec)) {
            hit_any = true;
            t_best = t;
            tmax_ub = __double2float_ru(cull_tmax(true, t, ray.tmax));
            rec = (int)ri;
            rec_cls = (int)((flags & RF_CLASS_MASK) >> RF_CLASS_SHIFT);
          }
        }
      } else {
        const double* bb = sc.rec_bounds + (size_t)ri * 6;
        if (!slab_test(bb[0], bb[1], bb[2], bb[3], bb[4], bb[5], ray.o, invd, nx, ny, nz, cull_tmax(hit_any, t_best, ray.tmax))) continue;
        if (COUNT) { c.prims++; if (flags & RF_FAST) c.sph++; else c.gen++; }
        pending = (int)ri;
      }
    }
    // ---- phase 3: the parked sphere/disk tests, run together once enough lanes hold one or nobody else can advance
    if (QUADRICS) {
      unsigned pm = __ballot_sync(FULL, pending >= 0);
      unsigned runnable = __ballot_sync(FULL, has_ray && pending < 0 && !(!have_cur && leaf_i >= leaf_n && leaf2 == 0));
      if (pm != 0 && (__popc(pm) >= kQuadricBatch || runnable == 0)) {
        if (pending >= 0) {
          const PrimRec* prec = sc.recs + pending;
          double t;
          if (quadric_test(sc, prec, prec->flags, ray, &t, bad)) {
            if (ANY) { hit_any = true; have_cur = false; leaf_i = leaf_n; leaf2 = 0; }
            else if (closer_hit(sc, t, (unsigned)pending, t_best, rec)) {
              hit_any = true;
              t_best = t;
              tmax_ub = __double2float_ru(cull_tmax(true, t, ray.tmax));
              rec = pending;
              rec_cls = (int)((prec->flags & RF_CLASS_MASK) >> RF_CLASS_SHIFT);
            }
          }
          pending = -1;
        }
      }
    }
    // ---- retire finished rays
    bool retire = has_ray && !have_cur && leaf_i >= leaf_n && leaf2 == 0 && pending < 0;
    if (retire) {
      if (MODE == 0) {
        double2 out;
        out.x = t_best;  // tHit, or the original tMax on a miss
        out.y = __longlong_as_double((long long)(((unsigned long long)(unsigned)rec_cls << 32) | (unsigned)rec));  // {hit_rec, shade class}
        ((double2*)(rays + lane))[3] = out;
        // MODE 0 with `occluded` set: one byte per QUEUE POSITION — shade class 0..3 of the hit, 4 = escaped — so that
        // the split stage streams over the queue and never touches the ray records
        if (occluded) occluded[qpos] = rec >= 0 ? (unsigned char)rec_cls : (unsigned char)(4 | rec_cls);
      } else if (MODE == 1 || MODE == 3) {
        occluded[lane] = hit_any ? 1 : 0;
      } else {
        const ShadowRec* sr = srays + lane;
        RadRec* pt = rads + lane;
        double pr = sr->pr, pg = sr->pg, pb = sr->pb;
        if (!hit_any) {
          pt->Lr += pr; pt->Lg += pg; pt->Lb += pb;
          if (sr->gt10) gt10++;
        } else if (!sr->pad) {  // Path, blocked: Li = 0, so L += beta*0 (NaN only if beta is not finite);
          pt->Lr += pr * 0.0; pt->Lg += pg * 0.0; pt->Lb += pb * 0.0;  // DirectLighting (pad = 1) adds nothing at all
        }
      }
      has_ray = false;
    }
  }
  if (COUNT) {
    c.nodes = warp_sum(c.nodes); c.prims = warp_sum(c.prims); c.tri = warp_sum(c.tri); c.sph = warp_sum(c.sph); c.gen = warp_sum(c.gen);
    if (lane_id == 0) {
      if (ANY) {
        atomicAdd(&ctr->snodes, c.nodes); atomicAdd(&ctr->sprims, c.prims);
        atomicAdd(&ctr->st_tri, c.tri); atomicAdd(&ctr->st_sph, c.sph); atomicAdd(&ctr->st_gen, c.gen);
      } else {
        atomicAdd(&ctr->nodes, c.nodes); atomicAdd(&ctr->prims, c.prims);
        atomicAdd(&ctr->t_tri, c.tri); atomicAdd(&ctr->t_sph, c.sph); atomicAdd(&ctr->t_gen, c.gen);
      }
    }
  }
  if (bad) atomicAdd(&ctr->efloat_panics, 1ULL);
  if (ovf) atomicAdd(&ctr->stack_overflows, 1ULL);
  if (MODE == 2 && gt10) atomicAdd(gt10_counter, gt10);
}

// ---- flat aggregate for small scenes ----------------------------------------------------------------------------
// With a few dozen primitives (BASELINE configs 1 and 2: 23 and 36) a tree buys nothing on a 32-wide machine: its rays
// leave the node loop at different steps and the warp runs at a third of its lanes.  For scenes of at most kFlatMax
// primitives the aggregate is therefore a flat table — the brute-force list of the reference's accelerator.Simple
// (pkg/accelerator/simple.go:47-79) behind a conservative filter: every lane tests ALL primitive bounds (float32,
// rounded outward, the same superset test as the BVH nodes) in one warp-uniform loop over a table held in shared
// memory, which leaves a bit mask of candidates; the candidates then get what they get in the tree: the own float64
// bound test with the running tMax (the parity spec of SURVEY §8a) and the float64/EFloat shape test, triangles first,
// spheres/disks after, so that the lanes of a warp run the same kind of test together.  The closest hit over the
// candidates does not depend on the order they are tested in (exact t ties aside).
constexpr int kFlatMax = 64;

// resident CTAs per SM: 4 for the closest-hit kernel (128 registers; at 96 it spills: config 2 extend 30.9 -> 33.1 ms), 5 for
// the any-hit kernels, which carry no best-hit state (shadow 10.5 -> 10.0 ms)
template <int MODE, bool COUNT>
__global__ void __launch_bounds__(kTraceThreads, MODE == 0 ? 4 : 5) k_trace_flat(DevScene sc, RayRec* __restrict__ rays, const ShadowRec* __restrict__ srays,
                                                                 RadRec* __restrict__ rads, unsigned char* __restrict__ occluded,
                                                                 const int* __restrict__ queue, const int* __restrict__ count, long long n_direct,
                                                                 int stack_cap, int* work_counter, TraceCounters* ctr, unsigned long long* gt10_counter) {
  constexpr bool ANY = MODE != 0;
  __shared__ float4 s_tab[2 * kFlatMax];
  __shared__ float4 s_grp[2 * kFlatMax];  // the distinct boxes of the table, each with the mask of the entries that have it
  __shared__ double s_bnd[kFlatMax][6];  // every entry's own float64 world bound (a triangle's: the min / max of its vertices)
  __shared__ double s_vtx[kFlatMax][9];  // a triangle entry's vertices: the watertight test reads them ALREADY PERMUTED (tri_test_idx)
  const int nf = sc.n_flat, ng = sc.n_flat_groups;
  for (int i = threadIdx.x; i < 2 * nf; i += blockDim.x) s_tab[i] = sc.flat[i];
  for (int i = threadIdx.x; i < 2 * ng; i += blockDim.x) s_grp[i] = sc.flat[2 * nf + i];
  for (int k = threadIdx.x; k < nf; k += blockDim.x) {
    const unsigned ri = __float_as_uint(sc.flat[2 * k].w);
    if ((sc.flat_tri_mask >> k) & 1ULL) {
      const double* d = sc.recs[ri].d;
      for (int a = 0; a < 3; a++) {
        const double v0 = d[a], v1 = d[3 + a], v2 = d[6 + a];
        double lo = v0 < v1 ? v0 : v1, hi = v0 < v1 ? v1 : v0;
        lo = v2 < lo ? v2 : lo; hi = v2 > hi ? v2 : hi;
        s_bnd[k][a] = lo; s_bnd[k][3 + a] = hi;
      }
      for (int a = 0; a < 9; a++) s_vtx[k][a] = d[a];
    } else {
      for (int a = 0; a < 6; a++) s_bnd[k][a] = sc.rec_bounds[(size_t)ri * 6 + a];
    }
  }
  __syncthreads();
  const long long n = queue ? (long long)*count : n_direct;
  const unsigned long long tri_mask = sc.flat_tri_mask;
  TravCnt c = {0, 0, 0, 0, 0};
  int bad = 0;
  unsigned long long gt10 = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long lane = queue ? queue[i] : i;
    Ray ray;
    int last_bit = 0;
    if (MODE >= 2) {
      const double2* q = (const double2*)(srays + lane);
      double2 a = q[0], b = q[1], c2 = q[2];
      ray.o = mk3(a.x, a.y, b.x);
      ray.d = mk3(b.y, c2.x, c2.y);
      ray.tmax = 1 - 0.0001;  // 1 - ShadowEpsilon (interaction.go:99)
    } else {
      const double2* q = (const double2*)(rays + lane);
      double2 a = q[0], b = q[1], c2 = q[2], d2 = q[3];
      ray.o = mk3(a.x, a.y, b.x);
      ray.d = mk3(b.y, c2.x, c2.y);
      ray.tmax = d2.x;
      if (MODE == 0) last_bit = (int)((unsigned long long)__double_as_longlong(d2.y) >> 32 & 1ULL) << 3;  // see k_trace
    }
    const V3 invd = mk3(1 / ray.d.x, 1 / ray.d.y, 1 / ray.d.z);  // bvh.go:665-666
    const int nx = invd.x < 0, ny = invd.y < 0, nz = invd.z < 0;
    const RayF32 rf = ray_f32(ray.o, invd);
    float tmax_ub = __double2float_ru(ray.tmax);
    unsigned long long mask = 0;
    if (ray.tmax > 0) {  // see k_trace: a ray with tMax <= 0 (or NaN) cannot hit anything
#pragma unroll 4
      for (int g = 0; g < ng; g++) {  // one test per DISTINCT box (the two triangles of a quad mostly share theirs)
        const float4 g0 = s_grp[2 * g], g1 = s_grp[2 * g + 1];
        const unsigned long long gm = ((unsigned long long)__float_as_uint(g1.w) << 32) | __float_as_uint(g0.w);
        mask |= slab_test_f32_maybe(g0, g1, rf, nx, ny, nz, tmax_ub) ? gm : 0ULL;
      }
      if (COUNT) c.nodes += ng;
    }
    int rec = -1, rec_cls = last_bit;
    bool hit_any = false;
    double t_best = ray.tmax;  // see closer_hit: ray.tmax keeps the original tMax
    // triangles
    unsigned long long m = mask & tri_mask;
    if (m) {
      const TriRay tray = tri_ray_setup(ray.d);
      while (m) {
        const int k = __ffsll((long long)m) - 1;
        m &= m - 1;
        const float4 t0 = s_tab[2 * k], t1 = s_tab[2 * k + 1];
        if (!ANY && hit_any && !slab_test_f32_maybe(t0, t1, rf, nx, ny, nz, tmax_ub)) continue;  // behind the hit found meanwhile
        const double* bb = s_bnd[k];
        if (!slab_test(bb[0], bb[1], bb[2], bb[3], bb[4], bb[5], ray.o, invd, nx, ny, nz, cull_tmax(hit_any, t_best, ray.tmax))) continue;
        if (COUNT) { c.prims++; c.tri++; }
        const unsigned ri = __float_as_uint(t0.w);
        double t;
        if (tri_test_idx(s_vtx[k], ray, tray, &t)) {
          if (ANY) { hit_any = true; break; }
          if (closer_hit(sc, t, ri, t_best, rec)) {
            hit_any = true;
            t_best = t;
            tmax_ub = __double2float_ru(cull_tmax(true, t, ray.tmax));
            rec = (int)ri;
            rec_cls = (int)((__float_as_uint(t1.w) & RF_CLASS_MASK) >> RF_CLASS_SHIFT);
          }
        }
      }
    }
    // spheres and disks
    m = (ANY && hit_any) ? 0ULL : (mask & ~tri_mask);
    while (m) {
      const int k = __ffsll((long long)m) - 1;
      m &= m - 1;
      const float4 t0 = s_tab[2 * k], t1 = s_tab[2 * k + 1];
      if (!ANY && hit_any && !slab_test_f32_maybe(t0, t1, rf, nx, ny, nz, tmax_ub)) continue;
      const unsigned ri = __float_as_uint(t0.w);
      const uint32_t flags = __float_as_uint(t1.w);
      const double* bb = s_bnd[k];
      if (!slab_test(bb[0], bb[1], bb[2], bb[3], bb[4], bb[5], ray.o, invd, nx, ny, nz, cull_tmax(hit_any, t_best, ray.tmax))) continue;
      if (COUNT) { c.prims++; if (flags & RF_FAST) c.sph++; else c.gen++; }
      double t;
      if (quadric_test(sc, sc.recs + ri, flags, ray, &t, bad)) {
        if (ANY) { hit_any = true; break; }
        if (closer_hit(sc, t, ri, t_best, rec)) {
          hit_any = true;
          t_best = t;
          tmax_ub = __double2float_ru(cull_tmax(true, t, ray.tmax));
          rec = (int)ri;
          rec_cls = (int)((flags & RF_CLASS_MASK) >> RF_CLASS_SHIFT);
        }
      }
    }
    // retire (same record layout as k_trace)
    if (MODE == 0) {
      double2 out;
      out.x = t_best;
      out.y = __longlong_as_double((long long)(((unsigned long long)(unsigned)rec_cls << 32) | (unsigned)rec));
      ((double2*)(rays + lane))[3] = out;
      if (occluded) occluded[i] = rec >= 0 ? (unsigned char)rec_cls : (unsigned char)(4 | rec_cls);
    } else if (MODE == 1 || MODE == 3) {
      occluded[lane] = hit_any ? 1 : 0;
    } else {
      const ShadowRec* sr = srays + lane;
      RadRec* pt = rads + lane;
      double pr = sr->pr, pg = sr->pg, pb = sr->pb;
      if (!hit_any) {
        pt->Lr += pr; pt->Lg += pg; pt->Lb += pb;
        if (sr->gt10) gt10++;
      } else if (!sr->pad) {
        pt->Lr += pr * 0.0; pt->Lg += pg * 0.0; pt->Lb += pb * 0.0;
      }
    }
  }
  if (COUNT) {
    c.nodes = warp_sum(c.nodes); c.prims = warp_sum(c.prims); c.tri = warp_sum(c.tri); c.sph = warp_sum(c.sph); c.gen = warp_sum(c.gen);
    if ((threadIdx.x & 31) == 0) {
      if (ANY) {
        atomicAdd(&ctr->snodes, c.nodes); atomicAdd(&ctr->sprims, c.prims);
        atomicAdd(&ctr->st_tri, c.tri); atomicAdd(&ctr->st_sph, c.sph); atomicAdd(&ctr->st_gen, c.gen);
      } else {
        atomicAdd(&ctr->nodes, c.nodes); atomicAdd(&ctr->prims, c.prims);
        atomicAdd(&ctr->t_tri, c.tri); atomicAdd(&ctr->t_sph, c.sph); atomicAdd(&ctr->t_gen, c.gen);
      }
    }
  }
  if (bad) atomicAdd(&ctr->efloat_panics, 1ULL);
  if (MODE == 2 && gt10) atomicAdd(gt10_counter, gt10);
}

// SoA <-> record packing for the batched API (host arrays are SoA float64)
__global__ void k_pack_rays(RaySoA in, RayRec* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    RayRec r;
    r.ox = in.ox[i]; r.oy = in.oy[i]; r.oz = in.oz[i]; r.dx = in.dx[i]; r.dy = in.dy[i]; r.dz = in.dz[i]; r.tmax = in.tmax[i];
    r.hit_rec = -1; r.pad = 0;
    out[i] = r;
  }
}
// prim[i] = primitive id (or -1), t[i] = tHit
__global__ void k_unpack_hits(DevScene sc, const RayRec* __restrict__ in, int* __restrict__ prim, int* __restrict__ rec_out, double* __restrict__ t, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int r = in[i].hit_rec;
    if (prim) prim[i] = r >= 0 ? (int)sc.recs[r].prim : -1;
    if (rec_out) rec_out[i] = r;
    t[i] = in[i].tmax;
  }
}

// world-space hit point and geometric normal for the batched Aggregate.Intersect API (gopbrt_trace_closest)
__global__ void k_hit_points(DevScene sc, RaySoA rays_in /* original rays (tmax = tHit after extend) */, const int* __restrict__ hit_rec,
                             long long n, int* __restrict__ prim_out, double* __restrict__ p_out, double* __restrict__ n_out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int rec = hit_rec[i];
    int prim = -1;
    Hit h;
    h.p = mk3(0, 0, 0); h.n = mk3(0, 0, 0);
    if (rec >= 0) {
      Ray r;
      r.o = mk3(rays_in.ox[i], rays_in.oy[i], rays_in.oz[i]);
      r.d = mk3(rays_in.dx[i], rays_in.dy[i], rays_in.dz[i]);
      r.tmax = rays_in.tmax[i];
      int bad = 0;
      hit_record(sc, rec, r, r.tmax, &h, &prim, bad);
    }
    prim_out[i] = prim;
    if (p_out) { p_out[3 * i] = h.p.x; p_out[3 * i + 1] = h.p.y; p_out[3 * i + 2] = h.p.z; }
    if (n_out) { n_out[3 * i] = h.n.x; n_out[3 * i + 1] = h.n.y; n_out[3 * i + 2] = h.n.z; }
  }
}

}  // namespace gp
