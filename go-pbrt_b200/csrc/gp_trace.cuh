// gp_trace.cuh — BVH traversal kernels: `extend` (closest hit == BVH.Intersect, pkg/accelerator/bvh.go:659-712) and
// `shadow` (any hit == BVH.IntersectP, bvh.go:713-765) over coalesced SoA ray queues.
//
// One thread per ray; persistent grid (a multiple of the SM count) striding over the queue.  The traversal stack
// lives in shared memory, interleaved [entry][thread] so a warp's pushes/pops hit 32 distinct banks.
// Node fetch = two 16-byte loads of a 32-byte node.  Node boxes are float32 rounded OUTWARD from the float64 union of
// the primitives' bounds and tested with the reference's float64 slab test, so a node is entered whenever the
// reference's own (float64) box would be; at the leaf each primitive is admitted only if ITS OWN float64 world bound
// passes Bounds3.IntersectP with the running tMax — the topology-independent parity spec of SURVEY §8a — and then
// gets the float64/EFloat shape test.  Near child first by split axis and ray sign, as the reference.
#pragma once
#include "gp_scene.cuh"

namespace gp {

constexpr int kStackDepth = 64;
constexpr int kTraceThreads = 128;

struct RaySoA { double *ox, *oy, *oz, *dx, *dy, *dz, *tmax; };

struct TraceCounters {
  unsigned long long nodes, prims, snodes, sprims, efloat_panics, stack_overflows, t_tri, t_sph, t_gen;
};

struct TravCnt { unsigned long long nodes, prims, tri, sph, gen; };

template <bool ANY, bool COUNT>
GP_D bool traverse(const DevScene& sc, Ray& ray, int* hit_rec, unsigned* stack /* this thread's column */, int stride,
                   TravCnt& cnt, int& bad, int& overflow) {
  V3 invd = mk3(1 / ray.d.x, 1 / ray.d.y, 1 / ray.d.z);
  int nx = invd.x < 0, ny = invd.y < 0, nz = invd.z < 0;
  int sp = 0;
  unsigned cur = 0;
  bool any = false;
  if (sc.n_nodes == 0) return false;
  for (;;) {
    float4 n0 = __ldg(sc.nodes + 2 * (size_t)cur);
    float4 n1 = __ldg(sc.nodes + 2 * (size_t)cur + 1);
    if (COUNT) cnt.nodes++;
    unsigned a = __float_as_uint(n0.w), b = __float_as_uint(n1.w);
    if (slab_test((double)n0.x, (double)n0.y, (double)n0.z, (double)n1.x, (double)n1.y, (double)n1.z, ray.o, invd, nx, ny, nz, ray.tmax)) {
      unsigned np = b >> 8;
      if (np > 0) {
        for (unsigned i = 0; i < np; i++) {
          unsigned ri = a + i;
          const PrimRec* rec = sc.recs + ri;
          uint32_t flags = rec->flags;
          bool cand;
          if ((flags & RK_KIND_MASK) == RK_TRIANGLE) {
            const double* d = rec->d;
            double x0 = go_min(go_min(d[0], d[3]), d[6]), x1 = go_max(go_max(d[0], d[3]), d[6]);
            double y0 = go_min(go_min(d[1], d[4]), d[7]), y1 = go_max(go_max(d[1], d[4]), d[7]);
            double z0 = go_min(go_min(d[2], d[5]), d[8]), z1 = go_max(go_max(d[2], d[5]), d[8]);
            cand = slab_test(x0, y0, z0, x1, y1, z1, ray.o, invd, nx, ny, nz, ray.tmax);
          } else {
            const double* bb = sc.rec_bounds + (size_t)ri * 6;
            cand = slab_test(bb[0], bb[1], bb[2], bb[3], bb[4], bb[5], ray.o, invd, nx, ny, nz, ray.tmax);
          }
          if (!cand) continue;
          if (COUNT) {
            cnt.prims++;
            if ((flags & RK_KIND_MASK) == RK_TRIANGLE) cnt.tri++;
            else if (flags & RF_FAST) cnt.sph++;
            else cnt.gen++;
          }
          double t;
          if (prim_test(sc, rec, flags, ray, &t, bad)) {
            if (ANY) return true;
            ray.tmax = t;  // r.TMax = tHit (primitive.go:51)
            *hit_rec = (int)ri;
            any = true;
          }
        }
        if (sp == 0) break;
        cur = stack[(--sp) * stride];
      } else {
        if (sp >= kStackDepth) { overflow = 1; break; }
        int neg = (b & 3) == 0 ? nx : ((b & 3) == 1 ? ny : nz);
        if (neg) { stack[(sp++) * stride] = cur + 1; cur = a; }
        else { stack[(sp++) * stride] = a; cur = cur + 1; }
      }
    } else {
      if (sp == 0) break;
      cur = stack[(--sp) * stride];
    }
  }
  return any;
}

GP_D unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// closest hit.  queue == nullptr: ray i is lane i (batched API); otherwise lane = queue[i] for i < *count.
// out: rays.tmax[lane] = tHit (unchanged on a miss), hit_rec[lane] = leaf-record index or -1.
template <bool COUNT>
__global__ void __launch_bounds__(kTraceThreads) k_extend(DevScene sc, RaySoA rays, int* __restrict__ hit_rec, const int* __restrict__ queue,
                                                          const int* __restrict__ count, long long n_direct, TraceCounters* ctr) {
  __shared__ unsigned s_stack[kStackDepth * kTraceThreads];
  long long n = queue ? (long long)*count : n_direct;
  TravCnt c = {0, 0, 0, 0, 0};
  int bad = 0, ovf = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long lane = queue ? queue[i] : i;
    Ray r;
    r.o = mk3(rays.ox[lane], rays.oy[lane], rays.oz[lane]);
    r.d = mk3(rays.dx[lane], rays.dy[lane], rays.dz[lane]);
    r.tmax = rays.tmax[lane];
    int rec = -1;
    traverse<false, COUNT>(sc, r, &rec, s_stack + threadIdx.x, kTraceThreads, c, bad, ovf);
    rays.tmax[lane] = r.tmax;
    hit_rec[lane] = rec;
  }
  if (COUNT) {
    c.nodes = warp_sum(c.nodes); c.prims = warp_sum(c.prims); c.tri = warp_sum(c.tri); c.sph = warp_sum(c.sph); c.gen = warp_sum(c.gen);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&ctr->nodes, c.nodes); atomicAdd(&ctr->prims, c.prims);
      atomicAdd(&ctr->t_tri, c.tri); atomicAdd(&ctr->t_sph, c.sph); atomicAdd(&ctr->t_gen, c.gen);
    }
  }
  if (bad) atomicAdd(&ctr->efloat_panics, 1ULL);
  if (ovf) atomicAdd(&ctr->stack_overflows, 1ULL);
}

// any hit over the shadow-ray queue.  occluded[lane] = 1/0.
template <bool COUNT>
__global__ void __launch_bounds__(kTraceThreads) k_anyhit(DevScene sc, RaySoA rays, unsigned char* __restrict__ occluded,
                                                          const int* __restrict__ queue, const int* __restrict__ count, long long n_direct,
                                                          TraceCounters* ctr) {
  __shared__ unsigned s_stack[kStackDepth * kTraceThreads];
  long long n = queue ? (long long)*count : n_direct;
  TravCnt c = {0, 0, 0, 0, 0};
  int bad = 0, ovf = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long lane = queue ? queue[i] : i;
    Ray r;
    r.o = mk3(rays.ox[lane], rays.oy[lane], rays.oz[lane]);
    r.d = mk3(rays.dx[lane], rays.dy[lane], rays.dz[lane]);
    r.tmax = rays.tmax[lane];
    int rec = -1;
    bool hit = traverse<true, COUNT>(sc, r, &rec, s_stack + threadIdx.x, kTraceThreads, c, bad, ovf);
    occluded[lane] = hit ? 1 : 0;
  }
  if (COUNT) {
    c.nodes = warp_sum(c.nodes); c.prims = warp_sum(c.prims); c.tri = warp_sum(c.tri); c.sph = warp_sum(c.sph); c.gen = warp_sum(c.gen);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&ctr->snodes, c.nodes); atomicAdd(&ctr->sprims, c.prims);
      atomicAdd(&ctr->t_tri, c.tri); atomicAdd(&ctr->t_sph, c.sph); atomicAdd(&ctr->t_gen, c.gen);
    }
  }
  if (bad) atomicAdd(&ctr->efloat_panics, 1ULL);
  if (ovf) atomicAdd(&ctr->stack_overflows, 1ULL);
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
