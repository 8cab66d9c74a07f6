// gp_build.cuh — on-device BVH build of libgopbrt_cuda: the GPU form of accelerator.NewBVH (pkg/accelerator/bvh.go:223-270).
//
// The reference builds its tree recursively on one goroutine (RecursiveBuild, bvh.go:272-411).  Here the tree is built
// breadth-first, one tree LEVEL per round, every node of the level and every primitive in it at once:
//   bin      every primitive of an unfinished node files its box under (node, axis, bin of its centroid): 16 bins x 3 axes,
//            counts by atomicAdd, bin boxes by atomicMin / atomicMax on order-preserving integer images of the floats;
//   split    one thread per node sweeps the 3 x 16 bins for the cheapest surface-area-heuristic plane (the same binned SAH
//            as the host builder in gp_bvh.h; a node whose centroids coincide, or a level past 32, is halved by position);
//   scatter  a device-wide prefix sum of the "goes right" flags gives every primitive its place in its child's range; boxes,
//            centroids and indices move there together, so the next level streams them in order; the children's boxes and
//            centroid bounds are accumulated on the way.
// Nothing is ever gathered through an index: a level costs a few sequential passes over n x 44 bytes.  The finished binary
// tree is laid out in the same 4-record child groups the traversal kernels read (gp_bvh.h / gp_trace.cuh), groups numbered
// level by level, and the leaf-ordered primitive records are written from the uploaded vertex / triangle arrays.
// Results (hit / miss, primitive, t) do not depend on the tree (DESIGN §2 parity spec): every node box is the exact union of
// float32 boxes rounded OUTWARD from the primitives' float64 bounds, so no node ever culls a primitive whose own bound passes.
// Everything is deterministic: node ids, child-group ids and leaf order come from prefix sums, never from atomic counters.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "gp_bvh.h"
#include "gp_scene.cuh"

namespace gpbuild {

using gp::PrimRec;

constexpr int kBins = 16;
constexpr int kBinWords = 7;                       // count, lo.xyz, hi.xyz
constexpr int kNodeBinWords = 3 * kBins * kBinWords;  // 336 words per unfinished node and level
constexpr unsigned kDone = 0xffffffffu;            // node_of[] of a position whose leaf is final
constexpr int kHalveFromLevel = 32;                // SAH above, halving by position below: depth <= 32 + log2(n) < kStackDepth

// order-preserving images: a < b  <=>  ord(a) < ord(b) as unsigned integers (-0 < +0, as in Go's math.Min / math.Max)
__host__ __device__ __forceinline__ unsigned ord_f(float f) {
#ifdef __CUDA_ARCH__
  unsigned u = __float_as_uint(f);
#else
  unsigned u; memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float unord_f(unsigned o) {
  unsigned u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
__device__ __forceinline__ unsigned long long ord_d(double d) {
  unsigned long long u = (unsigned long long)__double_as_longlong(d);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __forceinline__ double unord_d(unsigned long long o) {
  unsigned long long u = (o & 0x8000000000000000ull) ? (o & 0x7fffffffffffffffull) : ~o;
  double d; memcpy(&d, &u, 8); return d;
}
constexpr unsigned kOrdMaxInit = 0u, kOrdMinInit = 0xffffffffu;  // identities of atomicMax / atomicMin on the images

// monotone atomics with a read first: most primitives do not move a bound that many others have already pushed
__device__ __forceinline__ void amin(unsigned* p, unsigned v) { if (v < *(volatile unsigned*)p) atomicMin(p, v); }
__device__ __forceinline__ void amax(unsigned* p, unsigned v) { if (v > *(volatile unsigned*)p) atomicMax(p, v); }
__device__ __forceinline__ void amin64(unsigned long long* p, unsigned long long v) { if (v < *(volatile unsigned long long*)p) atomicMin(p, v); }
__device__ __forceinline__ void amax64(unsigned long long* p, unsigned long long v) { if (v > *(volatile unsigned long long*)p) atomicMax(p, v); }

// ---------------------------------------------------------------- device-wide exclusive prefix sum (uint32)
// Three launches per level of a block-sum pyramid: per-block totals, the (recursive) scan of those totals, the per-block
// scan seeded with them.  1024 threads x 4 items per block.
constexpr int kScanThreads = 1024, kScanItems = 4, kScanBlock = kScanThreads * kScanItems;

__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* total) {  // v: this thread's sum
  __shared__ unsigned s_warp[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned w = s_warp[lane];
    unsigned winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
    s_warp[lane] = winc - w;  // exclusive over the warps
    if (lane == 31 && total) *total = winc;
  }
  __syncthreads();
  unsigned r = s_warp[warp] + inc - v;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const unsigned* __restrict__ in, unsigned* __restrict__ block_sums, long long n) {
  long long base = (long long)blockIdx.x * kScanBlock + (long long)threadIdx.x * kScanItems;
  unsigned v = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) if (base + k < n) v += in[base + k];
  __shared__ unsigned s_total;
  block_exclusive_scan(v, &s_total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = s_total;
}
// single block: exclusive scan of up to kScanBlock values in place; *total (may be null) receives the sum
__global__ void __launch_bounds__(kScanThreads) k_scan_small(unsigned* data, long long n, unsigned* total) {
  long long base = (long long)threadIdx.x * kScanItems;
  unsigned x[kScanItems], v = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) { x[k] = base + k < n ? data[base + k] : 0u; v += x[k]; }
  __shared__ unsigned s_total;
  unsigned off = block_exclusive_scan(v, &s_total);
#pragma unroll
  for (int k = 0; k < kScanItems; k++) { if (base + k < n) data[base + k] = off; off += x[k]; }
  if (threadIdx.x == 0 && total) *total = s_total;
}
__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const unsigned* in, unsigned* out, const unsigned* __restrict__ block_offsets, long long n)  /* in may alias out */ {
  long long base = (long long)blockIdx.x * kScanBlock + (long long)threadIdx.x * kScanItems;
  unsigned x[kScanItems], v = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) { x[k] = base + k < n ? in[base + k] : 0u; v += x[k]; }
  unsigned off = block_exclusive_scan(v, nullptr) + block_offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; k++) { if (base + k < n) out[base + k] = off; off += x[k]; }
}
// out[i] = sum of in[0..i); *total = sum of all (device memory).  tmp: at least scan_tmp_words(n) words.
inline size_t scan_tmp_words(long long n) {
  size_t w = 0;
  while (n > kScanBlock) { n = (n + kScanBlock - 1) / kScanBlock; w += (size_t)n; }
  return w + 1;
}
inline void exclusive_scan(const unsigned* in, unsigned* out, long long n, unsigned* tmp, unsigned* total, cudaStream_t st, uint64_t* launches) {
  if (n <= 0) { cudaMemsetAsync(total, 0, sizeof(unsigned), st); return; }
  if (n <= kScanBlock) {
    if (in != out) cudaMemcpyAsync(out, in, (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToDevice, st);
    k_scan_small<<<1, kScanThreads, 0, st>>>(out, n, total);
    if (launches) *launches += 1;
    return;
  }
  long long nb = (n + kScanBlock - 1) / kScanBlock;
  k_scan_reduce<<<(unsigned)nb, kScanThreads, 0, st>>>(in, tmp, n);
  if (launches) *launches += 1;
  exclusive_scan(tmp, tmp, nb, tmp + nb, total, st, launches);  // block offsets in place; the grand total falls out at the top
  k_scan_apply<<<(unsigned)nb, kScanThreads, 0, st>>>(in, out, tmp, n);
  if (launches) *launches += 1;
}

// ---------------------------------------------------------------- build state
struct Nodes {         // structure of arrays over node ids (root = 0; the children of a node are left, left + 1)
  unsigned* start;     // first position of the node's primitives
  unsigned* count;
  unsigned* left;      // id of the left child; 0 while the node is a leaf
  unsigned* meta;      // split axis (bits 0-1) | split bin (bits 8-15; 255 = halved by position) | level (bits 16-23)
  unsigned* box;       // 6 words per node: order images of lo.xyz, hi.xyz (float32, outward)
  unsigned* cbox;      // 6 words per node: order images of the centroid bounds
};
struct Pos {           // per position (double-buffered): the primitive there, its box and centroid, the node it is in
  float4* a;           // lo.xyz, centroid.x
  float4* b;           // hi.xyz, centroid.y
  float* c;            // centroid.z
  unsigned* idx;       // primitive index
  unsigned* node;      // node id, or kDone once the position belongs to a finished leaf
};

struct PrepOut {       // written by k_prim_bounds
  unsigned long long world[6];  // order images of the float64 world bound (Go Min / Max semantics)
  unsigned root_box[6], root_cbox[6];
  int error;           // first validation failure (see kErr*)
};
enum { kErrNone = 0, kErrVertexIndex = 1, kErrTriangleIndex = 2, kErrMaterialIndex = 3, kErrXfOnTriangle = 4, kErrNonFinite = 5, kErrKind = 6 };

// Per primitive: float64 world bound -> float32 outward box + centroid, the world bound, the root's box.
// Triangles are bounded here from the uploaded vertices (Go Min / Max of finite values = compare-select, up to the sign of
// a zero, which the order images restore); spheres / disks arrive with the bound the host computed in the reference's
// arithmetic (Transform.TransformBounds, transform.go:336-345): qbounds[6 * qslot[i]].
__global__ void k_prim_bounds(const int4* __restrict__ prims, long long n, const double* __restrict__ vertices, long long n_vertices,
                              const int4* __restrict__ triangles, long long n_triangles, const int* __restrict__ qslot,
                              const double* __restrict__ qbounds, int n_materials, Pos P, PrepOut* out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
  bool ok = false;
  int err = 0;
  if (i < n) {
    int4 p = prims[i];
    if (p.z >= n_materials) err = kErrMaterialIndex;
    if (p.x == GOPBRT_SHAPE_TRIANGLE) {
      if (p.y < 0 || p.y >= n_triangles) err = kErrTriangleIndex;
      else if (p.w >= 0) err = kErrXfOnTriangle;
      else {
        int4 t = triangles[p.y];
        if (t.x < 0 || t.y < 0 || t.z < 0 || t.x >= n_vertices || t.y >= n_vertices || t.z >= n_vertices) err = kErrVertexIndex;
        else {
          const int v[3] = {t.x, t.y, t.z};
          for (int k = 0; k < 3; k++) {
            double a = vertices[3 * (size_t)v[0] + k], b = vertices[3 * (size_t)v[1] + k], c = vertices[3 * (size_t)v[2] + k];
            double lo = a < b ? a : b, hi = a < b ? b : a;
            mn[k] = c < lo ? c : lo; mx[k] = c > hi ? c : hi;
          }
          ok = true;
        }
      }
    } else if (p.x == GOPBRT_SHAPE_SPHERE || p.x == GOPBRT_SHAPE_DISK) {
      const double* q = qbounds + 6 * (size_t)qslot[i];
      for (int k = 0; k < 3; k++) { mn[k] = q[k]; mx[k] = q[3 + k]; }
      ok = true;
    } else err = kErrKind;
    if (ok) for (int k = 0; k < 3; k++) if (!(fabs(mn[k]) < 1.7e308) || !(fabs(mx[k]) < 1.7e308)) { ok = false; err = kErrNonFinite; }
  }
  if (err) atomicCAS(&out->error, 0, err);
  if (i < n) {
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}, ce[3] = {0, 0, 0};
    if (ok) for (int k = 0; k < 3; k++) { lo[k] = __double2float_rd(mn[k]); hi[k] = __double2float_ru(mx[k]); ce[k] = (float)(0.5 * mn[k] + 0.5 * mx[k]); }
    P.a[i] = make_float4(lo[0], lo[1], lo[2], ce[0]);
    P.b[i] = make_float4(hi[0], hi[1], hi[2], ce[1]);
    P.c[i] = ce[2];
    P.idx[i] = (unsigned)i;
    P.node[i] = 0;
    if (ok) {
      for (int k = 0; k < 3; k++) {
        amin64(&out->world[k], ord_d(mn[k])); amax64(&out->world[3 + k], ord_d(mx[k]));
        amin(&out->root_box[k], ord_f(lo[k])); amax(&out->root_box[3 + k], ord_f(hi[k]));
        amin(&out->root_cbox[k], ord_f(ce[k])); amax(&out->root_cbox[3 + k], ord_f(ce[k]));
      }
    }
  }
}

__global__ void k_init_root(Nodes N, const PrepOut* prep, unsigned n) {
  N.start[0] = 0; N.count[0] = n; N.left[0] = 0; N.meta[0] = 0;
  for (int k = 0; k < 6; k++) { N.box[k] = prep->root_box[k]; N.cbox[k] = prep->root_cbox[k]; }
}

// level [lb, le): act[i - lb] = 1 where node i still has to be split
__global__ void k_level_active(Nodes N, unsigned lb, unsigned le, unsigned max_prims, unsigned* __restrict__ act) {
  unsigned i = lb + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < le) act[i - lb] = N.count[i] > max_prims ? 1u : 0u;
}
// act_list[rank] = node id, for the unfinished nodes of the level; rank_of[i - lb] = rank
__global__ void k_level_list(unsigned lb, unsigned le, const unsigned* __restrict__ act, const unsigned* __restrict__ rank, unsigned* __restrict__ act_list) {
  unsigned i = lb + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < le && act[i - lb]) act_list[rank[i - lb]] = i;
}

__device__ __forceinline__ int bin_of(float c, float lo, float scale) {
  int k = (int)((c - lo) * scale);
  return k < 0 ? 0 : (k >= kBins ? kBins - 1 : k);
}

// bins[word][active rank]: word = (axis * kBins + bin) * kBinWords + {0: count, 1-3: ~image of lo, 4-6: image of hi}.
// The lo rows hold the COMPLEMENT of the order image, so that min becomes max and an all-zero memset is the identity of
// every row (count 0, lo = +inf side, hi = -inf side).
__global__ void k_bin(Pos P, Nodes N, long long n, unsigned lb, const unsigned* __restrict__ act, const unsigned* __restrict__ rank,
                      unsigned* __restrict__ bins, unsigned n_act) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  unsigned node = P.node[p];
  if (node == kDone || node < lb || !act[node - lb]) return;
  const unsigned a = rank[node - lb];
  float4 A = P.a[p], B = P.b[p];
  const float ce[3] = {A.w, B.w, P.c[p]};
  const unsigned lo[3] = {ord_f(A.x), ord_f(A.y), ord_f(A.z)}, hi[3] = {ord_f(B.x), ord_f(B.y), ord_f(B.z)};
  const unsigned* cb = N.cbox + 6 * (size_t)node;
#pragma unroll
  for (int ax = 0; ax < 3; ax++) {
    float clo = unord_f(cb[ax]), chi = unord_f(cb[3 + ax]);
    float ext = chi - clo;
    if (!(ext > 0.f)) continue;
    int k = bin_of(ce[ax], clo, (float)kBins / ext);
    unsigned* w = bins + (size_t)((ax * kBins + k) * kBinWords) * n_act + a;
    atomicAdd(w, 1u);
#pragma unroll
    for (int c = 0; c < 3; c++) { amax(w + (size_t)(1 + c) * n_act, ~lo[c]); amax(w + (size_t)(4 + c) * n_act, hi[c]); }
  }
}

__device__ __forceinline__ float box_area(const float* lo, const float* hi) {
  float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
  if (!(dx >= 0.f) || !(dy >= 0.f) || !(dz >= 0.f)) return 0.f;
  return 2.f * (dx * dy + dx * dz + dy * dz);
}

// one thread per unfinished node: binned SAH over 3 x 16 bins (cost = area_left * n_left + area_right * n_right, as the host
// builder), creation of the two children (ids next + 2 * rank, next + 2 * rank + 1; ranges; empty boxes for the scatter pass)
__global__ void k_split(Nodes N, const unsigned* __restrict__ act_list, unsigned n_act, const unsigned* __restrict__ bins, unsigned next, unsigned level) {
  unsigned a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_act) return;
  const unsigned node = act_list[a];
  const unsigned start = N.start[node], cnt = N.count[node];
  float best = INFINITY;
  int best_axis = 0, best_bin = -1;
  unsigned best_left = 0;
  if (level < (unsigned)kHalveFromLevel) {
    for (int ax = 0; ax < 3; ax++) {
      float clo = unord_f(N.cbox[6 * (size_t)node + ax]), chi = unord_f(N.cbox[6 * (size_t)node + 3 + ax]);
      if (!(chi - clo > 0.f)) continue;
      float ra[kBins];
      unsigned rc[kBins];
      float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
      unsigned c = 0;
      for (int k = kBins - 1; k > 0; k--) {
        const unsigned* w = bins + (size_t)((ax * kBins + k) * kBinWords) * n_act + a;
        unsigned bc = w[0];
        if (bc) for (int q = 0; q < 3; q++) { lo[q] = fminf(lo[q], unord_f(~w[(size_t)(1 + q) * n_act])); hi[q] = fmaxf(hi[q], unord_f(w[(size_t)(4 + q) * n_act])); }
        c += bc;
        ra[k] = box_area(lo, hi); rc[k] = c;
      }
      for (int q = 0; q < 3; q++) { lo[q] = INFINITY; hi[q] = -INFINITY; }
      c = 0;
      for (int k = 0; k < kBins - 1; k++) {
        const unsigned* w = bins + (size_t)((ax * kBins + k) * kBinWords) * n_act + a;
        unsigned bc = w[0];
        if (bc) for (int q = 0; q < 3; q++) { lo[q] = fminf(lo[q], unord_f(~w[(size_t)(1 + q) * n_act])); hi[q] = fmaxf(hi[q], unord_f(w[(size_t)(4 + q) * n_act])); }
        c += bc;
        if (c == 0 || rc[k + 1] == 0) continue;
        float cost = box_area(lo, hi) * (float)c + ra[k + 1] * (float)rc[k + 1];
        if (cost < best) { best = cost; best_bin = k; best_axis = ax; best_left = c; }
      }
    }
  }
  unsigned sbin = 255u, n_left = cnt / 2;
  if (best_bin >= 0) { sbin = (unsigned)best_bin; n_left = best_left; }
  else {  // halved by position: the split "axis" that orders the two children for the traversal is the widest centroid axis
    float e[3];
    for (int q = 0; q < 3; q++) e[q] = unord_f(N.cbox[6 * (size_t)node + 3 + q]) - unord_f(N.cbox[6 * (size_t)node + q]);
    best_axis = (e[0] > e[1] && e[0] > e[2]) ? 0 : (e[1] > e[2] ? 1 : 2);
  }
  const unsigned l = next + 2 * a;
  N.left[node] = l;
  N.meta[node] = (unsigned)best_axis | (sbin << 8) | (level << 16);
  N.start[l] = start; N.count[l] = n_left; N.left[l] = 0; N.meta[l] = (level + 1) << 16;
  N.start[l + 1] = start + n_left; N.count[l + 1] = cnt - n_left; N.left[l + 1] = 0; N.meta[l + 1] = (level + 1) << 16;
  for (int c = 0; c < 2; c++)
    for (int q = 0; q < 3; q++) {
      N.box[6 * (size_t)(l + c) + q] = kOrdMinInit; N.box[6 * (size_t)(l + c) + 3 + q] = kOrdMaxInit;
      N.cbox[6 * (size_t)(l + c) + q] = kOrdMinInit; N.cbox[6 * (size_t)(l + c) + 3 + q] = kOrdMaxInit;
    }
}

// flag[p] = 1 where position p goes to the right child of its (unfinished) node
__global__ void k_side(Pos P, Nodes N, long long n, unsigned lb, const unsigned* __restrict__ act, unsigned* __restrict__ flag) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  unsigned node = P.node[p], f = 0;
  if (node != kDone && node >= lb && act[node - lb]) {
    unsigned meta = N.meta[node];
    unsigned sbin = (meta >> 8) & 255u;
    if (sbin == 255u) f = ((unsigned)p - N.start[node]) >= N.count[N.left[node]] ? 1u : 0u;
    else {
      int ax = meta & 3u;
      float ce = ax == 0 ? P.a[p].w : (ax == 1 ? P.b[p].w : P.c[p]);
      float clo = unord_f(N.cbox[6 * (size_t)node + ax]), chi = unord_f(N.cbox[6 * (size_t)node + 3 + ax]);
      f = bin_of(ce, clo, (float)kBins / (chi - clo)) > (int)sbin ? 1u : 0u;
    }
  }
  flag[p] = f;
}

// moves every position to its place in the next level's order (stable inside each child), stamps it with its child's id (or
// kDone when that child is a leaf already), and accumulates the children's boxes and centroid bounds
__global__ void k_scatter(Pos P, Pos Q, Nodes N, long long n, unsigned lb, const unsigned* __restrict__ act, const unsigned* __restrict__ flag,
                          const unsigned* __restrict__ scan, unsigned max_prims) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  unsigned node = P.node[p];
  float4 A = P.a[p], B = P.b[p];
  float C = P.c[p];
  unsigned id = P.idx[p];
  long long dst = p;
  unsigned stamp = node;
  if (node != kDone && node >= lb && act[node - lb]) {
    const unsigned start = N.start[node], l = N.left[node], n_left = N.count[l];
    const unsigned right_before = scan[p] - scan[start];
    const unsigned f = flag[p];
    dst = f ? (long long)start + n_left + right_before : (long long)start + ((unsigned)p - start - right_before);
    const unsigned child = l + f;
    stamp = N.count[child] > max_prims ? child : kDone;
    unsigned* bx = N.box + 6 * (size_t)child;
    unsigned* cb = N.cbox + 6 * (size_t)child;
    amin(bx + 0, ord_f(A.x)); amin(bx + 1, ord_f(A.y)); amin(bx + 2, ord_f(A.z));
    amax(bx + 3, ord_f(B.x)); amax(bx + 4, ord_f(B.y)); amax(bx + 5, ord_f(B.z));
    amin(cb + 0, ord_f(A.w)); amin(cb + 1, ord_f(B.w)); amin(cb + 2, ord_f(C));
    amax(cb + 3, ord_f(A.w)); amax(cb + 4, ord_f(B.w)); amax(cb + 5, ord_f(C));
  } else if (node != kDone && node >= lb) {
    stamp = kDone;  // a node of this level that is a leaf
  }
  Q.a[dst] = A; Q.b[dst] = B; Q.c[dst] = C; Q.idx[dst] = id; Q.node[dst] = stamp;
}

// ---------------------------------------------------------------- child groups (the layout gp_trace.cuh reads)
// An interior node at an EVEN level owns a group of four records: the children of its left child (or that child itself + an
// empty slot when it is a leaf) and likewise on the right (gp_bvh.h flatten_quads).  gflag marks those nodes; their group
// index is 1 + the prefix sum (group 0 = the root's own record and three unused slots).
__global__ void k_group_flags(Nodes N, unsigned n_nodes, unsigned* __restrict__ gflag) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_nodes) gflag[i] = (N.left[i] != 0 && (((N.meta[i] >> 16) & 1u) == 0)) ? 1u : 0u;
}
__device__ __forceinline__ unsigned node_word(const Nodes& N, unsigned x, const unsigned* __restrict__ grank) {
  if (N.left[x] == 0) return gpbvh::leaf_word(N.start[x], (int)N.count[x]);
  unsigned l = N.left[x];
  unsigned axl = N.left[l] ? (N.meta[l] & 3u) : 0u, axr = N.left[l + 1] ? (N.meta[l + 1] & 3u) : 0u;
  return gpbvh::inner_word(1u + grank[x], (int)(N.meta[x] & 3u), (int)axl, (int)axr);
}
__device__ __forceinline__ void put_record(float4* out, size_t rec, const Nodes& N, unsigned x, const unsigned* __restrict__ grank) {
  const unsigned* b = N.box + 6 * (size_t)x;
  out[2 * rec] = make_float4(unord_f(b[0]), unord_f(b[1]), unord_f(b[2]), __uint_as_float(node_word(N, x, grank)));
  out[2 * rec + 1] = make_float4(unord_f(b[3]), unord_f(b[4]), unord_f(b[5]), __uint_as_float(N.count[x]));
}
__device__ __forceinline__ void put_empty(float4* out, size_t rec) {
  out[2 * rec] = make_float4(INFINITY, INFINITY, INFINITY, __uint_as_float(gpbvh::kEmptyWord));
  out[2 * rec + 1] = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
}
__global__ void k_emit_groups(Nodes N, unsigned n_nodes, const unsigned* __restrict__ gflag, const unsigned* __restrict__ grank, float4* __restrict__ out) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { put_record(out, 0, N, 0, grank); put_empty(out, 1); put_empty(out, 2); put_empty(out, 3); }
  if (i >= n_nodes || !gflag[i]) return;
  const size_t g = 4 * (size_t)(1u + grank[i]);
  const unsigned l = N.left[i], r = l + 1;
  if (N.left[l]) { put_record(out, g, N, N.left[l], grank); put_record(out, g + 1, N, N.left[l] + 1, grank); }
  else { put_record(out, g, N, l, grank); put_empty(out, g + 1); }
  if (N.left[r]) { put_record(out, g + 2, N, N.left[r], grank); put_record(out, g + 3, N, N.left[r] + 1, grank); }
  else { put_record(out, g + 2, N, r, grank); put_empty(out, g + 3); }
}

// ---------------------------------------------------------------- leaf-ordered primitive records
// rec r = the primitive at position r of the final order.  Triangles are written from the vertex arrays; spheres / disks copy
// the record the host prepared (qrecs[qslot[prim]], flags and shade class included) and their float64 bound.
__global__ void k_make_records(const unsigned* __restrict__ order, long long n, const int4* __restrict__ prims, const double* __restrict__ vertices,
                               const int4* __restrict__ triangles, const unsigned char* __restrict__ mat_not_lambert, int n_materials,
                               const int* __restrict__ qslot, const PrimRec* __restrict__ qrecs, const double* __restrict__ qbounds,
                               PrimRec* __restrict__ recs, double* __restrict__ rec_bounds, unsigned* __restrict__ class_mask) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const unsigned pi = order[r];
  const int4 p = prims[pi];
  PrimRec rec;
  if (p.x == GOPBRT_SHAPE_TRIANGLE) {
    const int4 t = triangles[p.y];
    const int v[3] = {t.x, t.y, t.z};
    for (int k = 0; k < 3; k++) for (int c = 0; c < 3; c++) rec.d[3 * k + c] = vertices[3 * (size_t)v[k] + c];
    const bool lambert = p.z >= 0 && p.z < n_materials && !mat_not_lambert[p.z];
    rec.flags = (uint32_t)gp::RK_TRIANGLE | gp::RF_FAST | (t.w ? gp::RF_REVERSE : 0) | ((lambert ? 0u : 1u) << gp::RF_CLASS_SHIFT);
    rec.prim = pi;
  } else {
    rec = qrecs[qslot[pi]];
    const double* q = qbounds + 6 * (size_t)qslot[pi];
    for (int k = 0; k < 6; k++) rec_bounds[6 * (size_t)r + k] = q[k];
  }
  recs[r] = rec;
  const unsigned m = 1u << ((rec.flags & gp::RF_CLASS_MASK) >> gp::RF_CLASS_SHIFT);
  if (!(*(volatile unsigned*)class_mask & m)) atomicOr(class_mask, m);
}

// ---------------------------------------------------------------- host driver
template <class T>
struct Buf {
  T* p = nullptr;
  cudaError_t alloc(size_t count) { return cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T)); }
  ~Buf() { if (p) cudaFree(p); }
};

struct Input {                 // everything lives in DEVICE memory
  const int4* prims = nullptr;        // gopbrt_primitive verbatim: {shape_kind, shape_index, material, prim_to_world}
  long long n = 0;
  const double* vertices = nullptr;   long long n_vertices = 0;
  const int4* triangles = nullptr;    long long n_triangles = 0;   // gopbrt_triangle verbatim: {v0, v1, v2, reverse_orientation}
  const int* qslot = nullptr;         // per primitive: slot of a sphere / disk in qrecs / qbounds (unused for triangles); may be null
  const PrimRec* qrecs = nullptr;
  const double* qbounds = nullptr;
  const unsigned char* mat_not_lambert = nullptr;  int n_materials = 0;
  int max_prims = 2;
};
struct Output {
  float4* nodes = nullptr;       // cudaMalloc'ed here, owned by the caller afterwards: 2 float4 per record
  size_t n_records = 0;
  PrimRec* recs = nullptr;       // caller-allocated, n records
  double* rec_bounds = nullptr;  // caller-allocated, 6 n doubles
  double world[6] = {0, 0, 0, 0, 0, 0};
  bool world_valid = false;
  int depth = 0;
  unsigned class_mask = 0;
  uint64_t launches = 0;
  bool cuda_error = false;       // build() failed on a CUDA call (as opposed to an invalid scene description)
  // debug: the binary tree and the final order, downloaded when `keep_host` is set (GOPBRT_CHECK_BVH)
  bool keep_host = false;
  std::vector<unsigned> h_order;
};

#define GPB_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { error = std::string(#call) + ": " + cudaGetErrorString(e__); out.cuda_error = true; return false; } } while (0)

// Builds the tree over in.n primitives on the current device, on stream st.  Returns false and sets `error` on failure
// (CUDA error, invalid index in the scene description, tree deeper than max_depth).
inline bool build(const Input& in, Output& out, int max_depth, cudaStream_t st, std::string& error) {
  const long long n = in.n;
  if (n <= 0 || n > (long long)gpbvh::kMaxLeafFirst) { error = "primitive count out of range for the device builder"; return false; }
  const unsigned max_prims = (unsigned)std::max(1, std::min(in.max_prims, gpbvh::kMaxLeafPrims));
  const size_t node_cap = 2 * (size_t)n + 2;
  const unsigned act_cap = (unsigned)(n / (max_prims + 1) + 1);  // an unfinished node holds more than max_prims primitives
  Buf<float4> pa[2], pb[2];
  Buf<float> pc[2];
  Buf<unsigned> pidx[2], pnode[2], flag, scan, tmp, act, rank, act_list, bins, total;
  Buf<unsigned> n_start, n_count, n_left, n_meta, n_box, n_cbox;
  Buf<PrepOut> prep;
  for (int k = 0; k < 2; k++) {
    GPB_CUDA(pa[k].alloc(n)); GPB_CUDA(pb[k].alloc(n)); GPB_CUDA(pc[k].alloc(n)); GPB_CUDA(pidx[k].alloc(n)); GPB_CUDA(pnode[k].alloc(n));
  }
  GPB_CUDA(flag.alloc(n)); GPB_CUDA(scan.alloc(n));
  GPB_CUDA(tmp.alloc(scan_tmp_words(std::max<long long>(n, (long long)node_cap)) + 8));
  GPB_CUDA(act.alloc(node_cap)); GPB_CUDA(rank.alloc(node_cap)); GPB_CUDA(act_list.alloc(act_cap)); GPB_CUDA(total.alloc(4));
  GPB_CUDA(bins.alloc((size_t)act_cap * kNodeBinWords));
  GPB_CUDA(n_start.alloc(node_cap)); GPB_CUDA(n_count.alloc(node_cap)); GPB_CUDA(n_left.alloc(node_cap)); GPB_CUDA(n_meta.alloc(node_cap));
  GPB_CUDA(n_box.alloc(6 * node_cap)); GPB_CUDA(n_cbox.alloc(6 * node_cap));
  GPB_CUDA(prep.alloc(1));
  Nodes N{n_start.p, n_count.p, n_left.p, n_meta.p, n_box.p, n_cbox.p};
  Pos P[2] = {{pa[0].p, pb[0].p, pc[0].p, pidx[0].p, pnode[0].p}, {pa[1].p, pb[1].p, pc[1].p, pidx[1].p, pnode[1].p}};
  unsigned* h_total = nullptr;
  GPB_CUDA(cudaHostAlloc((void**)&h_total, 4 * sizeof(unsigned), cudaHostAllocDefault));
  struct HostFree { unsigned* p; ~HostFree() { cudaFreeHost(p); } } host_free{h_total};

  PrepOut h_prep;
  for (int k = 0; k < 3; k++) {
    h_prep.world[k] = ~0ull; h_prep.world[3 + k] = 0ull;
    h_prep.root_box[k] = kOrdMinInit; h_prep.root_box[3 + k] = kOrdMaxInit;
    h_prep.root_cbox[k] = kOrdMinInit; h_prep.root_cbox[3 + k] = kOrdMaxInit;
  }
  h_prep.error = 0;
  GPB_CUDA(cudaMemcpyAsync(prep.p, &h_prep, sizeof(h_prep), cudaMemcpyHostToDevice, st));
  const int T = 256;
  const unsigned gn = (unsigned)((n + T - 1) / T);
  k_prim_bounds<<<gn, T, 0, st>>>(in.prims, n, in.vertices, in.n_vertices, in.triangles, in.n_triangles, in.qslot, in.qbounds, in.n_materials, P[0], prep.p);
  k_init_root<<<1, 1, 0, st>>>(N, prep.p, (unsigned)n);
  out.launches += 2;
  GPB_CUDA(cudaMemcpyAsync(&h_prep, prep.p, sizeof(h_prep), cudaMemcpyDeviceToHost, st));
  GPB_CUDA(cudaStreamSynchronize(st));
  switch (h_prep.error) {
    case kErrNone: break;
    case kErrVertexIndex: error = "vertex index out of range"; return false;
    case kErrTriangleIndex: error = "triangle index out of range"; return false;
    case kErrMaterialIndex: error = "material index out of range"; return false;
    case kErrXfOnTriangle: error = "TransformedPrimitive around a triangle is not supported"; return false;
    case kErrNonFinite: error = "non-finite primitive bound"; return false;
    default: error = "unknown shape kind"; return false;
  }
  for (int k = 0; k < 6; k++) out.world[k] = unord_d(h_prep.world[k]);
  out.world_valid = true;

  // ---- level by level
  unsigned lb = 0, le = 1, level = 0;
  int cur = 0;
  for (;;) {
    const unsigned nl = le - lb;
    k_level_active<<<(nl + T - 1) / T, T, 0, st>>>(N, lb, le, max_prims, act.p);
    exclusive_scan(act.p, rank.p, nl, tmp.p, total.p, st, &out.launches);
    GPB_CUDA(cudaMemcpyAsync(h_total, total.p, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    GPB_CUDA(cudaStreamSynchronize(st));
    out.launches += 1;
    const unsigned n_act = h_total[0];
    if (n_act == 0) break;
    if ((int)level + 1 >= max_depth) { error = "BVH deeper than the traversal stack"; return false; }
    if (n_act > act_cap || (size_t)le + 2 * (size_t)n_act > node_cap) { error = "device BVH build: node pool exhausted"; return false; }
    k_level_list<<<(nl + T - 1) / T, T, 0, st>>>(lb, le, act.p, rank.p, act_list.p);
    GPB_CUDA(cudaMemsetAsync(bins.p, 0, (size_t)n_act * kNodeBinWords * sizeof(unsigned), st));  // the identity of every row (see k_bin)
    k_bin<<<gn, T, 0, st>>>(P[cur], N, n, lb, act.p, rank.p, bins.p, n_act);
    k_split<<<(n_act + 127) / 128, 128, 0, st>>>(N, act_list.p, n_act, bins.p, le, level);
    k_side<<<gn, T, 0, st>>>(P[cur], N, n, lb, act.p, flag.p);
    exclusive_scan(flag.p, scan.p, n, tmp.p, total.p + 1, st, &out.launches);
    k_scatter<<<gn, T, 0, st>>>(P[cur], P[cur ^ 1], N, n, lb, act.p, flag.p, scan.p, max_prims);
    out.launches += 5;
    cur ^= 1;
    lb = le; le = le + 2 * n_act; level++;
  }
  const unsigned n_nodes = le;
  out.depth = (int)level;

  // ---- child groups
  Buf<unsigned> gflag, grank;
  GPB_CUDA(gflag.alloc(n_nodes)); GPB_CUDA(grank.alloc(n_nodes));
  k_group_flags<<<(n_nodes + T - 1) / T, T, 0, st>>>(N, n_nodes, gflag.p);
  exclusive_scan(gflag.p, grank.p, n_nodes, tmp.p, total.p + 2, st, &out.launches);
  GPB_CUDA(cudaMemcpyAsync(h_total, total.p + 2, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  GPB_CUDA(cudaStreamSynchronize(st));
  const size_t n_groups = 1 + (size_t)h_total[0];
  if (n_groups > gpbvh::kMaxGroups) { error = "more child groups than a node word addresses"; return false; }
  out.n_records = 4 * n_groups;
  GPB_CUDA(cudaMalloc((void**)&out.nodes, out.n_records * 2 * sizeof(float4)));
  k_emit_groups<<<(n_nodes + T - 1) / T, T, 0, st>>>(N, n_nodes, gflag.p, grank.p, out.nodes);
  // ---- leaf-ordered records
  Buf<unsigned> cmask;
  GPB_CUDA(cmask.alloc(1));
  GPB_CUDA(cudaMemsetAsync(cmask.p, 0, sizeof(unsigned), st));
  k_make_records<<<gn, T, 0, st>>>(P[cur].idx, n, in.prims, in.vertices, in.triangles, in.mat_not_lambert, in.n_materials, in.qslot, in.qrecs, in.qbounds,
                                   out.recs, out.rec_bounds, cmask.p);
  out.launches += 3;
  GPB_CUDA(cudaMemcpyAsync(&out.class_mask, cmask.p, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  if (out.keep_host) {
    out.h_order.resize(n);
    GPB_CUDA(cudaMemcpyAsync(out.h_order.data(), P[cur].idx, (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  }
  GPB_CUDA(cudaStreamSynchronize(st));
  GPB_CUDA(cudaGetLastError());
  return true;
}

}  // namespace gpbuild
