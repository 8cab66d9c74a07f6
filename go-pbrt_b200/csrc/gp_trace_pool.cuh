// gp_trace_pool.cuh — pooled BVH traversal: the `extend` (closest hit == BVH.Intersect, pkg/accelerator/bvh.go:659-712)
// and `shadow` (any hit == BVH.IntersectP, bvh.go:713-765) kernels with the rays of a warp held in SHARED MEMORY.
//
// Why: a ray's work comes in kinds of very different cost — an inner-node step (~100 instructions), the leaf's
// own-bound candidate test (~60), the float64 watertight triangle test (~250) and the EFloat sphere/disk test
// (~1500).  With one ray per thread (k_trace, gp_trace.cuh) the lanes of a warp are spread over those kinds and ncu
// shows 9-11 of 32 lanes active per issued instruction on every scene.  Here a warp owns a POOL of 32*SPL ray slots
// whose traversal state lives in shared memory, and every scheduler round it picks ONE kind of work, gathers up to 32
// slots waiting for exactly that kind and runs it with a (nearly) full warp:
//
//   X  retire finished rays and refill the freed slots from the queue (one atomicAdd per <= 256 rays per warp)
//   N  one traversal step: fetch the two sibling child records of the current inner node (64 contiguous bytes, float4
//      loads), test both float32 boxes, near child first by split axis and ray sign, the far child's record index on
//      the stack — or pop one stacked record and (closest hit) re-test its box against the ray's current tMax
//   L  walk the leaf's candidates: a primitive is admitted iff ITS OWN float64 world bound passes Bounds3.IntersectP
//      with the running tMax (the topology-independent parity spec); the first admitted one parks the slot
//   T  the parked triangle test (float64 watertight)
//   Q  the parked sphere/disk test (float64 + EFloat, sphere.go:64-268 / disk.go:64-159)
//
// A slot is a sequential state machine, so the per-ray visit order, candidate order and running tMax are exactly those
// of k_trace (and of the reference): only the interleaving BETWEEN rays changes.  All scheduling decisions are
// warp-uniform (ballots over the slot states); warps never synchronise with each other.
//
// Shared memory per slot: o, 1/d, tMax, triangle shear constants (10 doubles), 8 control words, and a stack of
// stack_cap 4-byte record indices, all laid out [field][slot] so a chunk's accesses spread over the banks.
#pragma once
#include "gp_trace.cuh"

namespace gp {

#ifndef GP_POOL_SPL
#define GP_POOL_SPL 2        // ray slots per lane: a warp owns 32*SPL slots
#endif
#ifndef GP_POOL_DSTEPS
#define GP_POOL_DSTEPS 1     // traversal steps an N round runs before the warp re-schedules
#endif
#ifndef GP_POOL_REFILL
#define GP_POOL_REFILL 16    // finished/free slots that trigger a retire + refill round
#endif
constexpr int kPoolSPL = GP_POOL_SPL;
constexpr int kPoolSlots = 32 * kPoolSPL;
constexpr int kPoolWarps = kTraceThreads / 32;
constexpr int kPoolDSteps = GP_POOL_DSTEPS;

enum : unsigned { PS_FREE = 0, PS_N = 1, PS_L = 2, PS_T = 3, PS_Q = 4, PS_DONE = 5 };
// double fields
enum : int { PD_OX = 0, PD_OY, PD_OZ, PD_IX, PD_IY, PD_IZ, PD_TMAX, PD_SX, PD_SY, PD_SZ, PD_N };
// word fields
enum : int { PU_CURA = 0, PU_CURB, PU_LEAFA, PU_LEAF /* nPrims<<16 | next */, PU_REC, PU_META, PU_LANE, PU_PEND, PU_N };
// meta word: bits 0-2 state, 3-10 sp, 11/12/13 sign of 1/d per axis, 14-15 kz, 16-17 shade class of the hit, 18 hit_any, 19 have_cur
constexpr unsigned PM_STATE = 7u, PM_SP_SHIFT = 3, PM_SP = 0xffu << 3, PM_NX = 1u << 11, PM_NY = 1u << 12, PM_NZ = 1u << 13,
                   PM_KZ_SHIFT = 14, PM_CLS_SHIFT = 16, PM_CLS = 3u << 16, PM_HIT = 1u << 18, PM_CUR = 1u << 19;

__host__ __device__ inline size_t pool_warp_bytes(int stack_cap) { return (size_t)kPoolSlots * (PD_N * 8 + PU_N * 4 + 4 * (size_t)stack_cap) + 32; }
inline size_t pool_smem_bytes(int stack_cap) { return kPoolWarps * pool_warp_bytes(stack_cap); }

template <int MODE, bool COUNT>
__global__ void __launch_bounds__(kTraceThreads, 4) k_trace_pool(DevScene sc, RayRec* __restrict__ rays, const ShadowRec* __restrict__ srays,
                                                                 PathRec* __restrict__ paths, unsigned char* __restrict__ occluded,
                                                                 const int* __restrict__ queue, const int* __restrict__ count, long long n_direct,
                                                                 int stack_cap, int* work_counter, TraceCounters* ctr, unsigned long long* gt10_counter) {
  constexpr bool ANY = MODE != 0;
  constexpr int NS = kPoolSlots;
  extern __shared__ __align__(16) unsigned char s_pool[];
  const unsigned FULL = 0xffffffffu;
  const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane_id) - 1u;
  unsigned char* wbase = s_pool + (size_t)warp * pool_warp_bytes(stack_cap);
  double* sd = (double*)wbase;                                   // [PD_N][NS]
  unsigned* su = (unsigned*)(wbase + (size_t)PD_N * 8 * NS);     // [PU_N][NS]
  unsigned* sstack = su + PU_N * NS;                             // [stack_cap][NS]
  unsigned char* slist = (unsigned char*)(sstack + (size_t)stack_cap * NS);  // [32]
#define SD(f, s) sd[(f) * NS + (s)]
#define SU(f, s) su[(f) * NS + (s)]

  const long long n = queue ? (long long)*count : n_direct;
  long long per_warp = n / ((long long)gridDim.x * kPoolWarps * 2);
  const int kChunk = per_warp >= kChunkMax ? kChunkMax : (per_warp <= 32 ? 32 : (int)(per_warp & ~31LL));
  long long w_next = 0, w_end = 0;  // warp-uniform chunk [w_next, w_end) of queue positions
  bool exhausted = false;
  TravCnt c = {0, 0, 0, 0, 0};
  int bad = 0, ovf = 0;
  unsigned long long gt10 = 0;

#pragma unroll
  for (int k = 0; k < kPoolSPL; k++) SU(PU_META, lane_id + 32 * k) = PS_FREE;
  __syncwarp();

  for (;;) {
    // ---- census of the warp's slots (lane l owns slots l, l+32, ...): one packed warp reduction, 8 bits per kind
    unsigned own[kPoolSPL];
    unsigned packed = 0;
    int nX = 0;
#pragma unroll
    for (int k = 0; k < kPoolSPL; k++) {
      own[k] = SU(PU_META, lane_id + 32 * k) & PM_STATE;
      packed += (own[k] >= PS_N && own[k] <= PS_Q) ? (1u << (8 * (own[k] - 1))) : 0u;
      nX += __popc(__ballot_sync(FULL, own[k] == PS_DONE));
    }
    packed = __reduce_add_sync(FULL, packed);
    const int nN = (int)(packed & 0xffu), nL = (int)((packed >> 8) & 0xffu), nT = (int)((packed >> 16) & 0xffu), nQ = (int)(packed >> 24);
    const int busy = nN + nL + nT + nQ;
    const int nFree = NS - busy - nX;
    const bool can_refill = !(exhausted && w_next >= w_end);
    const int nR = nX + (can_refill ? nFree : 0);
    // ---- this round's work: retire + refill once enough slots wait for it, else the kind with the most waiting slots
    //      (a round costs the same however many of its 32 lanes have a slot; ties go to the more expensive kind)
    unsigned pick;
    if (nR >= GP_POOL_REFILL || (nR > 0 && busy == 0)) pick = PS_DONE;
    else if (busy == 0) break;  // nothing in flight, nothing to retire, nothing left to claim
    else {
      const int cN = nN < 32 ? nN : 32, cL = nL < 32 ? nL : 32, cT = nT < 32 ? nT : 32, cQ = nQ < 32 ? nQ : 32;
      pick = PS_Q; int best = cQ;
      if (cT > best) { pick = PS_T; best = cT; }
      if (cL > best) { pick = PS_L; best = cL; }
      if (cN > best) { pick = PS_N; best = cN; }
    }
    // ---- gather up to 32 slots of the chosen kind (X takes finished slots and, while rays remain, free ones)
    int n_list = 0;
    {
      int run = 0;
#pragma unroll
      for (int k = 0; k < kPoolSPL; k++) {
        bool m = pick == PS_DONE ? (own[k] == PS_DONE || (can_refill && own[k] == PS_FREE)) : own[k] == pick;
        unsigned b = __ballot_sync(FULL, m);
        int r = run + __popc(b & lt_mask);
        if (m && r < 32) slist[r] = (unsigned char)(lane_id + 32 * k);
        run += __popc(b);
      }
      n_list = run < 32 ? run : 32;
    }
    __syncwarp();
    const int s = lane_id < n_list ? (int)slist[lane_id] : -1;

    if (pick == PS_DONE) {
      // ------------------------------------------------------------------ X: retire + refill
      if (s >= 0 && (SU(PU_META, s) & PM_STATE) == PS_DONE) {
        unsigned meta = SU(PU_META, s);
        long long lane = (long long)SU(PU_LANE, s);
        if (MODE == 0) {
          double2 out;
          out.x = SD(PD_TMAX, s);
          unsigned cls = (meta & PM_CLS) >> PM_CLS_SHIFT;
          out.y = __longlong_as_double((long long)(((unsigned long long)cls << 32) | SU(PU_REC, s)));  // {hit_rec, shade class}
          ((double2*)(rays + lane))[3] = out;
        } else if (MODE == 1) {
          occluded[lane] = (meta & PM_HIT) ? 1 : 0;
        } else {
          const ShadowRec* sr = srays + lane;
          PathRec* pt = paths + lane;
          double pr = sr->pr, pg = sr->pg, pb = sr->pb;
          if (!(meta & PM_HIT)) {
            pt->Lr += pr; pt->Lg += pg; pt->Lb += pb;
            if (sr->gt10) gt10++;
          } else {  // blocked: Li = 0, so L += beta*0 (NaN only if beta is not finite)
            pt->Lr += pr * 0.0; pt->Lg += pg * 0.0; pt->Lb += pb * 0.0;
          }
        }
      }
      int take = 0;
      if (can_refill) {
        if (w_next >= w_end) {
          int base = 0;
          if (lane_id == 0) base = atomicAdd(work_counter, kChunk);
          base = __shfl_sync(FULL, base, 0);
          if ((long long)base >= n) { exhausted = true; w_next = w_end = 0; }
          else { w_next = base; w_end = (long long)base + kChunk < n ? (long long)base + kChunk : n; }
        }
        long long avail = w_end - w_next;
        take = (long long)n_list < avail ? n_list : (int)avail;
      }
      if (s >= 0) {
        unsigned meta = PS_FREE;
        if (lane_id < take) {
          long long i = w_next + lane_id;
          long long lane = queue ? queue[i] : i;
          Ray ray;
          if (MODE == 2) {
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
          }
          V3 invd = mk3(1 / ray.d.x, 1 / ray.d.y, 1 / ray.d.z);  // bvh.go:665-666
          int nx = invd.x < 0, ny = invd.y < 0, nz = invd.z < 0;
          TriRay tray = tri_ray_setup(ray.d);
          SD(PD_OX, s) = ray.o.x; SD(PD_OY, s) = ray.o.y; SD(PD_OZ, s) = ray.o.z;
          SD(PD_IX, s) = invd.x; SD(PD_IY, s) = invd.y; SD(PD_IZ, s) = invd.z;
          SD(PD_TMAX, s) = ray.tmax;
          SD(PD_SX, s) = tray.Sx; SD(PD_SY, s) = tray.Sy; SD(PD_SZ, s) = tray.Sz;
          SU(PU_LANE, s) = (unsigned)lane;
          SU(PU_REC, s) = 0xffffffffu;
          SU(PU_LEAF, s) = 0;
          meta = PS_DONE | (nx ? PM_NX : 0u) | (ny ? PM_NY : 0u) | (nz ? PM_NZ : 0u) | ((unsigned)tray.kz << PM_KZ_SHIFT);
          if (sc.n_nodes > 0) {  // the root's own box (bvh.go:673-675)
            float4 r0 = __ldg(sc.nodes), r1 = __ldg(sc.nodes + 1);
            if (COUNT) c.nodes++;
            RayF32 rf = ray_f32(ray.o, invd);
            if (slab_test_f32_maybe(r0, r1, rf, nx, ny, nz, __double2float_ru(ray.tmax))) {
              unsigned ra = __float_as_uint(r0.w), rb = __float_as_uint(r1.w);
              if ((rb >> 8) != 0) {  // the root is a leaf
                SU(PU_LEAFA, s) = ra; SU(PU_LEAF, s) = (rb >> 8) << 16;
                meta = (meta & ~PM_STATE) | PS_L;
              } else {
                SU(PU_CURA, s) = ra; SU(PU_CURB, s) = rb;
                meta = (meta & ~PM_STATE) | PS_N | PM_CUR;
              }
            }
          }
        }
        SU(PU_META, s) = meta;
      }
      w_next += take;
    } else if (pick == PS_N) {
      // ------------------------------------------------------------------ N: ONE traversal step per slot
      // have_cur: (cur_a, cur_b) are the words of an inner node whose box has passed -> fetch its two child records
      // (64 contiguous bytes), test both boxes, near child first, far child's record index on the stack.
      // otherwise: pop one record index, fetch that record and (closest hit) re-test its box against the current tMax.
      // A slot whose next node is a leaf moves on to the L rounds.
      if (s >= 0) {
        unsigned meta = SU(PU_META, s);
        int sp = (int)((meta & PM_SP) >> PM_SP_SHIFT);
        const int nx = (meta & PM_NX) != 0, ny = (meta & PM_NY) != 0, nz = (meta & PM_NZ) != 0;
        bool have_cur = (meta & PM_CUR) != 0;
        unsigned cur_a = SU(PU_CURA, s), cur_b = SU(PU_CURB, s);
        const RayF32 rf = ray_f32(mk3(SD(PD_OX, s), SD(PD_OY, s), SD(PD_OZ, s)), mk3(SD(PD_IX, s), SD(PD_IY, s), SD(PD_IZ, s)));
        const float tub = __double2float_ru(SD(PD_TMAX, s));
#pragma unroll 1
        for (int step = 0; step < kPoolDSteps; step++) {
          if (have_cur) {
            if ((cur_b >> 8) != 0) break;  // a leaf
            const float4* pp = sc.nodes + 2 * (size_t)cur_a;
            float4 l0 = __ldg(pp), l1 = __ldg(pp + 1), r0 = __ldg(pp + 2), r1 = __ldg(pp + 3);
            if (COUNT) c.nodes += 2;
            bool pl = slab_test_f32_maybe(l0, l1, rf, nx, ny, nz, tub);
            bool pr = slab_test_f32_maybe(r0, r1, rf, nx, ny, nz, tub);
            int axis = cur_b & 3;
            int neg = axis == 0 ? nx : (axis == 1 ? ny : nz);
            unsigned la = __float_as_uint(l0.w), lb = __float_as_uint(l1.w), ra = __float_as_uint(r0.w), rb = __float_as_uint(r1.w);
            unsigned fa = neg ? ra : la, fb = neg ? rb : lb, sa = neg ? la : ra, sb = neg ? lb : rb;
            bool pf = neg ? pr : pl, ps = neg ? pl : pr;
            unsigned second_rec = neg ? cur_a : cur_a + 1;  // record index of the far child
            if (pf) {
              if (ps) {
                if (sp >= stack_cap) { ovf = 1; have_cur = false; sp = 0; break; }
                sstack[sp * NS + s] = second_rec; ++sp;
              }
              cur_a = fa; cur_b = fb;
            } else if (ps) {
              cur_a = sa; cur_b = sb;
            } else {
              have_cur = false;
            }
          } else {
            if (sp == 0) break;
            --sp;
            unsigned r = sstack[sp * NS + s];
            const float4* pp = sc.nodes + 2 * (size_t)r;
            float4 b0 = __ldg(pp), b1 = __ldg(pp + 1);
            if (COUNT) c.nodes++;
            if (ANY || slab_test_f32_maybe(b0, b1, rf, nx, ny, nz, tub)) { cur_a = __float_as_uint(b0.w); cur_b = __float_as_uint(b1.w); have_cur = true; }
          }
        }
        unsigned ns = have_cur ? (((cur_b >> 8) != 0) ? PS_L : PS_N) : (sp > 0 ? PS_N : PS_DONE);
        if (ns == PS_L) { SU(PU_LEAFA, s) = cur_a; SU(PU_LEAF, s) = (cur_b >> 8) << 16; have_cur = false; }
        SU(PU_CURA, s) = cur_a; SU(PU_CURB, s) = cur_b;
        SU(PU_META, s) = (meta & ~(PM_STATE | PM_SP | PM_CUR)) | ns | ((unsigned)sp << PM_SP_SHIFT) | (have_cur ? PM_CUR : 0u);
      }
    } else if (pick == PS_L) {
      // ------------------------------------------------------------------ L: the leaf's candidates in order
      // a primitive is admitted iff its own float64 world bound passes Bounds3.IntersectP with the running tMax; the
      // first admitted one parks the slot for a T or Q round
      if (s >= 0) {
        unsigned meta = SU(PU_META, s);
        const int nx = (meta & PM_NX) != 0, ny = (meta & PM_NY) != 0, nz = (meta & PM_NZ) != 0;
        const unsigned leaf_a = SU(PU_LEAFA, s);
        unsigned leaf = SU(PU_LEAF, s);
        unsigned ln = leaf >> 16, li = leaf & 0xffffu;
        const V3 o = mk3(SD(PD_OX, s), SD(PD_OY, s), SD(PD_OZ, s)), invd = mk3(SD(PD_IX, s), SD(PD_IY, s), SD(PD_IZ, s));
        const double tmax = SD(PD_TMAX, s);
        unsigned ns = (meta & PM_SP) != 0 ? PS_N : PS_DONE;  // leaf exhausted: back to the stack
        while (li < ln) {
          unsigned ri = leaf_a + li;
          li++;
          const PrimRec* prec = sc.recs + ri;
          uint32_t flags = prec->flags;
          if ((flags & RK_KIND_MASK) == RK_TRIANGLE) {
            // own bound = min/max of the (finite) vertices; a compare-select equals Go's Min/Max up to the sign of a
            // zero, which the slab test cannot observe
            const double2* q = (const double2*)prec;
            double2 v0 = q[0], v1 = q[1], v2 = q[2], v3 = q[3], v4 = q[4];  // {flags|prim, d0} {d1,d2} {d3,d4} {d5,d6} {d7,d8}
            double ax = v0.y, ay = v1.x, az = v1.y, bx = v2.x, by = v2.y, bz = v3.x, cx = v3.y, cy = v4.x, cz = v4.y;
            double x0 = ax < bx ? ax : bx, x1 = ax < bx ? bx : ax; x0 = cx < x0 ? cx : x0; x1 = cx > x1 ? cx : x1;
            double y0 = ay < by ? ay : by, y1 = ay < by ? by : ay; y0 = cy < y0 ? cy : y0; y1 = cy > y1 ? cy : y1;
            double z0 = az < bz ? az : bz, z1 = az < bz ? bz : az; z0 = cz < z0 ? cz : z0; z1 = cz > z1 ? cz : z1;
            if (!slab_test(x0, y0, z0, x1, y1, z1, o, invd, nx, ny, nz, tmax)) continue;
            if (COUNT) { c.prims++; c.tri++; }
            SU(PU_PEND, s) = ri; ns = PS_T;
            break;
          } else {
            const double* bb = sc.rec_bounds + (size_t)ri * 6;
            if (!slab_test(bb[0], bb[1], bb[2], bb[3], bb[4], bb[5], o, invd, nx, ny, nz, tmax)) continue;
            if (COUNT) { c.prims++; if (flags & RF_FAST) c.sph++; else c.gen++; }
            SU(PU_PEND, s) = ri; ns = PS_Q;
            break;
          }
        }
        SU(PU_LEAF, s) = (ln << 16) | li;
        SU(PU_META, s) = (meta & ~PM_STATE) | ns;
      }
    } else {
      // ------------------------------------------------------------------ T / Q: the parked shape test
      if (s >= 0) {
        unsigned meta = SU(PU_META, s);
        const unsigned ri = SU(PU_PEND, s);
        const PrimRec* prec = sc.recs + ri;
        Ray ray;
        ray.o = mk3(SD(PD_OX, s), SD(PD_OY, s), SD(PD_OZ, s));
        ray.tmax = SD(PD_TMAX, s);
        double t;
        bool hit;
        uint32_t flags;
        if (pick == PS_T) {
          const double2* q = (const double2*)prec;
          double2 v0 = q[0], v1 = q[1], v2 = q[2], v3 = q[3], v4 = q[4];
          flags = (uint32_t)(__double_as_longlong(v0.x) & 0xffffffffLL);
          V3 p0 = mk3(v0.y, v1.x, v1.y), p1 = mk3(v2.x, v2.y, v3.x), p2 = mk3(v3.y, v4.x, v4.y);
          TriRay tray;
          tray.kz = (int)((meta >> PM_KZ_SHIFT) & 3u);
          tray.kx = tray.kz + 1; if (tray.kx == 3) tray.kx = 0;
          tray.ky = tray.kx + 1; if (tray.ky == 3) tray.ky = 0;
          tray.Sx = SD(PD_SX, s); tray.Sy = SD(PD_SY, s); tray.Sz = SD(PD_SZ, s);
          ray.d = mk3(0, 0, 0);  // the watertight test reads the direction only through (k, S)
          hit = tri_test_pre(p0, p1, p2, ray, tray, &t, nullptr);
        } else {
          long long lane = (long long)SU(PU_LANE, s);
          const double2* q = MODE == 2 ? (const double2*)(srays + lane) : (const double2*)(rays + lane);
          double2 b = q[1], c2 = q[2];
          ray.d = mk3(b.y, c2.x, c2.y);
          flags = prec->flags;
          hit = quadric_test(sc, prec, flags, ray, &t, bad);
        }
        unsigned leaf = SU(PU_LEAF, s);
        unsigned ns = (leaf & 0xffffu) < (leaf >> 16) ? PS_L : ((meta & PM_SP) != 0 ? PS_N : PS_DONE);
        if (hit) {
          meta |= PM_HIT;
          if (ANY) ns = PS_DONE;
          else {
            SD(PD_TMAX, s) = t;  // r.TMax = tHit (primitive.go:51)
            SU(PU_REC, s) = ri;
            meta = (meta & ~PM_CLS) | (((flags & RF_CLASS_MASK) >> RF_CLASS_SHIFT) << PM_CLS_SHIFT);
          }
        }
        SU(PU_META, s) = (meta & ~PM_STATE) | ns;
      }
    }
    __syncwarp();
  }
#undef SD
#undef SU
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

}  // namespace gp
