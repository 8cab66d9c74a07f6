// gp_bvh.h — host-side BVH builder of libgopbrt_cuda (the "upload once" part of accelerator.NewBVH, bvh.go:223-270).
//
// The reference's RecursiveBuild is degenerate (SplitSAH only ever uses buckets 0 and 11, depth ~N/2, and
// its [64]-entry traversal stack overflows beyond ~100 primitives), so the tree is this backend's own: binned SAH
// (16 bins, all three axes), median fallback, bounded depth, built top-down with the large subtrees fanned out over
// host threads, flattened into 32-byte node records in groups of four grandchildren (see flatten_quads).
// Any tree that never culls a primitive whose own bound passes reproduces the reference's hits, and the
// node boxes written here are float32 rounded OUTWARD of the float64 union, so they never do.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#ifdef GPBVH_TIMING
#include <chrono>
#include <cstdio>
#endif

namespace gpbvh {

#ifdef GPBVH_TIMING
#define GPBVH_TICK(label) do { auto now_ = std::chrono::steady_clock::now(); fprintf(stderr, "[gpbvh] %-22s %.3f s\n", label, std::chrono::duration<double>(now_ - tick_).count()); tick_ = now_; } while (0)
#define GPBVH_TICK_INIT auto tick_ = std::chrono::steady_clock::now()
#else
#define GPBVH_TICK(label) do {} while (0)
#define GPBVH_TICK_INIT do {} while (0)
#endif

struct Box { double mn[3], mx[3]; };
struct Node32 { float mn[3]; uint32_t a; float mx[3]; uint32_t b; };  // the 32-byte device node
static_assert(sizeof(Node32) == 32, "node must be 32 bytes");

struct BNode {
  Box b;
  int64_t left = -1, right = -1;  // indices into the same vector
  int64_t first = 0;
  int n = 0;
  int axis = 0;
};

static inline void box_init(Box& b) { for (int k = 0; k < 3; k++) { b.mn[k] = INFINITY; b.mx[k] = -INFINITY; } }
static inline void box_add(Box& b, const Box& o) {
  for (int k = 0; k < 3; k++) { if (o.mn[k] < b.mn[k]) b.mn[k] = o.mn[k]; if (o.mx[k] > b.mx[k]) b.mx[k] = o.mx[k]; }
}
static inline double box_area(const Box& b) {
  double dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
  if (!(dx >= 0) || !(dy >= 0) || !(dz >= 0)) return 0;
  return 2 * (dx * dy + dx * dz + dy * dz);
}
static inline float round_down(double d) { float f = (float)d; if ((double)f > d) f = std::nextafterf(f, -INFINITY); return f; }
static inline float round_up(double d) { float f = (float)d; if ((double)f < d) f = std::nextafterf(f, INFINITY); return f; }

struct Builder {
  const Box* bounds;       // per primitive
  std::vector<float> cen;  // centroids, 3 per primitive
  std::vector<uint32_t> idx;
  int max_prims = 4;
  int max_depth_seen = 0;

  // builds [start,end) into the node pool `out` from *used on (a range of n primitives never needs more than 2n - 1
  // nodes), returns the root's index in the pool
  int64_t build(BNode* out, int64_t* used, int64_t start, int64_t end, int depth, int* maxd) {
    int64_t me = (*used)++;
    new (&out[me]) BNode();
    if (depth > *maxd) *maxd = depth;
    Box b;
    box_init(b);
    float cmn[3] = {INFINITY, INFINITY, INFINITY}, cmx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = start; i < end; i++) {
      uint32_t p = idx[i];
      box_add(b, bounds[p]);
      for (int k = 0; k < 3; k++) { float c = cen[3 * (size_t)p + k]; if (c < cmn[k]) cmn[k] = c; if (c > cmx[k]) cmx[k] = c; }
    }
    out[me].b = b;
    int64_t n = end - start;
    int axis0 = 0;
    { float ex = cmx[0] - cmn[0], ey = cmx[1] - cmn[1], ez = cmx[2] - cmn[2]; axis0 = (ex > ey && ex > ez) ? 0 : (ey > ez ? 1 : 2); }
    if (n <= max_prims) {  // leaf: at most max_prims (<= kMaxLeafPrims) primitives, the count rides in two bits of the node word
      out[me].first = start;
      out[me].n = (int)n;
      out[me].axis = axis0;
      return me;
    }
    int64_t mid = -1;
    int best_axis = axis0;
    if (!(cmx[axis0] > cmn[axis0])) {
      // a cluster of coincident centroids: no plane separates it, so it is halved by index (logarithmic depth; the halves'
      // boxes overlap, which costs those rays extra candidates and nothing else)
      out[me].axis = axis0;
      mid = (start + end) / 2;
      int64_t l = build(out, used, start, mid, depth + 1, maxd);
      int64_t r = build(out, used, mid, end, depth + 1, maxd);
      out[me].left = l;
      out[me].right = r;
      return me;
    }
    if (depth < 40 && n > 2) {
      const int NB = 16;
      double best = INFINITY;
      int best_bin = -1;
      for (int ax = 0; ax < 3; ax++) {
        float lo = cmn[ax], ext = cmx[ax] - cmn[ax];
        if (!(ext > 0)) continue;
        Box bb[NB];
        int64_t cnt[NB];
        for (int k = 0; k < NB; k++) { box_init(bb[k]); cnt[k] = 0; }
        float scale = NB / ext;
        for (int64_t i = start; i < end; i++) {
          uint32_t p = idx[i];
          int k = (int)((cen[3 * (size_t)p + ax] - lo) * scale);
          if (k >= NB) k = NB - 1;
          if (k < 0) k = 0;
          cnt[k]++;
          box_add(bb[k], bounds[p]);
        }
        double ra[NB];
        int64_t rc[NB];
        Box acc;
        box_init(acc);
        int64_t c = 0;
        for (int k = NB - 1; k > 0; k--) { box_add(acc, bb[k]); c += cnt[k]; ra[k] = box_area(acc); rc[k] = c; }
        box_init(acc);
        c = 0;
        for (int k = 0; k < NB - 1; k++) {
          box_add(acc, bb[k]);
          c += cnt[k];
          if (c == 0 || rc[k + 1] == 0) continue;
          double cost = box_area(acc) * (double)c + ra[k + 1] * (double)rc[k + 1];
          if (cost < best) { best = cost; best_bin = k; best_axis = ax; }
        }
      }
      if (best_bin >= 0) {
        float lo = cmn[best_axis], scale = NB / (cmx[best_axis] - cmn[best_axis]);
        auto it = std::partition(idx.begin() + start, idx.begin() + end, [&](uint32_t p) {
          int k = (int)((cen[3 * (size_t)p + best_axis] - lo) * scale);
          if (k >= NB) k = NB - 1;
          if (k < 0) k = 0;
          return k <= best_bin;
        });
        mid = it - idx.begin();
        if (mid == start || mid == end) mid = -1;
      }
    }
    if (mid < 0) {  // median split on the widest centroid axis: guarantees logarithmic depth
      best_axis = axis0;
      mid = (start + end) / 2;
      std::nth_element(idx.begin() + start, idx.begin() + mid, idx.begin() + end,
                       [&](uint32_t a, uint32_t c) { return cen[3 * (size_t)a + best_axis] < cen[3 * (size_t)c + best_axis]; });
    }
    out[me].axis = best_axis;
    int64_t l = build(out, used, start, mid, depth + 1, maxd);
    int64_t r = build(out, used, mid, end, depth + 1, maxd);
    out[me].left = l;
    out[me].right = r;
    return me;
  }
};

struct Result {
  std::vector<Node32> nodes;
  std::vector<uint32_t> order;  // leaf order -> primitive index
  int depth = 0;
};

// 32-byte record of BNode i.  `a` is the node word the traversal carries and stacks (gp_trace.cuh):
//   leaf:     a = first primitive record << 3 | (nPrims - 1) << 1 | 1
//   interior: a = child-group index << 7 | axisL << 5 | axisR << 3 | axis0 << 1   (group g = records 4g .. 4g+3), filled in
//             by flatten_quads; axisL / axisR = split axes of the left / right child when that child is interior
//   unused slot of a group: a = kEmptyWord and the empty box.
// `b` keeps the leaf's primitive count (host-side checks only; the kernels do not read it).
#ifdef __CUDACC__
#define GPBVH_HD __host__ __device__
#else
#define GPBVH_HD
#endif
constexpr int kMaxLeafPrims = 4;
constexpr uint32_t kEmptyWord = 0xffffffffu;
constexpr uint64_t kMaxLeafFirst = (1ull << 29) - 1, kMaxGroups = (1ull << 25) - 1;
GPBVH_HD static inline uint32_t leaf_word(uint64_t first, int n) { return (uint32_t)(first << 3) | (uint32_t)((n - 1) << 1) | 1u; }
GPBVH_HD static inline uint32_t inner_word(uint64_t group, int axis0, int axisL, int axisR) {
  return (uint32_t)(group << 7) | (uint32_t)(axisL << 5) | (uint32_t)(axisR << 3) | (uint32_t)(axis0 << 1);
}
static Node32 make_record(const BNode& n) {
  Node32 o;
  for (int k = 0; k < 3; k++) { o.mn[k] = round_down(n.b.mn[k]); o.mx[k] = round_up(n.b.mx[k]); }
  o.a = n.n > 0 ? leaf_word((uint64_t)n.first, n.n) : 0u;
  o.b = (uint32_t)n.n;
  return o;
}
static Node32 empty_record() {
  Node32 o;
  for (int k = 0; k < 3; k++) { o.mn[k] = INFINITY; o.mx[k] = -INFINITY; }
  o.a = kEmptyWord; o.b = 0;
  return o;
}
// Child-group layout ("quads"): the binary tree is kept as built, but an interior node's record points at a group of
// FOUR adjacent 32-byte records (128 contiguous, 128-byte aligned bytes): slots 0,1 = the children of its left child
// (or slot 0 = the left child itself when that is a leaf, slot 1 empty), slots 2,3 likewise for the right child.
// One fetch therefore brings two tree levels, and a traversal step needs one dependent memory round trip per TWO
// levels.  The three split axes ride in the parent's record, so the traversal visits the (up to four) children in
// exactly the binary tree's near-first order.  The skipped middle boxes only ever culled what their children's boxes
// cull too (a child box lies inside its parent's), so the set and order of visited leaves is unchanged.
// Subtrees are laid out depth-first for locality.
static void flatten_quads(const BNode* in, int64_t i, size_t slot, std::vector<Node32>& out) {
  const BNode& n = in[i];
  Node32 rec = make_record(n);
  if (n.n == 0) {
    size_t g = out.size();
    for (int k = 0; k < 4; k++) out.push_back(empty_record());
    const BNode& L = in[n.left];
    const BNode& R = in[n.right];
    uint32_t le = L.n == 0, re = R.n == 0;
    rec.a = inner_word(g / 4, n.axis, le ? L.axis : 0, re ? R.axis : 0);
    out[slot] = rec;
    if (le) { flatten_quads(in, L.left, g, out); flatten_quads(in, L.right, g + 1, out); }
    else flatten_quads(in, n.left, g, out);
    if (re) { flatten_quads(in, R.left, g + 2, out); flatten_quads(in, R.right, g + 3, out); }
    else flatten_quads(in, n.right, g + 2, out);
  } else {
    out[slot] = rec;
  }
}
// records flatten_quads appends for the subtree entered at node i (four per interior node reached at a slot)
static size_t quad_records(const BNode* in, int64_t i) {
  const BNode& n = in[i];
  if (n.n != 0) return 0;
  const BNode& L = in[n.left];
  const BNode& R = in[n.right];
  size_t c = 4;
  if (L.n == 0) c += quad_records(in, L.left) + quad_records(in, L.right);
  if (R.n == 0) c += quad_records(in, R.left) + quad_records(in, R.right);
  return c;
}
// flatten_quads into a pre-sized array: `cursor` is where this subtree's next group goes.  Same layout, record for
// record, as the appending version; subtrees below `fork_depth` slot levels are deferred to `tasks` (node, slot, first
// group index) so that they can be written concurrently.
struct QuadTask { int64_t node; size_t slot, cursor; };
static void flatten_quads_at(const BNode* in, int64_t i, size_t slot, Node32* out, size_t& cursor, int fork_depth,
                             std::vector<QuadTask>* tasks) {
  const BNode& n = in[i];
  if (n.n != 0) { out[slot] = make_record(n); return; }
  if (tasks && fork_depth == 0) {
    tasks->push_back({i, slot, cursor});
    cursor += quad_records(in, i);
    return;
  }
  Node32 rec = make_record(n);
  size_t g = cursor;
  cursor += 4;
  for (int k = 0; k < 4; k++) out[g + k] = empty_record();
  const BNode& L = in[n.left];
  const BNode& R = in[n.right];
  uint32_t le = L.n == 0, re = R.n == 0;
  rec.a = inner_word(g / 4, n.axis, le ? L.axis : 0, re ? R.axis : 0);
  out[slot] = rec;
  int fd = fork_depth > 0 ? fork_depth - 1 : 0;
  if (le) { flatten_quads_at(in, L.left, g, out, cursor, fd, tasks); flatten_quads_at(in, L.right, g + 1, out, cursor, fd, tasks); }
  else flatten_quads_at(in, n.left, g, out, cursor, fd, tasks);
  if (re) { flatten_quads_at(in, R.left, g + 2, out, cursor, fd, tasks); flatten_quads_at(in, R.right, g + 3, out, cursor, fd, tasks); }
  else flatten_quads_at(in, n.right, g + 2, out, cursor, fd, tasks);
}
static void flatten(const BNode* in, size_t n_nodes, int64_t root, std::vector<Node32>& out) {
  size_t total = 4 + quad_records(in, root);
  out.assign(total, empty_record());  // [0] root record, [1..3] padding: child groups start on 128-byte boundaries
  size_t cursor = 4;
  unsigned hw = std::thread::hardware_concurrency();
  if (n_nodes < 200000 || hw < 2) {
    flatten_quads_at(in, root, 0, out.data(), cursor, 0, nullptr);
    return;
  }
  std::vector<QuadTask> tasks;
  flatten_quads_at(in, root, 0, out.data(), cursor, 4, &tasks);  // top 4 slot levels (<= 256 subtrees) serially
  std::vector<std::thread> th;
  std::atomic<size_t> next{0};
  for (unsigned t = 0; t < hw; t++)
    th.emplace_back([&]() {
      for (;;) {
        size_t k = next.fetch_add(1);
        if (k >= tasks.size()) break;
        size_t c = tasks[k].cursor;
        flatten_quads_at(in, tasks[k].node, tasks[k].slot, out.data(), c, 0, nullptr);
      }
    });
  for (auto& t : th) t.join();
}

// Top of the tree serially until there are enough independent subtrees, then one host thread per subtree.
static Result build_bvh(const Box* bounds, int64_t n, int max_prims) {
  Result res;
  if (n == 0) return res;
  GPBVH_TICK_INIT;
  Builder B;
  B.bounds = bounds;
  B.max_prims = std::max(1, std::min(max_prims, kMaxLeafPrims));
  B.cen.resize(3 * (size_t)n);
  B.idx.resize(n);
  for (int64_t i = 0; i < n; i++) {
    B.idx[i] = (uint32_t)i;
    for (int k = 0; k < 3; k++) B.cen[3 * (size_t)i + k] = (float)(0.5 * bounds[i].mn[k] + 0.5 * bounds[i].mx[k]);
  }
  GPBVH_TICK("centroids");
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 4;
  // One uninitialised node pool for the whole tree: a range of k primitives never needs more than 2k - 1 nodes, so the
  // slab built by a thread owns pool[2 * first primitive ...) and writes absolute indices — nothing is copied or
  // re-indexed afterwards.  The pool has gaps (only the pages that are written get touched); the flattened output
  // follows the links, so its layout does not depend on where the nodes sit.
  struct Pool {
    BNode* p;
    explicit Pool(size_t count) : p((BNode*)malloc(count * sizeof(BNode))) {}
    ~Pool() { free(p); }
  } pool(2 * (size_t)n + 512);
  if (!pool.p) return res;
  BNode* nodes = pool.p;
  int maxd = 0;
  int64_t root = 0;
  size_t n_nodes = 0;
  if (n < 200000 || hw < 2) {
    int64_t used = 0;
    root = B.build(nodes, &used, 0, n, 0, &maxd);
    n_nodes = (size_t)used;
  } else {
    // split the index range into 2^k slabs by repeated median on the widest axis, build each slab on its own thread,
    // then join the slab roots under a small top tree
    int levels = 0;
    while ((1u << levels) < hw * 2 && levels < 7) levels++;
    struct Range { int64_t s, e; int axis; };
    std::vector<std::vector<Range>> lv(levels + 1);
    lv[0].push_back({0, n, 0});
    for (int l = 0; l < levels; l++) {
      lv[l + 1].resize(lv[l].size() * 2);
      std::vector<std::thread> th;
      for (size_t r = 0; r < lv[l].size(); r++) {
        th.emplace_back([&, r, l]() {
          Range& R = lv[l][r];
          float cmn[3] = {INFINITY, INFINITY, INFINITY}, cmx[3] = {-INFINITY, -INFINITY, -INFINITY};
          for (int64_t i = R.s; i < R.e; i++)
            for (int k = 0; k < 3; k++) { float c = B.cen[3 * (size_t)B.idx[i] + k]; if (c < cmn[k]) cmn[k] = c; if (c > cmx[k]) cmx[k] = c; }
          float ex = cmx[0] - cmn[0], ey = cmx[1] - cmn[1], ez = cmx[2] - cmn[2];
          int ax = (ex > ey && ex > ez) ? 0 : (ey > ez ? 1 : 2);
          R.axis = ax;
          int64_t mid = (R.s + R.e) / 2;
          std::nth_element(B.idx.begin() + R.s, B.idx.begin() + mid, B.idx.begin() + R.e,
                           [&](uint32_t a, uint32_t c) { return B.cen[3 * (size_t)a + ax] < B.cen[3 * (size_t)c + ax]; });
          lv[l + 1][2 * r] = {R.s, mid, 0};
          lv[l + 1][2 * r + 1] = {mid, R.e, 0};
        });
      }
      for (auto& t : th) t.join();
    }
    GPBVH_TICK("median top levels");
    size_t ns = lv[levels].size();
    std::vector<int64_t> sub_root(ns), sub_used(ns);
    std::vector<int> subd(ns, 0);
    {
      std::vector<std::thread> th;
      std::atomic<size_t> next{0};
      for (unsigned t = 0; t < hw; t++)
        th.emplace_back([&]() {
          for (;;) {
            size_t r = next.fetch_add(1);
            if (r >= ns) break;
            int64_t used = 2 * lv[levels][r].s;  // this slab's region of the pool
            int64_t first = used;
            sub_root[r] = B.build(nodes, &used, lv[levels][r].s, lv[levels][r].e, levels, &subd[r]);
            sub_used[r] = used - first;
          }
        });
      for (auto& t : th) t.join();
    }
    GPBVH_TICK("subtree builds");
    for (size_t r = 0; r < ns; r++) { n_nodes += (size_t)sub_used[r]; if (subd[r] > maxd) maxd = subd[r]; }
    // the top tree over the slab roots lives behind the last slab's region
    int64_t top = 2 * n;
    std::vector<int64_t> cur = sub_root;
    for (int l = levels - 1; l >= 0; l--) {
      std::vector<int64_t> up(lv[l].size());
      for (size_t r = 0; r < lv[l].size(); r++) {
        BNode bn;
        bn.left = cur[2 * r];
        bn.right = cur[2 * r + 1];
        bn.axis = lv[l][r].axis;
        bn.b = nodes[bn.left].b;
        box_add(bn.b, nodes[bn.right].b);
        up[r] = top;
        new (&nodes[top++]) BNode(bn);
        n_nodes++;
      }
      cur = up;
    }
    root = cur[0];
    GPBVH_TICK("top tree");
  }
  if ((uint64_t)n > kMaxLeafFirst) return res;  // the node word holds 29 bits of primitive index (536 M primitives)
  flatten(nodes, n_nodes, root, res.nodes);
  GPBVH_TICK("flatten");
  if (res.nodes.size() / 4 > kMaxGroups) { res.nodes.clear(); return res; }  // ... and 25 bits of child-group index
  res.order = B.idx;
  res.depth = maxd;
  return res;
}

}  // namespace gpbvh
