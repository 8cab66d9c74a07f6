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
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace gpbvh {

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

  // builds [start,end) into `out`, returns the root index within `out`
  int64_t build(std::vector<BNode>& out, int64_t start, int64_t end, int depth, int* maxd) {
    int64_t me = (int64_t)out.size();
    out.emplace_back();
    if (depth > *maxd) *maxd = depth;
    Box b, cb;
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
    if (n <= max_prims || !(cmx[axis0] > cmn[axis0])) {
      if (n <= 255 * 1024) {  // leaf (a degenerate cluster of coincident centroids stays one leaf; count field is 24 bits)
        out[me].first = start;
        out[me].n = (int)n;
        out[me].axis = axis0;
        return me;
      }
    }
    int64_t mid = -1;
    int best_axis = axis0;
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
    int64_t l = build(out, start, mid, depth + 1, maxd);
    int64_t r = build(out, mid, end, depth + 1, maxd);
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

// 32-byte record of BNode i.  leaf: {a = first primitive record, b = nPrims<<8 | axis}.
// interior: {a = index of its 4-record CHILD GROUP, b = axis0 | axis1<<2 | axis2<<4 | L_expanded<<6 | R_expanded<<7}
// (b>>8 == 0 marks an interior record), filled in by flatten_quads.
static Node32 make_record(const BNode& n) {
  Node32 o;
  for (int k = 0; k < 3; k++) { o.mn[k] = round_down(n.b.mn[k]); o.mx[k] = round_up(n.b.mx[k]); }
  o.a = (uint32_t)n.first;
  o.b = n.n > 0 ? (((uint32_t)n.n << 8) | (uint32_t)n.axis) : (uint32_t)n.axis;
  return o;
}
static Node32 empty_record() {
  Node32 o;
  for (int k = 0; k < 3; k++) { o.mn[k] = INFINITY; o.mx[k] = -INFINITY; }
  o.a = 0; o.b = 0;
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
static void flatten_quads(const std::vector<BNode>& in, int64_t i, size_t slot, std::vector<Node32>& out) {
  const BNode& n = in[i];
  Node32 rec = make_record(n);
  if (n.n == 0) {
    size_t g = out.size();
    for (int k = 0; k < 4; k++) out.push_back(empty_record());
    const BNode& L = in[n.left];
    const BNode& R = in[n.right];
    uint32_t le = L.n == 0, re = R.n == 0;
    rec.a = (uint32_t)g;
    rec.b = (uint32_t)n.axis | ((uint32_t)(le ? L.axis : 0) << 2) | ((uint32_t)(re ? R.axis : 0) << 4) | (le << 6) | (re << 7);
    out[slot] = rec;
    if (le) { flatten_quads(in, L.left, g, out); flatten_quads(in, L.right, g + 1, out); }
    else flatten_quads(in, n.left, g, out);
    if (re) { flatten_quads(in, R.left, g + 2, out); flatten_quads(in, R.right, g + 3, out); }
    else flatten_quads(in, n.right, g + 2, out);
  } else {
    out[slot] = rec;
  }
}
static void flatten(const std::vector<BNode>& in, int64_t root, std::vector<Node32>& out) {
  out.push_back(empty_record());  // [0] root record
  for (int k = 0; k < 3; k++) out.push_back(empty_record());  // [1..3] padding: child groups start on 128-byte boundaries
  flatten_quads(in, root, 0, out);
}

// Top of the tree serially until there are enough independent subtrees, then one host thread per subtree.
static Result build_bvh(const Box* bounds, int64_t n, int max_prims) {
  Result res;
  if (n == 0) return res;
  Builder B;
  B.bounds = bounds;
  B.max_prims = max_prims;
  B.cen.resize(3 * (size_t)n);
  B.idx.resize(n);
  for (int64_t i = 0; i < n; i++) {
    B.idx[i] = (uint32_t)i;
    for (int k = 0; k < 3; k++) B.cen[3 * (size_t)i + k] = (float)(0.5 * bounds[i].mn[k] + 0.5 * bounds[i].mx[k]);
  }
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 4;
  std::vector<BNode> nodes;
  int maxd = 0;
  if (n < 200000 || hw < 2) {
    nodes.reserve(2 * (size_t)n / std::max(1, max_prims) + 16);
    B.build(nodes, 0, n, 0, &maxd);
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
    size_t ns = lv[levels].size();
    std::vector<std::vector<BNode>> sub(ns);
    std::vector<int> subd(ns, 0);
    {
      std::vector<std::thread> th;
      std::atomic<size_t> next{0};
      for (unsigned t = 0; t < hw; t++)
        th.emplace_back([&]() {
          for (;;) {
            size_t r = next.fetch_add(1);
            if (r >= ns) break;
            sub[r].reserve(2 * (size_t)(lv[levels][r].e - lv[levels][r].s) / std::max(1, max_prims) + 16);
            B.build(sub[r], lv[levels][r].s, lv[levels][r].e, levels, &subd[r]);
          }
        });
      for (auto& t : th) t.join();
    }
    // stitch: top tree nodes first (recursively), subtrees appended with index fix-up
    std::vector<int64_t> sub_root(ns);
    for (size_t r = 0; r < ns; r++) {
      int64_t base = (int64_t)nodes.size();
      sub_root[r] = base;
      for (auto bn : sub[r]) {
        if (bn.left >= 0) { bn.left += base; bn.right += base; }
        nodes.push_back(bn);
      }
      std::vector<BNode>().swap(sub[r]);
      if (subd[r] > maxd) maxd = subd[r];
    }
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
        up[r] = (int64_t)nodes.size();
        nodes.push_back(bn);
      }
      cur = up;
    }
    // move the root to a known place: flatten() takes the root index
    res.nodes.reserve(nodes.size());
    flatten(nodes, cur[0], res.nodes);
    res.order = B.idx;
    res.depth = maxd;
    return res;
  }
  res.nodes.reserve(nodes.size());
  flatten(nodes, 0, res.nodes);
  res.order = B.idx;
  res.depth = maxd;
  return res;
}

}  // namespace gpbvh
