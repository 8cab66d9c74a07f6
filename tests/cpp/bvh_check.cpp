// Host-side check of the BVH builder / flattener (go-pbrt_b200/csrc/gp_bvh.h) — no CUDA needed.
// Walks the flattened child-group layout exactly as the traversal kernels read it and verifies:
//   every primitive sits in exactly one leaf, leaves hold <= maxPrims primitives, every record's float32 box contains
//   the float64 bounds of everything below it (outward rounding), child groups start on 128-byte boundaries and stay
//   inside the array, the recorded depth bounds the real one, and two builds of the same input are identical.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include "../../go-pbrt_b200/csrc/gp_bvh.h"

using namespace gpbvh;

struct Walk {
  const std::vector<Node32>& nodes;
  const std::vector<uint32_t>& order;
  const std::vector<Box>& b;
  int max_prims;
  std::vector<int> seen;
  long long errors = 0;
  int max_depth = 0;
  // returns the exact float64 bound of the subtree and checks the record's float32 box against it
  Box visit(size_t rec, int depth) {
    const Node32& n = nodes[rec];
    if (depth > max_depth) max_depth = depth;
    Box acc;
    box_init(acc);
    const bool leaf = (n.a & 1u) != 0 && n.a != kEmptyWord;
    if (leaf) {
      unsigned np = ((n.a >> 1) & 3u) + 1u, first = n.a >> 3;
      if ((int)np > max_prims || np != n.b) errors++;  // a leaf never exceeds maxPrims (coincident centroids are halved by index)
      for (unsigned k = 0; k < np; k++) {
        size_t r = (size_t)first + k;
        if (r >= order.size()) { errors++; continue; }
        seen[order[r]]++;
        box_add(acc, b[order[r]]);
      }
    } else {
      if (n.a == kEmptyWord) { errors++; return acc; }
      size_t g = 4 * (size_t)(n.a >> 7);
      if (g + 3 >= nodes.size() || g == 0) { errors++; return acc; }
      for (int k = 0; k < 4; k++) {
        const Node32& c = nodes[g + k];
        if (c.a != kEmptyWord) box_add(acc, visit(g + k, depth + 1));
        else if (!(c.mn[0] > c.mx[0]) || (k != 1 && k != 3)) errors++;  // only slots 1 and 3 may be unused, and they hold the empty box
      }
      for (int k = 0; k < 3; k++) if (((n.a >> (1 + 2 * k)) & 3u) > 2u) errors++;  // split axes
    }
    for (int k = 0; k < 3; k++)
      if (!((double)n.mn[k] <= acc.mn[k] && (double)n.mx[k] >= acc.mx[k])) errors++;
    return acc;
  }
};

static int check(size_t n, int max_prims, unsigned seed, bool clustered) {
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> U(-50, 50), R(0.01, 0.8);
  std::vector<Box> b(n);
  for (size_t i = 0; i < n; i++) {
    double c[3] = {U(rng), U(rng), clustered ? 0.0 : U(rng)};
    if (clustered && i % 3 == 0) { c[0] = 1.5; c[1] = -2.25; }  // many coincident centroids
    double r = R(rng);
    for (int k = 0; k < 3; k++) { b[i].mn[k] = c[k] - r; b[i].mx[k] = c[k] + r; }
  }
  Result r1 = build_bvh(b.data(), (int64_t)n, max_prims);
  Result r2 = build_bvh(b.data(), (int64_t)n, max_prims);
  int bad = 0;
  if (r1.nodes.size() != r2.nodes.size() || memcmp(r1.nodes.data(), r2.nodes.data(), r1.nodes.size() * sizeof(Node32)) != 0 || r1.order != r2.order) {
    printf("FAIL n=%zu: two builds differ\n", n); bad++;
  }
  if (n == 0) return bad + (r1.nodes.empty() ? 0 : 1);
  if (r1.nodes.size() % 4 != 0) { printf("FAIL n=%zu: node array not a whole number of groups\n", n); bad++; }
  Walk w{r1.nodes, r1.order, b, max_prims, std::vector<int>(n, 0)};
  w.visit(0, 0);
  long long missing = 0;
  for (size_t i = 0; i < n; i++) if (w.seen[i] != 1) missing++;
  // slot depth d corresponds to binary depth <= 2 d; the builder reports the binary depth
  if (w.errors || missing || w.max_depth * 2 > r1.depth + 2 + 1 || r1.depth > 2 * w.max_depth + 1) {
    printf("FAIL n=%zu maxPrims=%d: errors=%lld missing=%lld slot_depth=%d reported_depth=%d\n", n, max_prims, w.errors, missing, w.max_depth, r1.depth);
    bad++;
  }
  return bad;
}

int main() {
  int bad = 0;
  size_t sizes[] = {0, 1, 2, 3, 5, 17, 100, 1000, 30000, 250000};  // the last one takes the threaded path
  for (size_t n : sizes)
    for (int mp : {1, 2, 4}) {  // (maxPrims above kMaxLeafPrims = 4 is clamped by the builder)
      bad += check(n, mp, 7 + (unsigned)n, false);
      if (n >= 17 && n <= 30000) bad += check(n, mp, 11 + (unsigned)n, true);
    }
  printf(bad ? "bvh_check: %d FAILURES\n" : "bvh_check: ok\n", bad);
  return bad ? 1 : 0;
}
