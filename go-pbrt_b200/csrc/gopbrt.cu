// gopbrt.cu — C ABI of libgopbrt_cuda.so (include/gopbrt_cuda.h): scene upload, batched Intersect/IntersectP,
// and the wavefront render loop.  sm_100a only; there is no CPU fallback — every entry point needs a live B200.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off -shared
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gopbrt_cuda.h"
#include "gp_bvh.h"
#include "gp_comm.h"
#include "gp_render.cuh"
#include "gp_build.cuh"
#include "gp_kat.cuh"

using namespace gp;

// the ctypes / cgo views of these structs are checked against the same numbers (tests/test_abi.py)
static_assert(sizeof(gopbrt_transform) == 256 && sizeof(gopbrt_sphere) == 40 && sizeof(gopbrt_disk) == 40, "ABI layout");
static_assert(sizeof(gopbrt_triangle) == 16 && sizeof(gopbrt_primitive) == 16 && sizeof(gopbrt_material) == 48, "ABI layout");
static_assert(sizeof(gopbrt_texture) == 136 && sizeof(gopbrt_light) == 64 && sizeof(gopbrt_camera) == 288, "ABI layout");
static_assert(sizeof(gopbrt_sampler) == 24 && sizeof(gopbrt_integrator) == 32 && sizeof(gopbrt_film) == 56, "ABI layout");
static_assert(sizeof(gopbrt_render_options) == 16 && sizeof(gopbrt_stats) == 280 && sizeof(gopbrt_scene_desc) == 152, "ABI layout");

struct gopbrt_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::string last_error;
  std::atomic<uint64_t> launches{0};
  std::mutex mu;
  // Every scene handle of a device context submits to ONE stream (and the render loop captures CUDA graphs on it), so
  // device work of concurrent calls — each gRPC request of the reference renders on its own goroutine with its own scene,
  // SURVEY §8b — is serialised per context; gopbrt_cancel stays lock-free.
  std::mutex run_mu;
  int grid_gen[8] = {0, 0, 0, 0, 0, 0, 0, 0}, grid_shade[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // persistent grids of this device (SMs x resident CTAs)
  size_t trace_smem_limit = 0;  // largest dynamic shared memory size the traversal kernels have been opted into (> 48 KB only)
  // film reduce across GPUs (gopbrt_comm_init_rank / gopbrt_multi_init): this context's NCCL communicator and its place in it
  gpcomm::ncclComm_t comm = nullptr;
  int comm_rank = 0, comm_world = 1;
};

#define GP_CUDA(ctx, call)                                                                               \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess) {                                                                            \
      (ctx)->last_error = std::string(#call) + ": " + cudaGetErrorString(e__);                           \
      return GOPBRT_ERR_CUDA;                                                                            \
    }                                                                                                    \
  } while (0)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    return cudaMalloc((void**)&p, count * sizeof(T));
  }
  cudaError_t upload(const std::vector<T>& v, cudaStream_t s) {
    cudaError_t e = alloc(v.size());
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
  }
  // grow-only: keeps the allocation when it is already large enough (n then stays the capacity)
  cudaError_t ensure(size_t count) { return count <= n ? cudaSuccess : alloc(count + count / 4); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  ~DevBuf() { release(); }
};

// Scratch of the batched Intersect / IntersectP entry points, kept per scene and only ever grown: a host that traces batch
// after batch (the parity tests of the Go side, a baking tool) pays cudaMalloc / cudaFree and page-locking once, not per call.
struct BatchPool {
  DevBuf<RayRec> recs;
  DevBuf<double> rays, t, p, n;
  DevBuf<int> rec, prim;
  DevBuf<unsigned char> hit;
  double* stage = nullptr;  // pinned host staging: the caller's seven SoA arrays are gathered here and go up as ONE copy
  size_t stage_cap = 0;
  cudaError_t ensure_stage(size_t doubles) {
    if (doubles <= stage_cap) return cudaSuccess;
    if (stage) cudaFreeHost(stage);
    stage = nullptr; stage_cap = 0;
    cudaError_t e = cudaHostAlloc((void**)&stage, (doubles + doubles / 4) * sizeof(double), cudaHostAllocDefault);
    if (e == cudaSuccess) stage_cap = doubles + doubles / 4;
    return e;
  }
  ~BatchPool() { if (stage) cudaFreeHost(stage); }
};

struct Workspace {  // per-scene render workspace, kept between gopbrt_render calls of the same shape
  long long lanes = 0;
  long long cap = 0, cap_key[5] = {0, 0, 0, 0, 0};  // lane capacity granted to the last request, and what that request was
  size_t per_lane = 0;
  int auto_groups = 1;  // the lane-group count the automatic rule chose for that request
  size_t bytes_tables = 0, bytes_tilepix = 0;
  DevBuf<RayRec> ray;
  DevBuf<ShadowRec> sray;
  DevBuf<PathRec> path;
  DevBuf<RadRec> rad;
  DevBuf<FilmRec> fsum;
  DevBuf<int> i32;        // queues
  DevBuf<double> tables, tilepix, frames;
  DevBuf<unsigned char> occl;
  DevBuf<unsigned char> codes;  // per extend-queue position: the hit's shade class or 4 = escaped (extend -> split)
  DevBuf<int> cnt;
  DevBuf<RenderCounters> rctr;
  int* remaining_host = nullptr;  // pinned, device-mapped
  int* remaining_dev = nullptr;
  std::vector<cudaEvent_t> events;  // pool for GOPBRT_FLAG_TIME_KERNELS
  DevBuf<double> film;              // device film of gopbrt_render (host-film entry point), kept between calls
  // CUDA graph of kGraphIters wavefront iterations, re-used while the launch arguments stay the same
  cudaGraphExec_t graph_exec = nullptr;
  std::vector<unsigned char> graph_key;
  cudaEvent_t graph_ev[2] = {nullptr, nullptr};
  cudaEvent_t frame_ev[2] = {nullptr, nullptr};
  ~Workspace() {
    for (auto e : frame_ev) if (e) cudaEventDestroy(e);
    if (remaining_host) cudaFreeHost(remaining_host);
    for (auto e : events) cudaEventDestroy(e);
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    for (auto e : graph_ev) if (e) cudaEventDestroy(e);
  }
};

typedef void (*trace_fn)(DevScene, RayRec*, const ShadowRec*, RadRec*, unsigned char*, const int*, const int*, long long, int, int*,
                         TraceCounters*, unsigned long long*);

struct gopbrt_scene {
  gopbrt_ctx* ctx = nullptr;
  DevScene dev{};
  DevBuf<gpbvh::Node32> nodes;
  DevBuf<gpbvh::Node32> flat;  // small scenes: the flat aggregate's table (k_trace_flat)
  DevBuf<PrimRec> recs;
  DevBuf<double> rec_bounds;
  DevBuf<int4> prims;
  DevBuf<double> xf;
  DevBuf<int> xf_flags;
  DevBuf<SphereDev> spheres;
  DevBuf<DiskDev> disks;
  DevBuf<MaterialDev> materials;
  DevBuf<TextureDev> textures;
  DevBuf<LightDev> lights;
  DevBuf<double> light_cdf;
  DevBuf<TraceCounters> tctr;
  DevBuf<int> work;  // work counters of the persistent traversal warps (batched API)
  int stack_cap = 8;
  unsigned class_mask = 0;  // shade classes (RF_CLASS_*) that occur among the scene's primitives
  // traversal kernels of this scene: [0] extend, [1] extend + counters, [2] shadow, [3] shadow + counters, [4] batched any-hit,
  // [5] per-segment shadow (DirectLighting / UniformSampleAll), [6] the same + counters
  trace_fn trace_k[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int trace_grid[7] = {0, 0, 0, 0, 0, 0, 0};
  size_t trace_smem = 0;
  double world[6] = {0, 0, 0, 0, 0, 0};
  uint64_t bvh_nodes = 0, bvh_depth = 0;
  std::atomic<int> cancel{0};
  Workspace ws;
  BatchPool pool;
  std::mutex mu;
};

static int grid_for(gopbrt_ctx* ctx, const void* kernel, int block, size_t smem = 0) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  return ctx->sm_count * per_sm;  // persistent grid: a whole number of resident CTAs per SM
}

extern "C" {

int gopbrt_abi_version(void) { return GOPBRT_ABI_VERSION; }

int gopbrt_init(int device, gopbrt_ctx** out) {
  if (!out) return GOPBRT_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return GOPBRT_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GOPBRT_ERR_CUDA;
  if (prop.major != 10) return GOPBRT_ERR_CUDA;  // sm_100a cubin only: no fallback
  if (cudaSetDevice(device) != cudaSuccess) return GOPBRT_ERR_CUDA;
  gopbrt_ctx* c = new gopbrt_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return GOPBRT_ERR_CUDA; }
  *out = c;
  return GOPBRT_OK;
}

void gopbrt_shutdown(gopbrt_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->comm) {
    std::string e;
    if (gpcomm::Api* a = gpcomm::api(e)) a->CommDestroy(c->comm);
  }
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* gopbrt_last_error(const gopbrt_ctx* c) { return c ? c->last_error.c_str() : "null ctx"; }
uint64_t gopbrt_launch_count(const gopbrt_ctx* c) { return c ? c->launches.load() : 0; }

}  // extern "C"

// ------------------------------------------------------------------------------------------------ scene build (host)
static M4 m4_from(const double* a) {
  M4 m;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m.m[i][j] = a[i * 4 + j];
  return m;
}
static bool is_translation_only(const double* a) {
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
    if (j == 3 && i < 3) continue;
    if (a[i * 4 + j] != (i == j ? 1.0 : 0.0)) return false;
  }
  return true;
}
static bool is_identity(const double* a) {
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) if (a[i * 4 + j] != (i == j ? 1.0 : 0.0)) return false;
  return true;
}

// Bounds3 with the reference's Union/UnionPoint semantics (bounds.go:209-238: Go Min/Max)
struct HB { double mn[3], mx[3]; bool valid = false; };
static void hb_union_point(HB& b, V3 p) {
  double v[3] = {p.x, p.y, p.z};
  if (!b.valid) { for (int k = 0; k < 3; k++) { b.mn[k] = v[k]; b.mx[k] = v[k]; } b.valid = true; }
  for (int k = 0; k < 3; k++) { b.mn[k] = go_min(b.mn[k], v[k]); b.mx[k] = go_max(b.mx[k], v[k]); }
}
static void hb_union(HB& b, const HB& o) {
  if (!o.valid) return;
  if (!b.valid) b = o;
  for (int k = 0; k < 3; k++) { b.mn[k] = go_min(b.mn[k], o.mn[k]); b.mx[k] = go_max(b.mx[k], o.mx[k]); }
}
// Transform.TransformBounds (transform.go:336-345)
static HB xf_bounds(const M4& m, const HB& b) {
  HB r;
  hb_union_point(r, xf_point(m, mk3(b.mn[0], b.mn[1], b.mn[2]), mk3(0, 0, 0), nullptr));
  for (int i = 1; i < 8; i++) {
    V3 c = mk3((i & 1) ? b.mx[0] : b.mn[0], (i & 2) ? b.mx[1] : b.mn[1], (i & 4) ? b.mx[2] : b.mn[2]);  // bounds.go:114-120
    hb_union_point(r, xf_point(m, c, mk3(0, 0, 0), nullptr));
  }
  return r;
}

// NewUniformLightDistribution + NewDistribution1D (lightdistribution.go:25-34, sampling.go:11-40): func[i] = 1; returns FuncInt
static double uniform_light_cdf(int nl, std::vector<double>& cdf) {
  cdf.assign(nl + 1, 0.0);
  for (int i = 1; i < nl + 1; i++) cdf[i] = cdf[i - 1] + 1.0 / (double)nl;
  double func_int = nl ? cdf[nl] : 0.0;
  if (func_int == 0.0) { for (int i = 1; i < nl + 1; i++) cdf[i] = (double)i / (double)nl; }
  else { for (int i = 1; i < nl + 1; i++) cdf[i] /= func_int; }
  return func_int;
}

// fn(begin, end, chunk index) over [0, n) on the host's cores (one chunk per thread, in index order)
template <class F>
static void parallel_chunks(int64_t n, F fn, int* n_chunks_out = nullptr) {
  unsigned hw = std::thread::hardware_concurrency();
  int nt = (n < 100000 || hw < 2) ? 1 : (int)std::min<unsigned>(hw, 64);
  if (n_chunks_out) *n_chunks_out = nt;
  if (nt == 1) { fn((int64_t)0, n, 0); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; t++) th.emplace_back([=]() { fn(n * t / nt, n * (t + 1) / nt, t); });
  for (auto& t : th) t.join();
}

// Everything gopbrt_scene_create derives on the host: validated tables, the primitives' reference-arithmetic world bounds, the
// BVH, the leaf-ordered records.  Built once per scene; uploaded once per device (gopbrt_multi_scene_create replicates it).
struct HostScene {
  std::vector<double> xf;
  std::vector<int> xf_flags;
  std::vector<SphereDev> spheres;
  std::vector<DiskDev> disks;
  std::vector<int4> prims;
  HB world;
  gpbvh::Result bvh;
  std::vector<PrimRec> recs;
  std::vector<double> rec_bounds;
  std::vector<gpbvh::Node32> flat;  // the flat aggregate's entries, then its box groups (n_flat_groups records)
  unsigned long long flat_tri_mask = 0;
  int n_flat_entries = 0, n_flat_groups = 0;
  std::vector<MaterialDev> mats;
  std::vector<TextureDev> texs;
  std::vector<LightDev> lights;
  std::vector<double> cdf;
  double func_int = 0;
  int nl = 0;
  // Device build (gp_build.cuh): the BVH, the child groups and the leaf-ordered records are made on the GPU from the caller's
  // own arrays (valid for the duration of the create call).  The host only prepares what needs the reference's arithmetic:
  // bounds and records of spheres / disks (compact lists addressed through qslot), and the per-material shade class.
  bool device_build = false;
  const gopbrt_scene_desc* desc = nullptr;
  std::vector<int> qslot;
  std::vector<PrimRec> qrecs;
  std::vector<double> qbounds;
  std::vector<unsigned char> mat_not_lambert;
  int max_prims = 2;
};

static int build_host_scene(const gopbrt_scene_desc* d, HostScene& H, std::string& error) {
  auto bad = [&](const char* msg) { error = msg; return GOPBRT_ERR_INVALID; };
  std::vector<double>& xf = H.xf;
  std::vector<int>& xf_flags = H.xf_flags;
  std::vector<SphereDev>& spheres = H.spheres;
  std::vector<DiskDev>& disks = H.disks;
  std::vector<int4>& prims = H.prims;
  HB& world = H.world;
  gpbvh::Result& bvh = H.bvh;
  std::vector<PrimRec>& recs = H.recs;
  std::vector<double>& rec_bounds = H.rec_bounds;
  std::vector<gpbvh::Node32>& flat = H.flat;
  unsigned long long& flat_tri_mask = H.flat_tri_mask;
  std::vector<MaterialDev>& mats = H.mats;
  std::vector<TextureDev>& texs = H.texs;
  std::vector<LightDev>& lights = H.lights;
  std::vector<double>& cdf = H.cdf;
  if (d->n_primitives < 0 || d->n_primitives > 0x7fffffff) return bad("n_primitives out of range");
  // GOPBRT_HOST_TIMING=1: host wall-clock of the build's phases on stderr (tuning aid)
  const bool host_timing = getenv("GOPBRT_HOST_TIMING") != nullptr;
  auto hw0 = std::chrono::steady_clock::now();
  auto hw_tick = [&](const char* what) {
    if (!host_timing) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[gopbrt scene] %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - hw0).count());
    hw0 = now;
  };

  // ---- transforms
  xf.assign(32 * (size_t)d->n_transforms, 0.0);
  xf_flags.assign(d->n_transforms, 0);
  for (int i = 0; i < d->n_transforms; i++) {
    memcpy(&xf[32 * (size_t)i], d->transforms[i].m, 16 * sizeof(double));
    memcpy(&xf[32 * (size_t)i + 16], d->transforms[i].minv, 16 * sizeof(double));
    if (is_translation_only(d->transforms[i].m) && is_translation_only(d->transforms[i].minv)) xf_flags[i] |= XF_TRANSLATION;
    if (is_identity(d->transforms[i].m)) xf_flags[i] |= XF_IDENTITY;  // Transform.IsIdentity looks at Matrix only (transform.go:167-173)
  }
  auto xf_ok = [&](int i) { return i >= 0 && i < d->n_transforms; };

  // ---- shapes (NewSphere sphere.go:19-32, NewDisk disk.go:22-35)
  spheres.resize(d->n_spheres);
  for (int i = 0; i < d->n_spheres; i++) {
    const gopbrt_sphere& s = d->spheres[i];
    if (!xf_ok(s.object_to_world)) return bad("sphere transform index out of range");
    SphereDev o;
    o.radius = s.radius;
    o.zMin = go_clamp(go_min(s.z_min, s.z_max), -s.radius, s.radius);
    o.zMax = go_clamp(go_max(s.z_min, s.z_max), -s.radius, s.radius);
    o.thetaMin = go_acos(go_clamp(go_min(s.z_min, s.z_max) / s.radius, -1, 1));
    o.thetaMax = go_acos(go_clamp(go_max(s.z_min, s.z_max) / s.radius, -1, 1));
    o.phiMax = kPi / 180.0 * go_clamp(s.phi_max_deg, 0, 360);
    o.xf = s.object_to_world;
    o.flags = (s.reverse_orientation ? RF_REVERSE : 0);
    if (!(o.zMin > -o.radius) && !(o.zMax < o.radius) && o.phiMax >= 2 * kPi) o.flags |= RF_FULL;
    spheres[i] = o;
  }
  disks.resize(d->n_disks);
  for (int i = 0; i < d->n_disks; i++) {
    const gopbrt_disk& s = d->disks[i];
    if (!xf_ok(s.object_to_world)) return bad("disk transform index out of range");
    DiskDev o;
    o.height = s.height; o.radius = s.radius; o.innerRadius = s.inner_radius;
    o.phiMax = kPi / 180.0 * go_clamp(s.phi_max_deg, 0, 360);
    o.xf = s.object_to_world;
    o.flags = (s.reverse_orientation ? RF_REVERSE : 0);
    disks[i] = o;
  }

  // ---- primitives: world bounds (Primitive.WorldBound) in the reference's arithmetic
  int64_t np = d->n_primitives;
  H.desc = d;
  H.device_build = np > kFlatMax && !getenv("GOPBRT_HOST_BVH");
  // validates primitive i and returns its float64 world bound; nullptr on success, else the error text
  auto prim_bound = [&](int64_t i, HB& b) -> const char* {
    const gopbrt_primitive& p = d->primitives[i];
    if (p.material >= d->n_materials) return "material index out of range";
    if (p.prim_to_world >= d->n_transforms) return "prim_to_world index out of range";
    if (p.shape_kind == GOPBRT_SHAPE_SPHERE) {
      if (p.shape_index < 0 || p.shape_index >= d->n_spheres) return "sphere index out of range";
      const SphereDev& s = spheres[p.shape_index];
      HB ob; ob.valid = true;
      ob.mn[0] = -s.radius; ob.mn[1] = -s.radius; ob.mn[2] = s.zMin; ob.mx[0] = s.radius; ob.mx[1] = s.radius; ob.mx[2] = s.zMax;  // sphere.go:46-51
      b = xf_bounds(m4_from(&xf[32 * (size_t)s.xf]), ob);
    } else if (p.shape_kind == GOPBRT_SHAPE_DISK) {
      if (p.shape_index < 0 || p.shape_index >= d->n_disks) return "disk index out of range";
      const DiskDev& s = disks[p.shape_index];
      HB ob; ob.valid = true;
      ob.mn[0] = -s.radius; ob.mn[1] = -s.radius; ob.mn[2] = s.height; ob.mx[0] = s.radius; ob.mx[1] = s.radius; ob.mx[2] = s.height;  // disk.go:41-54
      b = xf_bounds(m4_from(&xf[32 * (size_t)s.xf]), ob);
    } else if (p.shape_kind == GOPBRT_SHAPE_TRIANGLE) {
      if (p.shape_index < 0 || p.shape_index >= d->n_triangles) return "triangle index out of range";
      const gopbrt_triangle& t = d->triangles[p.shape_index];
      for (int k = 0; k < 3; k++) {
        if (t.v[k] < 0 || t.v[k] >= d->n_vertices) return "vertex index out of range";
        hb_union_point(b, mk3(d->vertices[3 * (size_t)t.v[k]], d->vertices[3 * (size_t)t.v[k] + 1], d->vertices[3 * (size_t)t.v[k] + 2]));
      }
      if (p.prim_to_world >= 0) return "TransformedPrimitive around a triangle is not supported";
    } else {
      return "unknown shape kind";
    }
    if (p.prim_to_world >= 0) b = xf_bounds(m4_from(&xf[32 * (size_t)p.prim_to_world]), b);  // primitive.go:127-129
    return nullptr;
  };
  // the traversal record of primitive pi (leaf order is applied by the caller)
  auto make_record = [&](uint32_t pi) -> PrimRec {
    const gopbrt_primitive& p = d->primitives[pi];
    PrimRec rec;
    memset(&rec, 0, sizeof(rec));
    rec.prim = pi;
    rec.flags = (uint32_t)p.shape_kind;
    if (p.shape_kind == GOPBRT_SHAPE_TRIANGLE) {
      const gopbrt_triangle& t = d->triangles[p.shape_index];
      for (int k = 0; k < 3; k++) for (int c = 0; c < 3; c++) rec.d[3 * k + c] = d->vertices[3 * (size_t)t.v[k] + c];
      rec.flags |= RF_FAST | (t.reverse_orientation ? RF_REVERSE : 0);
    } else if (p.shape_kind == GOPBRT_SHAPE_SPHERE) {
      const SphereDev& s = spheres[p.shape_index];
      rec.flags |= (uint32_t)(s.flags & (RF_REVERSE | RF_FULL));
      bool fast = (s.flags & RF_FULL) && (xf_flags[s.xf] & XF_TRANSLATION) && (p.prim_to_world < 0 || (xf_flags[p.prim_to_world] & XF_TRANSLATION));
      if (fast) {
        rec.flags |= RF_FAST;
        rec.d[0] = s.radius;
        const double* inv = &xf[32 * (size_t)s.xf + 16];  // worldToObject.Matrix = objectToWorld.MatrixInverse
        rec.d[1] = inv[3]; rec.d[2] = inv[7]; rec.d[3] = inv[11];
        if (p.prim_to_world >= 0) {
          const double* pinv = &xf[32 * (size_t)p.prim_to_world + 16];
          rec.d[4] = pinv[3]; rec.d[5] = pinv[7]; rec.d[6] = pinv[11];
          rec.flags |= RF_HAS_P2W;
        }
      }
    } else {
      rec.flags |= (uint32_t)(disks[p.shape_index].flags & RF_REVERSE);
    }
    {  // shade class (used to bin the shade queue so that warps shade one kind of hit)
      uint32_t cls = (p.shape_kind == GOPBRT_SHAPE_TRIANGLE) ? 0u : 2u;
      bool lambert = p.material >= 0 && p.material < d->n_materials && d->materials[p.material].kind == GOPBRT_MAT_MATTE &&
                     !(go_clamp(d->materials[p.material].sigma, 0, 90) != 0);
      if (!lambert) cls |= 1u;
      rec.flags |= cls << RF_CLASS_SHIFT;
    }
    return rec;
  };
  H.max_prims = d->max_prims_in_node > 0 ? std::min(255, d->max_prims_in_node) : 4;
  H.max_prims = std::min(H.max_prims, 2);
  if (const char* mp = getenv("GOPBRT_MAX_PRIMS")) H.max_prims = std::max(1, std::min(255, atoi(mp)));
  if (H.device_build) {
    // spheres / disks only: validated, bounded and given their records here; triangles are handled on the device
    std::vector<int64_t> cnt(65, 0);
    int n_chunks = 1;
    parallel_chunks(np, [&](int64_t i0, int64_t i1, int chunk) {
      int64_t c = 0;
      for (int64_t i = i0; i < i1; i++) c += d->primitives[i].shape_kind != GOPBRT_SHAPE_TRIANGLE;
      cnt[chunk + 1] = c;
    }, &n_chunks);
    for (int c = 0; c < n_chunks; c++) cnt[c + 1] += cnt[c];
    const int64_t nq = cnt[n_chunks];
    if (nq > 0) {
      H.qslot.assign(np, -1);
      H.qrecs.resize(nq);
      H.qbounds.resize(6 * (size_t)nq);
      std::vector<const char*> err(64, nullptr);
      parallel_chunks(np, [&](int64_t i0, int64_t i1, int chunk) {
        int64_t slot = cnt[chunk];
        for (int64_t i = i0; i < i1; i++) {
          if (d->primitives[i].shape_kind == GOPBRT_SHAPE_TRIANGLE) continue;
          HB b;
          if (const char* e = prim_bound(i, b)) { err[chunk] = e; return; }
          H.qslot[i] = (int)slot;
          for (int k = 0; k < 3; k++) { H.qbounds[6 * (size_t)slot + k] = b.mn[k]; H.qbounds[6 * (size_t)slot + 3 + k] = b.mx[k]; }
          H.qrecs[slot] = make_record((uint32_t)i);
          slot++;
        }
      });
      for (int c = 0; c < n_chunks; c++) if (err[c]) return bad(err[c]);
    }
    H.mat_not_lambert.resize(std::max(1, d->n_materials));
    for (int m = 0; m < d->n_materials; m++)
      H.mat_not_lambert[m] = !(d->materials[m].kind == GOPBRT_MAT_MATTE && !(go_clamp(d->materials[m].sigma, 0, 90) != 0));
    hw_tick("sphere/disk records");
  }
  std::vector<gpbvh::Box> pb(H.device_build ? 0 : np);
  if (!H.device_build) {
    prims.resize(np);
    std::vector<HB> part(64);
    std::vector<const char*> err(64, nullptr);
    int n_chunks = 1;
    parallel_chunks(np, [&](int64_t i0, int64_t i1, int chunk) {
      HB& wpart = part[chunk];
      for (int64_t i = i0; i < i1; i++) {
        const gopbrt_primitive& p = d->primitives[i];
        prims[i] = make_int4(p.shape_kind, p.shape_index, p.material, p.prim_to_world);
        HB b;
        if (const char* e = prim_bound(i, b)) { err[chunk] = e; return; }
        for (int k = 0; k < 3; k++) { pb[i].mn[k] = b.mn[k]; pb[i].mx[k] = b.mx[k]; }
        hb_union(wpart, b);
      }
    }, &n_chunks);
    for (int c = 0; c < n_chunks; c++) {
      if (err[c]) return bad(err[c]);
      if (part[c].valid) hb_union(world, part[c]);  // chunks in index order: the same union sequence, merely bracketed
    }
  }

  hw_tick("tables + prim bounds");
  // ---- BVH
  // maxPrimsInNode (bvh.go:223-231) is an UPPER bound on a leaf; the tree is this backend's own, and every candidate of a
  // leaf costs a record fetch plus a float64 bound test per ray, so leaves hold at most two primitives (measured on
  // B200, FAST mode: config 2 frame 246 -> 225 ms, config 4 38.4 -> 36.3 ms against leaves of four; one per leaf is no
  // better).  Results do not depend on the tree.  GOPBRT_MAX_PRIMS overrides (tuning aid).
  if (!H.device_build) {
    bvh = gpbvh::build_bvh(pb.data(), np, H.max_prims);
    if ((int64_t)bvh.order.size() != np || (np > 0 && bvh.nodes.empty())) return bad("BVH build failed (out of host memory, or more primitives / nodes than a node word addresses)");
    if (bvh.depth >= kStackDepth - 1) return bad("BVH deeper than the traversal stack");
    hw_tick("bvh build + flatten");

    // ---- leaf-ordered primitive records + their float64 bounds
    recs.resize(np);
    rec_bounds.resize(6 * (size_t)np);
    parallel_chunks(np, [&](int64_t r0, int64_t r1, int) {
      for (int64_t r = r0; r < r1; r++) {
        uint32_t pi = bvh.order[r];
        recs[r] = make_record(pi);
        for (int k = 0; k < 3; k++) { rec_bounds[6 * (size_t)r + k] = pb[pi].mn[k]; rec_bounds[6 * (size_t)r + 3 + k] = pb[pi].mx[k]; }
      }
    });
  }

  hw_tick("leaf-ordered records");
  // ---- flat aggregate (scenes of at most kFlatMax primitives): one table entry per leaf-ordered record, triangles first
  flat.clear();
  flat_tri_mask = 0;
  if (np > 0 && np <= kFlatMax && !getenv("GOPBRT_NO_FLAT")) {
    for (int pass = 0; pass < 2; pass++)
      for (int64_t r = 0; r < np; r++) {
        const bool tri = (recs[r].flags & RK_KIND_MASK) == RK_TRIANGLE;
        if (tri != (pass == 0)) continue;
        gpbvh::Node32 e;
        for (int k = 0; k < 3; k++) { e.mn[k] = gpbvh::round_down(rec_bounds[6 * (size_t)r + k]); e.mx[k] = gpbvh::round_up(rec_bounds[6 * (size_t)r + 3 + k]); }
        e.a = (uint32_t)r;
        e.b = recs[r].flags;
        if (tri) flat_tri_mask |= 1ULL << flat.size();
        flat.push_back(e);
      }
    // Box groups: entries whose float32 boxes are bit-identical (the two triangles of an axis-aligned or upright quad) are
    // tested ONCE in the kernel's warp-uniform loop; a group record is {box, a | b << 32 = the mask of its entries}.
    H.n_flat_entries = (int)flat.size();
    std::vector<gpbvh::Node32> groups;
    for (int k = 0; k < H.n_flat_entries; k++) {
      size_t g = 0;
      for (; g < groups.size(); g++)
        if (memcmp(groups[g].mn, flat[k].mn, sizeof(float) * 3) == 0 && memcmp(groups[g].mx, flat[k].mx, sizeof(float) * 3) == 0) break;
      if (g == groups.size()) { gpbvh::Node32 e = flat[k]; e.a = 0; e.b = 0; groups.push_back(e); }
      if (k < 32) groups[g].a |= 1u << k; else groups[g].b |= 1u << (k - 32);
    }
    H.n_flat_groups = (int)groups.size();
    flat.insert(flat.end(), groups.begin(), groups.end());
  }

  // ---- materials / textures / lights
  mats.resize(d->n_materials);
  for (int i = 0; i < d->n_materials; i++) {
    const gopbrt_material& m = d->materials[i];
    if (m.tex_a >= d->n_textures || m.tex_b >= d->n_textures) return bad("texture index out of range");
    if (m.kind < 0 || m.kind > 2 || m.tex_a < 0 || (m.kind == 2 && m.tex_b < 0)) return bad("bad material");
    mats[i] = MaterialDev{m.kind, m.tex_a, m.tex_b, 0, m.sigma, m.eta, m.u_rough, m.v_rough};
  }
  texs.resize(d->n_textures);
  for (int i = 0; i < d->n_textures; i++) {
    const gopbrt_texture& t = d->textures[i];
    // checkerboard children must come earlier in the table: the texture graph is a DAG by construction (tex_eval recurses)
    if (t.kind == GOPBRT_TEX_CHECKERBOARD && (t.tex1 < 0 || t.tex2 < 0 || t.tex1 >= i || t.tex2 >= i)) return bad("checkerboard child index must be smaller than the texture's own");
    TextureDev o;
    o.kind = t.kind; o.mapping = t.mapping; o.tex1 = t.tex1; o.tex2 = t.tex2;
    if (t.kind == GOPBRT_TEX_CHECKERBOARD && (t.tex1 < 0 || t.tex1 >= d->n_textures || t.tex2 < 0 || t.tex2 >= d->n_textures)) return bad("bad checkerboard children");
    for (int k = 0; k < 3; k++) { o.rgb[k] = t.rgb[k]; o.vs[k] = t.vs[k]; o.vt[k] = t.vt[k]; }
    o.ds = t.ds; o.dt = t.dt; o.su = t.su; o.sv = t.sv; o.du = t.du; o.dv = t.dv;
    texs[i] = o;
  }
  lights.resize(d->n_lights);
  for (int i = 0; i < d->n_lights; i++) {
    const gopbrt_light& l = d->lights[i];
    LightDev o;
    o.kind = l.kind; o.shape_kind = l.shape_kind; o.shape_index = l.shape_index; o.two_sided = l.two_sided;
    for (int k = 0; k < 3; k++) { o.rgb[k] = l.rgb[k]; o.v[k] = l.v[k]; }
    if (l.kind == GOPBRT_LIGHT_DIFFUSE_AREA) {
      if (l.shape_kind == GOPBRT_SHAPE_SPHERE) { if (l.shape_index < 0 || l.shape_index >= d->n_spheres) return bad("light shape index"); }
      else if (l.shape_kind == GOPBRT_SHAPE_DISK) { if (l.shape_index < 0 || l.shape_index >= d->n_disks) return bad("light shape index"); }
      else return bad("area light shape must be a sphere or a disk");
    }
    lights[i] = o;
  }
  H.nl = d->n_lights;
  H.func_int = uniform_light_cdf(H.nl, cdf);
  return GOPBRT_OK;
}


// GOPBRT_CHECK_BVH: walks a device-built tree on the host, the way the traversal kernels read it — every primitive in exactly
// one leaf, leaves within maxPrims, every record's float32 box around the float64 bounds of everything below it.
static bool check_device_bvh(const gopbrt_scene_desc* d, const HostScene& H, const gpbvh::Node32* d_nodes, size_t n_records,
                             const std::vector<unsigned>& order, int max_prims, std::string& err) {
  std::vector<gpbvh::Node32> nodes(n_records);
  if (cudaMemcpy(nodes.data(), d_nodes, n_records * sizeof(gpbvh::Node32), cudaMemcpyDeviceToHost) != cudaSuccess) { err = "download failed"; return false; }
  const size_t np = order.size();
  std::vector<unsigned char> seen(np, 0);
  auto bound = [&](unsigned pi, double* mn, double* mx) {
    const gopbrt_primitive& p = d->primitives[pi];
    if (p.shape_kind == GOPBRT_SHAPE_TRIANGLE) {
      const gopbrt_triangle& t = d->triangles[p.shape_index];
      for (int k = 0; k < 3; k++) { mn[k] = INFINITY; mx[k] = -INFINITY; }
      for (int v = 0; v < 3; v++) for (int k = 0; k < 3; k++) { double c = d->vertices[3 * (size_t)t.v[v] + k]; mn[k] = std::min(mn[k], c); mx[k] = std::max(mx[k], c); }
    } else {
      const double* q = &H.qbounds[6 * (size_t)H.qslot[pi]];
      for (int k = 0; k < 3; k++) { mn[k] = q[k]; mx[k] = q[3 + k]; }
    }
  };
  long long errors = 0;
  struct Range { double mn[3], mx[3]; };
  std::function<Range(size_t, int)> visit = [&](size_t rec, int depth) -> Range {
    Range acc;
    for (int k = 0; k < 3; k++) { acc.mn[k] = INFINITY; acc.mx[k] = -INFINITY; }
    if (rec >= nodes.size() || depth > 70) { errors++; return acc; }
    const gpbvh::Node32& n = nodes[rec];
    if (n.a == gpbvh::kEmptyWord) { errors++; return acc; }
    if (n.a & 1u) {
      unsigned cnt = ((n.a >> 1) & 3u) + 1u, first = n.a >> 3;
      if ((int)cnt > max_prims) errors++;
      for (unsigned k = 0; k < cnt; k++) {
        size_t r = (size_t)first + k;
        if (r >= np) { errors++; continue; }
        unsigned pi = order[r];
        if (pi >= np || seen[pi]) { errors++; continue; }
        seen[pi] = 1;
        double mn[3], mx[3];
        bound(pi, mn, mx);
        for (int q = 0; q < 3; q++) { acc.mn[q] = std::min(acc.mn[q], mn[q]); acc.mx[q] = std::max(acc.mx[q], mx[q]); }
      }
    } else {
      size_t g = 4 * (size_t)(n.a >> 7);
      if (g == 0 || g + 3 >= nodes.size()) { errors++; return acc; }
      for (int k = 0; k < 4; k++) {
        if (nodes[g + k].a == gpbvh::kEmptyWord) { if (k != 1 && k != 3) errors++; continue; }
        Range c = visit(g + k, depth + 1);
        for (int q = 0; q < 3; q++) { acc.mn[q] = std::min(acc.mn[q], c.mn[q]); acc.mx[q] = std::max(acc.mx[q], c.mx[q]); }
      }
    }
    for (int q = 0; q < 3; q++) if (!((double)n.mn[q] <= acc.mn[q] && (double)n.mx[q] >= acc.mx[q])) errors++;
    return acc;
  };
  visit(0, 0);
  long long missing = 0;
  for (size_t i = 0; i < np; i++) missing += !seen[i];
  if (errors || missing) { err = std::to_string(errors) + " structural / box errors, " + std::to_string(missing) + " primitives not in any leaf"; return false; }
  return true;
}

// uploads a host scene to the context's device and binds the traversal kernels
static int upload_scene(gopbrt_ctx* ctx, HostScene& H, gopbrt_scene** out) {
  std::lock_guard<std::mutex> g(ctx->mu);
  std::lock_guard<std::mutex> grun(ctx->run_mu);
  GP_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<double>& xf = H.xf;
  std::vector<int>& xf_flags = H.xf_flags;
  std::vector<SphereDev>& spheres = H.spheres;
  std::vector<DiskDev>& disks = H.disks;
  std::vector<int4>& prims = H.prims;
  HB world = H.world;  // (a copy: the device build fills it per device, and devices upload concurrently)
  gpbvh::Result& bvh = H.bvh;
  std::vector<PrimRec>& recs = H.recs;
  std::vector<double>& rec_bounds = H.rec_bounds;
  std::vector<gpbvh::Node32>& flat = H.flat;
  const unsigned long long flat_tri_mask = H.flat_tri_mask;
  std::vector<MaterialDev>& mats = H.mats;
  std::vector<TextureDev>& texs = H.texs;
  std::vector<LightDev>& lights = H.lights;
  std::vector<double>& cdf = H.cdf;
  const double func_int = H.func_int;
  const int nl = H.nl;
  gopbrt_scene* sc = new gopbrt_scene();
  sc->ctx = ctx;
  cudaStream_t st = ctx->stream;
  std::vector<gpbvh::Node32>& nodes = bvh.nodes;
  auto up0 = std::chrono::steady_clock::now();
  const bool host_timing = getenv("GOPBRT_HOST_TIMING") != nullptr;
  int dev_depth = 0;
  size_t dev_records = 0;
  bool quadrics = false;
  if (H.device_build) {
    // ---- the caller's arrays go up as they are; tree, child groups and leaf-ordered records are made on the device
    const gopbrt_scene_desc* d = H.desc;
    const long long np = d->n_primitives;
    static_assert(sizeof(gopbrt_primitive) == sizeof(int4) && sizeof(gopbrt_triangle) == sizeof(int4), "descriptor tables are uploaded verbatim");
    DevBuf<double> d_vertices, d_qbounds;
    DevBuf<int4> d_triangles;
    DevBuf<int> d_qslot;
    DevBuf<PrimRec> d_qrecs;
    DevBuf<unsigned char> d_mnl;
    auto up = [&](void* dst, const void* src, size_t bytes) { return bytes == 0 || cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st) == cudaSuccess; };
    quadrics = !H.qrecs.empty();
    bool ok = sc->prims.alloc((size_t)np) == cudaSuccess && d_vertices.alloc(3 * (size_t)d->n_vertices) == cudaSuccess &&
              d_triangles.alloc((size_t)d->n_triangles) == cudaSuccess && sc->recs.alloc((size_t)np) == cudaSuccess &&
              sc->rec_bounds.alloc(quadrics ? 6 * (size_t)np : 0) == cudaSuccess && d_qslot.upload(H.qslot, st) == cudaSuccess &&
              d_qrecs.upload(H.qrecs, st) == cudaSuccess && d_qbounds.upload(H.qbounds, st) == cudaSuccess && d_mnl.upload(H.mat_not_lambert, st) == cudaSuccess &&
              up(sc->prims.p, d->primitives, (size_t)np * sizeof(int4)) && up(d_vertices.p, d->vertices, 3 * (size_t)d->n_vertices * sizeof(double)) &&
              up(d_triangles.p, d->triangles, (size_t)d->n_triangles * sizeof(int4));
    if (!ok) { ctx->last_error = std::string("scene upload: ") + cudaGetErrorString(cudaGetLastError()); delete sc; return GOPBRT_ERR_CUDA; }
    if (host_timing) {
      cudaStreamSynchronize(st);
      fprintf(stderr, "[gopbrt scene] %-22s %8.3f ms\n", "upload (raw arrays)", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - up0).count());
      up0 = std::chrono::steady_clock::now();
    }
    gpbuild::Input in;
    in.prims = sc->prims.p; in.n = np;
    in.vertices = d_vertices.p; in.n_vertices = d->n_vertices;
    in.triangles = d_triangles.p; in.n_triangles = d->n_triangles;
    in.qslot = d_qslot.p; in.qrecs = d_qrecs.p; in.qbounds = d_qbounds.p;
    in.mat_not_lambert = d_mnl.p; in.n_materials = d->n_materials;
    in.max_prims = H.max_prims;
    gpbuild::Output bo;
    bo.recs = sc->recs.p; bo.rec_bounds = sc->rec_bounds.p;
    bo.keep_host = getenv("GOPBRT_CHECK_BVH") != nullptr;
    std::string berr;
    if (!gpbuild::build(in, bo, kStackDepth - 1, st, berr)) {
      if (bo.nodes) cudaFree(bo.nodes);
      ctx->last_error = berr;
      delete sc;
      cudaGetLastError();
      return bo.cuda_error ? GOPBRT_ERR_CUDA : GOPBRT_ERR_INVALID;
    }
    ctx->launches += bo.launches;
    sc->nodes.p = (gpbvh::Node32*)bo.nodes; sc->nodes.n = bo.n_records;
    dev_depth = bo.depth; dev_records = bo.n_records;
    sc->class_mask = bo.class_mask;
    world.valid = bo.world_valid;
    for (int k = 0; k < 3; k++) { world.mn[k] = bo.world[k]; world.mx[k] = bo.world[3 + k]; }
    if (host_timing) {
      fprintf(stderr, "[gopbrt scene] %-22s %8.3f ms (%zu records, depth %d, %llu launches)\n", "device build", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - up0).count(),
              bo.n_records, bo.depth, (unsigned long long)bo.launches);
      up0 = std::chrono::steady_clock::now();
    }
    if (bo.keep_host) {
      std::string cerr_;
      if (!check_device_bvh(d, H, sc->nodes.p, bo.n_records, bo.h_order, H.max_prims, cerr_)) { ctx->last_error = "GOPBRT_CHECK_BVH: " + cerr_; delete sc; return GOPBRT_ERR_INVALID; }
      if (host_timing) fprintf(stderr, "[gopbrt scene] device BVH checked on the host: ok\n");
    }
  }
  bool ok = (H.device_build || (sc->nodes.upload(nodes, st) == cudaSuccess && sc->recs.upload(recs, st) == cudaSuccess &&
            sc->rec_bounds.upload(rec_bounds, st) == cudaSuccess && sc->prims.upload(prims, st) == cudaSuccess)) && sc->flat.upload(flat, st) == cudaSuccess &&
            sc->xf.upload(xf, st) == cudaSuccess && sc->xf_flags.upload(xf_flags, st) == cudaSuccess &&
            sc->spheres.upload(spheres, st) == cudaSuccess && sc->disks.upload(disks, st) == cudaSuccess &&
            sc->materials.upload(mats, st) == cudaSuccess && sc->textures.upload(texs, st) == cudaSuccess &&
            sc->lights.upload(lights, st) == cudaSuccess && sc->light_cdf.upload(cdf, st) == cudaSuccess &&
            sc->tctr.alloc(1) == cudaSuccess && cudaMemsetAsync(sc->tctr.p, 0, sizeof(TraceCounters), st) == cudaSuccess &&
            sc->work.alloc(2) == cudaSuccess && cudaMemsetAsync(sc->work.p, 0, 2 * sizeof(int), st) == cudaSuccess &&
            cudaStreamSynchronize(st) == cudaSuccess;
  if (host_timing)
    fprintf(stderr, "[gopbrt scene] %-22s %8.3f ms\n", "upload (tables)", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - up0).count());
  if (!ok) {
    ctx->last_error = std::string("scene upload: ") + cudaGetErrorString(cudaGetLastError());
    delete sc;
    return GOPBRT_ERR_CUDA;
  }
  DevScene& D = sc->dev;
  D.nodes = (const float4*)sc->nodes.p; D.recs = sc->recs.p; D.rec_bounds = sc->rec_bounds.p; D.prims = sc->prims.p;
  D.xf = sc->xf.p; D.xf_flags = sc->xf_flags.p; D.spheres = sc->spheres.p; D.disks = sc->disks.p;
  D.materials = sc->materials.p; D.textures = sc->textures.p; D.lights = sc->lights.p; D.light_cdf = sc->light_cdf.p;
  D.n_lights = nl; D.light_func_int = func_int; D.n_nodes = H.device_build ? (int)dev_records : (int)nodes.size();
  D.flat = (const float4*)sc->flat.p; D.n_flat = H.n_flat_entries; D.n_flat_groups = H.n_flat_groups; D.flat_tri_mask = flat_tri_mask;
  // Distant.Preprocess → Bounds3.BoundingSphere (distant.go:36-38, bounds.go:105-112)
  D.world_radius = 0;
  if (world.valid) {
    V3 c = (mk3(world.mn[0], world.mn[1], world.mn[2]) + mk3(world.mx[0], world.mx[1], world.mx[2])) / 2.0;
    bool inside = c.x >= world.mn[0] && c.x <= world.mx[0] && c.y >= world.mn[1] && c.y <= world.mx[1] && c.z >= world.mn[2] && c.z <= world.mx[2];
    if (inside) D.world_radius = sqrt(dist2(c, mk3(world.mx[0], world.mx[1], world.mx[2])));
    for (int k = 0; k < 3; k++) { sc->world[k] = world.mn[k]; sc->world[3 + k] = world.mx[k]; }
  }
  for (const PrimRec& r : recs) sc->class_mask |= 1u << ((r.flags & RF_CLASS_MASK) >> RF_CLASS_SHIFT);
  const int depth = H.device_build ? dev_depth : bvh.depth;
  sc->bvh_nodes = H.device_build ? dev_records : nodes.size();
  sc->bvh_depth = (uint64_t)depth;
  // traversal stack, [entry][thread] in dynamic shared memory: a step over a 4-record child group (two tree levels)
  // stacks at most three records
  sc->stack_cap = std::min(250, 3 * ((depth + 2) / 2) + 4);
  if (const char* e = getenv("GOPBRT_STACK_CAP")) sc->stack_cap = std::max(4, atoi(e));  // tuning aid (overflows are counted)
  sc->trace_smem = (size_t)sc->stack_cap * kTraceThreads * sizeof(unsigned);  // one node word per stacked node
  // scenes of triangles only get the traversal kernels that carry no sphere / disk (EFloat) code
  for (const PrimRec& r : recs) if ((r.flags & RK_KIND_MASK) != RK_TRIANGLE) { quadrics = true; break; }
  if (getenv("GOPBRT_GENERAL_TRACE")) quadrics = true;  // tuning aid: the general kernels on a triangle-only scene
  static const trace_fn k_general[7] = {k_trace<0, false, true>, k_trace<0, true, true>, k_trace<2, false, true>, k_trace<2, true, true>,
                                        k_trace<1, false, true>, k_trace<3, false, true>, k_trace<3, true, true>};
  static const trace_fn k_triangles[7] = {k_trace<0, false, false>, k_trace<0, true, false>, k_trace<2, false, false>, k_trace<2, true, false>,
                                          k_trace<1, false, false>, k_trace<3, false, false>, k_trace<3, true, false>};
  for (int k = 0; k < 7; k++) sc->trace_k[k] = quadrics ? k_general[k] : k_triangles[k];
  if (D.n_flat > 0) {  // small scene: the flat aggregate answers every query (no traversal stack)
    sc->trace_k[0] = k_trace_flat<0, false>; sc->trace_k[1] = k_trace_flat<0, true>; sc->trace_k[2] = k_trace_flat<2, false>;
    sc->trace_k[3] = k_trace_flat<2, true>; sc->trace_k[4] = k_trace_flat<1, false>;
    sc->trace_k[5] = k_trace_flat<3, false>; sc->trace_k[6] = k_trace_flat<3, true>;
    sc->trace_smem = 0;
  }
  // cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the kernel, not to the scene: it only ever grows (a later, shallower
  // scene must not lower the limit under a deeper one that is still alive)
  if (sc->trace_smem > 48 * 1024 && sc->trace_smem > ctx->trace_smem_limit) {
    for (int k = 0; k < 14; k++) {
      cudaError_t e = cudaFuncSetAttribute((const void*)(k < 7 ? k_general[k] : k_triangles[k - 7]), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc->trace_smem);
      if (e != cudaSuccess) { ctx->last_error = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); delete sc; return GOPBRT_ERR_CUDA; }
    }
    ctx->trace_smem_limit = sc->trace_smem;
  }
  for (int k = 0; k < 7; k++) sc->trace_grid[k] = grid_for(ctx, (const void*)sc->trace_k[k], kTraceThreads, sc->trace_smem);
  *out = sc;
  return GOPBRT_OK;
}

extern "C" int gopbrt_scene_create(gopbrt_ctx* ctx, const gopbrt_scene_desc* d, gopbrt_scene** out) {
  if (!ctx || !d || !out) return GOPBRT_ERR_INVALID;
  *out = nullptr;
  HostScene H;
  std::string err;
  int rc = build_host_scene(d, H, err);
  if (rc != GOPBRT_OK) { std::lock_guard<std::mutex> g(ctx->mu); ctx->last_error = err; return rc; }
  return upload_scene(ctx, H, out);
}

extern "C" void gopbrt_scene_destroy(gopbrt_scene* sc) {
  if (!sc) return;
  std::lock_guard<std::mutex> grun(sc->ctx->run_mu);
  cudaSetDevice(sc->ctx->device);
  cudaStreamSynchronize(sc->ctx->stream);
  delete sc;
}

extern "C" int gopbrt_scene_world_bound(const gopbrt_scene* sc, double out6[6]) {
  if (!sc || !out6) return GOPBRT_ERR_INVALID;
  memcpy(out6, sc->world, sizeof(sc->world));
  return GOPBRT_OK;
}

extern "C" int gopbrt_cancel(gopbrt_scene* sc) {
  if (!sc) return GOPBRT_ERR_INVALID;
  sc->cancel.store(1);
  return GOPBRT_OK;
}

// ------------------------------------------------------------------------------------------------ batched trace API
static RaySoA soa7(double* base, long long n) {
  RaySoA r;
  r.ox = base; r.oy = base + n; r.oz = base + 2 * n; r.dx = base + 3 * n; r.dy = base + 4 * n; r.dz = base + 5 * n; r.tmax = base + 6 * n;
  return r;
}

// closest hit over device-resident SoA rays: packs them into RayRec, traces, unpacks.  prim (may be null) receives
// primitive ids, rec (may be null) leaf-record indices, t the hit distance (tmax where nothing was hit).
static int trace_closest_soa_device(gopbrt_scene* sc, int64_t n, const double* rays_soa7, int32_t* prim, int32_t* rec, double* t, cudaStream_t st) {
  gopbrt_ctx* ctx = sc->ctx;
  DevBuf<RayRec>& recs = sc->pool.recs;
  GP_CUDA(ctx, recs.ensure((size_t)n));
  RaySoA r = soa7(const_cast<double*>(rays_soa7), n);
  int gs = ctx->sm_count * 8;
  k_pack_rays<<<gs, 256, 0, st>>>(r, recs.p, n);
  long long need = (n + kTraceThreads - 1) / kTraceThreads;
  size_t smem = sc->trace_smem;
  GP_CUDA(ctx, cudaMemsetAsync(sc->work.p, 0, sizeof(int), st));
  sc->trace_k[0]<<<(int)std::min<long long>(sc->trace_grid[0], need), kTraceThreads, smem, st>>>(sc->dev, recs.p, nullptr, nullptr, nullptr, nullptr, nullptr, n,
                                                                                                sc->stack_cap, sc->work.p, sc->tctr.p, nullptr);
  k_unpack_hits<<<gs, 256, 0, st>>>(sc->dev, recs.p, prim, rec, t, n);
  ctx->launches += 3;
  GP_CUDA(ctx, cudaGetLastError());
  GP_CUDA(ctx, cudaStreamSynchronize(st));  // the pooled records are reused by the next call
  return GOPBRT_OK;
}

// the traversal kernels index rays with 32-bit counters
static const int64_t kMaxBatchRays = 0x7fffff00;

extern "C" int gopbrt_trace_closest_device(gopbrt_scene* sc, int64_t n, const double* rays_soa7, int32_t* prim, double* t, void* stream) {
  if (!sc || n < 0 || n > kMaxBatchRays || (n > 0 && (!rays_soa7 || !t))) return GOPBRT_ERR_INVALID;
  if (n == 0) return GOPBRT_OK;
  std::lock_guard<std::mutex> grun(sc->ctx->run_mu);
  GP_CUDA(sc->ctx, cudaSetDevice(sc->ctx->device));
  return trace_closest_soa_device(sc, n, rays_soa7, prim, nullptr, t, stream ? (cudaStream_t)stream : sc->ctx->stream);
}

static int trace_any_soa_device(gopbrt_scene* sc, int64_t n, const double* rays_soa7, uint8_t* hit, void* stream);
extern "C" int gopbrt_trace_any_device(gopbrt_scene* sc, int64_t n, const double* rays_soa7, uint8_t* hit, void* stream) {
  if (!sc || n < 0 || n > kMaxBatchRays || (n > 0 && (!rays_soa7 || !hit))) return GOPBRT_ERR_INVALID;
  if (n == 0) return GOPBRT_OK;
  std::lock_guard<std::mutex> grun(sc->ctx->run_mu);
  GP_CUDA(sc->ctx, cudaSetDevice(sc->ctx->device));
  return trace_any_soa_device(sc, n, rays_soa7, hit, stream);
}
// callers hold the context's run_mu
static int trace_any_soa_device(gopbrt_scene* sc, int64_t n, const double* rays_soa7, uint8_t* hit, void* stream) {
  gopbrt_ctx* ctx = sc->ctx;
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  DevBuf<RayRec>& recs = sc->pool.recs;
  GP_CUDA(ctx, recs.ensure((size_t)n));
  RaySoA r = soa7(const_cast<double*>(rays_soa7), n);
  k_pack_rays<<<ctx->sm_count * 8, 256, 0, st>>>(r, recs.p, n);
  long long need = (n + kTraceThreads - 1) / kTraceThreads;
  size_t smem = sc->trace_smem;
  GP_CUDA(ctx, cudaMemsetAsync(sc->work.p + 1, 0, sizeof(int), st));
  sc->trace_k[4]<<<(int)std::min<long long>(sc->trace_grid[4], need), kTraceThreads, smem, st>>>(sc->dev, recs.p, nullptr, nullptr, hit, nullptr, nullptr, n,
                                                                                                sc->stack_cap, sc->work.p + 1, sc->tctr.p, nullptr);
  ctx->launches += 2;
  GP_CUDA(ctx, cudaGetLastError());
  GP_CUDA(ctx, cudaStreamSynchronize(st));
  return GOPBRT_OK;
}

extern "C" int gopbrt_trace_closest(gopbrt_scene* sc, int64_t n, const double* ox, const double* oy, const double* oz, const double* dx,
                                    const double* dy, const double* dz, const double* tmax, int32_t* prim, double* t, double* p, double* nrm) {
  if (!sc || n < 0 || n > kMaxBatchRays || (n > 0 && (!ox || !oy || !oz || !dx || !dy || !dz || !tmax || !prim || !t))) return GOPBRT_ERR_INVALID;
  if (n == 0) return GOPBRT_OK;
  gopbrt_ctx* ctx = sc->ctx;
  std::lock_guard<std::mutex> g(sc->mu);
  std::lock_guard<std::mutex> grun(sc->ctx->run_mu);
  GP_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  BatchPool& B = sc->pool;
  DevBuf<double>&rays = B.rays, &tt = B.t, &pp = B.p, &nn = B.n;
  DevBuf<int>&rec = B.rec, &pr = B.prim;
  GP_CUDA(ctx, rays.ensure(7 * (size_t)n));
  GP_CUDA(ctx, tt.ensure(n));
  GP_CUDA(ctx, rec.ensure(n));
  GP_CUDA(ctx, pr.ensure(n));
  if (p) GP_CUDA(ctx, pp.ensure(3 * (size_t)n));
  if (nrm) GP_CUDA(ctx, nn.ensure(3 * (size_t)n));
  GP_CUDA(ctx, B.ensure_stage(7 * (size_t)n));
  const double* src[7] = {ox, oy, oz, dx, dy, dz, tmax};
  for (int k = 0; k < 7; k++) memcpy(B.stage + (size_t)k * n, src[k], (size_t)n * sizeof(double));
  GP_CUDA(ctx, cudaMemcpyAsync(rays.p, B.stage, 7 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  int rc = trace_closest_soa_device(sc, n, rays.p, nullptr, rec.p, tt.p, st);
  if (rc != GOPBRT_OK) return rc;
  RaySoA r = soa7(rays.p, n);
  r.tmax = tt.p;
  k_hit_points<<<ctx->sm_count * 4, 128, 0, st>>>(sc->dev, r, rec.p, n, pr.p, p ? pp.p : nullptr, nrm ? nn.p : nullptr);
  ctx->launches++;
  GP_CUDA(ctx, cudaGetLastError());
  GP_CUDA(ctx, cudaMemcpyAsync(prim, pr.p, n * sizeof(int), cudaMemcpyDeviceToHost, st));
  GP_CUDA(ctx, cudaMemcpyAsync(t, tt.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (p) GP_CUDA(ctx, cudaMemcpyAsync(p, pp.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (nrm) GP_CUDA(ctx, cudaMemcpyAsync(nrm, nn.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
  GP_CUDA(ctx, cudaStreamSynchronize(st));
  return GOPBRT_OK;
}

extern "C" int gopbrt_trace_any(gopbrt_scene* sc, int64_t n, const double* ox, const double* oy, const double* oz, const double* dx,
                                const double* dy, const double* dz, const double* tmax, uint8_t* hit) {
  if (!sc || n < 0 || n > kMaxBatchRays || (n > 0 && (!ox || !oy || !oz || !dx || !dy || !dz || !tmax || !hit))) return GOPBRT_ERR_INVALID;
  if (n == 0) return GOPBRT_OK;
  gopbrt_ctx* ctx = sc->ctx;
  std::lock_guard<std::mutex> g(sc->mu);
  std::lock_guard<std::mutex> grun(sc->ctx->run_mu);
  GP_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  BatchPool& B = sc->pool;
  DevBuf<double>& rays = B.rays;
  DevBuf<unsigned char>& h = B.hit;
  GP_CUDA(ctx, rays.ensure(7 * (size_t)n));
  GP_CUDA(ctx, h.ensure(n));
  GP_CUDA(ctx, B.ensure_stage(7 * (size_t)n));
  const double* src[7] = {ox, oy, oz, dx, dy, dz, tmax};
  for (int k = 0; k < 7; k++) memcpy(B.stage + (size_t)k * n, src[k], (size_t)n * sizeof(double));
  GP_CUDA(ctx, cudaMemcpyAsync(rays.p, B.stage, 7 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  int rc = trace_any_soa_device(sc, n, rays.p, h.p, st);
  if (rc != GOPBRT_OK) return rc;
  GP_CUDA(ctx, cudaMemcpyAsync(hit, h.p, n, cudaMemcpyDeviceToHost, st));
  GP_CUDA(ctx, cudaStreamSynchronize(st));
  return GOPBRT_OK;
}

// Film geometry of a render: NewFilm's cropped pixel bounds (film.go:43-48), pbrt.Render's tile grid (integrator.go:297-299)
// and the extent of a lane's FilmTile storage.  Returns an error text, or nullptr.
static const char* film_params(const gopbrt_film* film, int64_t tile_size, RenderParams& P) {
  if (tile_size < 1) return "tile_size < 1";
  if (!(film->filter_radius[0] > 0) || !(film->filter_radius[1] > 0)) return "filter radius must be positive";
  P.tile_size = tile_size;
  P.cx0 = (long long)ceil((double)film->width * film->crop[0]); P.cy0 = (long long)ceil((double)film->height * film->crop[1]);
  P.cx1 = (long long)ceil((double)film->width * film->crop[2]); P.cy1 = (long long)ceil((double)film->height * film->crop[3]);
  long long fw = P.cx1 - P.cx0, fh = P.cy1 - P.cy0;
  if (fw <= 0 || fh <= 0) return "empty film";
  P.ntx = (fw + P.tile_size - 1) / P.tile_size; P.nty = (fh + P.tile_size - 1) / P.tile_size;  // integrator.go:297-299
  P.ntiles = P.ntx * P.nty;
  P.frx = film->filter_radius[0]; P.fry = film->filter_radius[1];
  // Columns / rows of a tile's pixel bounds (GetFilmTile, film.go:106-113: [ceil(x0 - 0.5 - r), floor(x1 - 0.5 + r) + 1), x0
  // integer, x1 - x0 <= tileSize) that AddSample can actually touch: samples lie strictly below x1, so the last pixel a
  // sample reaches is ceil(x1 - 0.5 + r) - 1 — one short of the bound when x1 - 0.5 + r is an integer (box filter r = 0.5).
  // The untouched last column / row stays zero in the reference's tile and adds exactly nothing in MergeFilmTile, so it is
  // not stored: tileSize 1 with the box filter needs 2 x 2 pixels = one 128-byte line per lane.
  long long tpw = std::min<long long>(fw, (long long)ceil((double)P.tile_size - 0.5 + P.frx) - (long long)ceil(-0.5 - P.frx));
  long long tph = std::min<long long>(fh, (long long)ceil((double)P.tile_size - 0.5 + P.fry) - (long long)ceil(-0.5 - P.fry));
  P.tpw = (int)tpw; P.tph = (int)tph;
  return nullptr;
}

// ------------------------------------------------------------------------------------------------ render
static int render_impl(gopbrt_scene* sc, const gopbrt_camera* cam, const gopbrt_sampler* smp, const gopbrt_integrator* ig,
                       const gopbrt_film* film, const gopbrt_render_options* opt, double* d_film, gopbrt_stats* stats) {
  gopbrt_ctx* ctx = sc->ctx;
  auto bad = [&](const char* msg) { ctx->last_error = msg; return GOPBRT_ERR_INVALID; };
  if (ig->kind == GOPBRT_INTEGRATOR_PATH) {
    if (ig->light_strategy == GOPBRT_LIGHTS_SPATIAL) {  // CreateLightSampleDistribution returns nil (lightdistribution.go:11-19): Path.Li panics
      ctx->last_error = "Path: LightSampleStrategy Spatial has no distribution in the reference (nil LightDistribution, path.go:80 panics)";
      return GOPBRT_ERR_UNSUPPORTED;
    }
    if (ig->light_strategy != GOPBRT_LIGHTS_UNIFORM && ig->light_strategy != GOPBRT_LIGHTS_POWER) return bad("Path: unknown light sample strategy");
    if (ig->max_depth > 255) return bad("Path: max_depth out of range (the bounce count is kept in 8 bits)");
  } else if (ig->kind == GOPBRT_INTEGRATOR_DIRECT_LIGHTING) {
    if (ig->light_strategy != GOPBRT_DL_SAMPLE_ONE && ig->light_strategy != GOPBRT_DL_SAMPLE_ALL) return bad("DirectLighting: unknown lighting strategy");
    if (ig->light_strategy == GOPBRT_DL_SAMPLE_ALL && sc->dev.n_lights > 31) {
      ctx->last_error = "DirectLighting/UniformSampleAll: more than 31 lights (one shadow segment per light and lane)";
      return GOPBRT_ERR_UNSUPPORTED;
    }
    if (ig->max_depth > 250) return bad("DirectLighting: max_depth out of range");
  } else return bad("unknown integrator kind");
  if (smp->kind != GOPBRT_SAMPLER_STRATIFIED && smp->kind != GOPBRT_SAMPLER_RANDOM) return bad("unknown sampler");
  if (smp->n_sampled_dimensions > 255 || smp->n_sampled_dimensions < 0) return bad("n_sampled_dimensions out of range");
  int rank = opt ? opt->rank : 0, world = opt ? opt->world : 1;
  int flags = opt ? opt->flags : 0;
  if (world < 1 || rank < 0 || rank >= world) return bad("bad rank/world");
  // A cancel request is consumed by the render call that observes it: the running one, or — when it arrived before any
  // wavefront was launched (ctx already done, integrator.go:332-336) — this one, right here.
  if (sc->cancel.exchange(0)) return GOPBRT_ERR_CANCELLED;

  RenderParams P;
  memset(&P, 0, sizeof(P));  // padding too: the struct is part of the CUDA-graph cache key
  P.raster_to_camera = m4_from(cam->raster_to_camera);
  P.camera_to_world = m4_from(cam->camera_to_world);
  P.lens_radius = cam->lens_radius; P.focal_distance = cam->focal_distance;
  P.shutter_open = cam->shutter_open; P.shutter_close = cam->shutter_close;
  P.sampler_kind = smp->kind; P.xs = smp->x_samples; P.ys = smp->y_samples; P.jitter = smp->jitter;
  P.ndims = smp->kind == GOPBRT_SAMPLER_STRATIFIED ? smp->n_sampled_dimensions : 0;
  P.mode = smp->mode;
  P.spp = smp->kind == GOPBRT_SAMPLER_STRATIFIED ? smp->x_samples * smp->y_samples : smp->x_samples;
  if (P.spp < 1) return bad("samples per pixel < 1");
  P.max_depth = ig->max_depth; P.rr_threshold = ig->rr_threshold;
  P.integrator = ig->kind;
  P.direct_levels = std::max(1, ig->max_depth / 2);
  const bool direct_all = ig->kind == GOPBRT_INTEGRATOR_DIRECT_LIGHTING && ig->light_strategy == GOPBRT_DL_SAMPLE_ALL;
  P.n_seg = direct_all ? std::max(1, sc->dev.n_lights) : 1;
  P.direct_all = direct_all ? 1 : 0;
  P.light_power = (ig->kind == GOPBRT_INTEGRATOR_PATH && ig->light_strategy == GOPBRT_LIGHTS_POWER) ? 1 : 0;
  // every sample of a lane through the integer corner of the lane's one pixel, filter weight 1: see film_add_uniform
  P.uniform_fp = (ig->tile_size == 1 && smp->kind == GOPBRT_SAMPLER_STRATIFIED && smp->n_sampled_dimensions >= 1 && !getenv("GOPBRT_NO_UNIFORM_FP")) ? 1 : 0;
  // ... and then a lane's last sample needs no trip through the regeneration queue (gp_render.cuh lane_on_last_sample)
  P.last_in_place = (P.uniform_fp && P.mode == GOPBRT_MODE_FAST && ig->kind == GOPBRT_INTEGRATOR_PATH && !getenv("GOPBRT_NO_LAST_IN_PLACE")) ? 1 : 0;
  if (const char* ferr = film_params(film, ig->tile_size, P)) return bad(ferr);
  long long fw = P.cx1 - P.cx0, fh = P.cy1 - P.cy0;
  long long tpw = P.tpw, tph = P.tph;
  if (P.mode == GOPBRT_MODE_STRICT) { P.rank = rank; P.world = world; P.s_rank = 0; P.s_world = 1; }
  else { P.rank = 0; P.world = 1; P.s_rank = rank; P.s_world = world; }
  // FAST mode: a pixel's samples are independent, so a tile can be worked on by several lanes at once ("lane groups",
  // each taking every groups-th sample of this rank's share).  That turns the long per-lane sample chains of a 64-spp
  // frame — hundreds of wavefront iterations whose tail runs nearly empty — into many short ones, paid for with lane
  // state in HBM.  Bits 8..15 of the flags choose the group count; 0 = automatic.
  P.groups = 1;
  const long long tiles_rank = (P.ntiles - P.rank + P.world - 1) / P.world;
  bool auto_groups = false;
  int share = 1;
  if (P.mode == GOPBRT_MODE_FAST) {
    int want = (flags >> 8) & 0xff;
    share = std::max(1, (P.spp - 1 + P.s_world - 1) / P.s_world);  // samples 1..spp-1 of every pixel, split over the ranks
    // automatic (decided below, once the lane-state budget is known): ONE sample per lane while that state stays small, else two
    // (measured on one B200, config 2 / 63 spp: 8 samples per lane 186 ms, 4: 167 ms, 2: 149 ms, 1: 148 ms — the last halving
    // doubles 21 GB of lane state for 1 %; on a rank's share of the frame, where the state is small, it is worth 4-5 %: world 2
    // 69.1 -> 66.5 ms, world 4 34.7 -> 33.3 ms, world 8 17.5 -> 16.6 ms)
    if (want == 0) { want = std::min(64, share); auto_groups = true; }
    P.groups = std::max(1, std::min(want, share));
  }

  // GOPBRT_HOST_TIMING=1: host wall-clock of the call's phases on stderr (tuning aid)
  const bool host_timing = getenv("GOPBRT_HOST_TIMING") != nullptr;
  auto hw0 = std::chrono::steady_clock::now();
  auto hw_tick = [&](const char* what) {
    if (!host_timing) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[gopbrt host] %-18s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - hw0).count());
    hw0 = now;
  };
  cudaStream_t st = ctx->stream;
  GP_CUDA(ctx, cudaSetDevice(ctx->device));
  GP_CUDA(ctx, cudaMemsetAsync(d_film, 0, (size_t)fw * fh * 4 * sizeof(double), st));

  // ---- workspace
  // the stratified 1-D tables exist only in STRICT mode (FAST derives a sample's stratum from a hashed permutation)
  const size_t table_doubles = P.mode == GOPBRT_MODE_FAST ? 0 : (size_t)P.ndims * P.spp;
  const size_t frame_doubles = P.integrator == GOPBRT_INTEGRATOR_DIRECT_LIGHTING ? (size_t)P.direct_levels * 8 : 0;
  const size_t n_seg = (size_t)P.n_seg;  // shadow segments (and shadow-queue entries) per lane
  size_t per_lane = sizeof(RayRec) + n_seg * (sizeof(ShadowRec) + 1) + sizeof(PathRec) + sizeof(RadRec) + sizeof(FilmRec) + (8 + n_seg) * 4 + 1 + table_doubles * 8 + (P.uniform_fp ? 0 : (size_t)tpw * tph * 4 * 8) + frame_doubles * 8;
  Workspace& W = sc->ws;
  // Lane state of ONE scene handle: at most 40 % of the device's memory (GOPBRT_LANE_BUDGET_GB overrides) and 80 % of what is
  // free right now — several scenes render on one context (one per gRPC request in the reference) and must not starve each other.
  // The device is asked for its free memory only when the request differs from the one this workspace was sized for: the
  // query takes a driver lock that monitoring tools (NVML, nvidia-smi) hold for milliseconds at a time.
  const long long cap_key[5] = {tiles_rank, (long long)P.groups * 2 + (auto_groups ? 1 : 0), (long long)per_lane, opt ? (long long)opt->max_lanes : 0, (long long)P.mode};
  long long cap = W.cap;
  if (W.lanes != 0 && memcmp(cap_key, W.cap_key, sizeof(cap_key)) == 0) {
    if (auto_groups) P.groups = W.auto_groups;
  } else {
    size_t free_b = 0, total_b = 0;
    GP_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
    size_t held = W.lanes ? (size_t)W.lanes * W.per_lane : 0;
    double budget = 0.40 * (double)total_b;
    if (const char* e = getenv("GOPBRT_LANE_BUDGET_GB")) budget = std::max(0.25, atof(e)) * 1e9;
    budget = std::min(budget, (double)(free_b + held) * 0.80);
    cap = (long long)(budget / (double)per_lane);
    if (opt && opt->max_lanes > 0) cap = std::min<long long>(cap, opt->max_lanes);
    if (auto_groups) {
      // one sample per lane only while the lane state stays within a quarter of the device's memory, else two per lane;
      // then the groups shrink until one pass holds every lane (more passes would serialise what the groups parallelise)
      if ((double)tiles_rank * P.groups * (double)per_lane > 0.25 * (double)total_b) P.groups = std::max(1, std::min(64, (share + 1) / 2));
      if (tiles_rank * P.groups > cap) P.groups = (int)std::max<long long>(1, std::min<long long>(P.groups, cap / std::max<long long>(1, tiles_rank)));
      W.auto_groups = P.groups;
    }
    W.cap = cap; W.per_lane = per_lane;
    memcpy(W.cap_key, cap_key, sizeof(cap_key));
  }
  long long lanes_total = tiles_rank * P.groups;
  long long lanes = std::max<long long>(1, std::min(lanes_total, cap));
  if (lanes * (long long)n_seg > 0x7fffff00LL) lanes = 0x7fffff00LL / (long long)n_seg;
  size_t bt = table_doubles * lanes, bp = P.uniform_fp ? 0 : (size_t)tpw * tph * 4 * lanes;
  if (W.lanes != lanes || W.bytes_tables != bt || W.bytes_tilepix != bp || W.frames.n != frame_doubles * (size_t)lanes || W.sray.n != n_seg * (size_t)lanes ||
      W.occl.n != (direct_all ? n_seg * (size_t)lanes : 0) || W.codes.n != (size_t)lanes) {
    // the shape fields are cleared first: a failed allocation below must not leave a "same shape" workspace with null buffers
    W.lanes = 0; W.bytes_tables = 0; W.bytes_tilepix = 0;
    if (W.graph_exec) { cudaGraphExecDestroy(W.graph_exec); W.graph_exec = nullptr; W.graph_key.clear(); }
    W.ray.release(); W.sray.release(); W.path.release(); W.rad.release(); W.fsum.release(); W.i32.release(); W.tables.release(); W.tilepix.release(); W.frames.release(); W.occl.release(); W.codes.release();
    GP_CUDA(ctx, W.ray.alloc((size_t)lanes));
    GP_CUDA(ctx, W.sray.alloc(n_seg * (size_t)lanes));
    GP_CUDA(ctx, W.occl.alloc(direct_all ? n_seg * (size_t)lanes : 0));
    GP_CUDA(ctx, W.path.alloc((size_t)lanes));
    GP_CUDA(ctx, W.rad.alloc((size_t)lanes));
    GP_CUDA(ctx, W.fsum.alloc((size_t)lanes));
    GP_CUDA(ctx, W.i32.alloc((8 + n_seg) * (size_t)lanes));
    GP_CUDA(ctx, W.codes.alloc((size_t)lanes));
    GP_CUDA(ctx, W.tables.alloc(std::max<size_t>(bt, 1)));
    GP_CUDA(ctx, W.tilepix.alloc(bp));
    GP_CUDA(ctx, W.frames.alloc(frame_doubles * (size_t)lanes));
    if (!W.cnt.p) GP_CUDA(ctx, W.cnt.alloc(16));
    if (!W.rctr.p) GP_CUDA(ctx, W.rctr.alloc(1));
    if (!W.remaining_host) {
      GP_CUDA(ctx, cudaHostAlloc((void**)&W.remaining_host, sizeof(int), cudaHostAllocMapped));
      GP_CUDA(ctx, cudaHostGetDevicePointer((void**)&W.remaining_dev, W.remaining_host, 0));
    }
    W.lanes = lanes; W.bytes_tables = bt; W.bytes_tilepix = bp;
  }
  hw_tick("workspace");
  Lanes L;
  memset(&L, 0, sizeof(L));
  L.n = lanes;
  L.ray = W.ray.p; L.sray = W.sray.p; L.path = W.path.p; L.rad = W.rad.p; L.fsum = W.fsum.p;
  int* ip = W.i32.p;
  Queues Q;
  memset(&Q, 0, sizeof(Q));
  Q.extend = ip; Q.extend_next = ip + lanes; Q.regen = ip + 2 * lanes; Q.regen_next = ip + 3 * lanes; Q.shade[0] = ip + 4 * lanes; Q.shade[1] = ip + 5 * lanes; Q.shade[2] = ip + 6 * lanes; Q.shade[3] = ip + 7 * lanes;
  Q.shadow = ip + 8 * lanes;  // n_seg entries per lane
  Q.cnt = W.cnt.p;
  L.tables = W.tables.p; L.tilepix = W.tilepix.p; L.frames = W.frames.p; L.occl = W.occl.p;
  L.tile_stride = (long long)tpw * tph * 4;

  GP_CUDA(ctx, cudaMemsetAsync(W.rctr.p, 0, sizeof(RenderCounters), st));
  GP_CUDA(ctx, cudaMemsetAsync(sc->tctr.p, 0, sizeof(TraceCounters), st));
  const bool count = (flags & GOPBRT_FLAG_COUNT_TRAVERSAL) != 0;

  // stage kernels of this render: raygen by (sampler mode, integrator), shade by (integrator / strategy, sampler mode) and,
  // for Path, one launch per shade class
  typedef void (*gen_fn)(DevScene, Lanes, RenderParams, Queues, const int*, const int*, RenderCounters*);
  typedef void (*shade_fn)(DevScene, Lanes, RenderParams, Queues, RenderCounters*);
  static const gen_fn gen_tab[8] = {k_generate<0, 0, 0>, k_generate<0, 1, 0>, k_generate<1, 0, 0>, k_generate<1, 1, 0>,
                                    k_generate<0, 0, 1>, k_generate<0, 1, 1>, k_generate<1, 0, 1>, k_generate<1, 1, 1>};
  static const shade_fn shade_tab[12] = {k_shade<0, 0, 0>, k_shade<0, 0, 1>, k_shade<0, 0, 2>, k_shade<0, 0, 3>,
                                         k_shade<0, 1, 0>, k_shade<0, 1, 1>, k_shade<0, 1, 2>, k_shade<0, 1, 3>,
                                         k_shade<1, 0, -1>, k_shade<1, 1, -1>, k_shade<2, 0, -1>, k_shade<2, 1, -1>};
  const int fast = P.mode == GOPBRT_MODE_FAST ? 1 : 0;
  const int gen_i = P.uniform_fp * 4 + fast * 2 + (P.integrator == GOPBRT_INTEGRATOR_PATH ? 0 : 1);
  const int shade_kind = P.integrator == GOPBRT_INTEGRATOR_PATH ? 0 : (direct_all ? 2 : 1);
  const int shade_i0 = shade_kind == 0 ? fast * 4 : 8 + (shade_kind - 1) * 2 + fast;
  int shade_list[4] = {shade_i0, 0, 0, 0}, n_shade = 0;  // Path: one launch per shade class that occurs in the scene
  if (shade_kind == 0) { for (int k = 0; k < 4; k++) if ((sc->class_mask >> k) & 1u) shade_list[n_shade++] = shade_i0 + k; }
  else n_shade = 1;
  if (!ctx->grid_gen[gen_i]) ctx->grid_gen[gen_i] = grid_for(ctx, (const void*)gen_tab[gen_i], 128);
  for (int k = 0; k < n_shade; k++)
    if (!ctx->grid_shade[shade_list[k]]) ctx->grid_shade[shade_list[k]] = grid_for(ctx, (const void*)shade_tab[shade_list[k]], 128);
  const gen_fn k_gen = gen_tab[gen_i];
  const size_t smem = sc->trace_smem;
  const int k_ext = count ? 1 : 0, k_any = direct_all ? (count ? 6 : 5) : (count ? 3 : 2);
  const int scap = sc->stack_cap;
  const int g_small = ctx->sm_count * 8;
  const int g_gen = ctx->grid_gen[gen_i];
  constexpr int kGraphIters = 8;

  cudaEvent_t* ev = W.frame_ev;  // the frame's two timing events live with the workspace (no create / destroy per frame)
  for (int k = 0; k < 2; k++) if (!ev[k]) GP_CUDA(ctx, cudaEventCreate(&ev[k]));
  GP_CUDA(ctx, cudaEventRecord(ev[0], st));
  uint64_t iterations = 0, launches0 = ctx->launches.load();
  int rc = GOPBRT_OK;
  // optional per-stage timing: CUDA events on the launching stream around every stage launch
  const bool timing = (flags & GOPBRT_FLAG_TIME_KERNELS) != 0;
  enum { ST_RAYGEN = 0, ST_EXTEND, ST_SHADE, ST_SHADOW, ST_FILM, ST_TAIL, ST_N };
  std::vector<int> ev_stage;
  size_t ev_used = 0;
  auto tick = [&](int stage) {
    if (!timing) return;
    if (ev_used >= W.events.size()) { cudaEvent_t e; cudaEventCreate(&e); W.events.push_back(e); }
    cudaEventRecord(W.events[ev_used++], st);
    ev_stage.push_back(stage);
  };
  uint64_t n_extend = 0, n_shadow = 0;
  // debug aid: GOPBRT_ITER_LOG=<file> synchronises every iteration and logs the queue sizes (implies per-stage timing)
  const char* iter_log_path = getenv("GOPBRT_ITER_LOG");
  std::vector<int> iter_counts, iter_shadow, iter_hits;
  for (long long base = 0; base < lanes_total && rc == GOPBRT_OK; base += lanes) {
    P.lane_base = base;
    P.lanes_active = std::min(lanes, lanes_total - base);
    if (bp) GP_CUDA(ctx, cudaMemsetAsync(W.tilepix.p, 0, bp * sizeof(double), st));
    GP_CUDA(ctx, cudaMemsetAsync(W.cnt.p, 0, 16 * sizeof(int), st));
    tick(ST_RAYGEN);
    k_gen<<<g_gen, 128, 0, st>>>(sc->dev, L, P, Q, nullptr, nullptr, W.rctr.p);
    ctx->launches += 1;
    // Wavefront iterations as a CUDA graph of kGraphIters iterations (6 launches each): in STRICT mode the last lanes
    // need hundreds of iterations over nearly empty queues, where the cost of an iteration IS its launch latency.
    // One graph stays queued ahead of the one the host is waiting for, so the device never idles on the host.
    const bool use_graph = !timing && !iter_log_path && !count && !getenv("GOPBRT_NO_GRAPH");
    if (use_graph) {
      auto enqueue_iteration = [&]() {
        sc->trace_k[k_ext]<<<sc->trace_grid[k_ext], kTraceThreads, smem, st>>>(sc->dev, L.ray, nullptr, L.rad, W.codes.p, Q.extend, Q.cnt + 0, 0, scap, Q.cnt + 6, sc->tctr.p, nullptr);
        k_split_hits<<<g_small, 256, 0, st>>>(L, Q, W.codes.p);
        for (int k = 0; k < n_shade; k++) shade_tab[shade_list[k]]<<<ctx->grid_shade[shade_list[k]], 128, 0, st>>>(sc->dev, L, P, Q, W.rctr.p);
        sc->trace_k[k_any]<<<sc->trace_grid[k_any], kTraceThreads, smem, st>>>(sc->dev, nullptr, L.sray, L.rad, L.occl, Q.shadow, Q.cnt + 2, 0, scap, Q.cnt + 7, sc->tctr.p, &W.rctr.p->radiance_gt10);
        k_advance<<<1, 32, 0, st>>>(Q, W.rctr.p, W.remaining_dev);
        std::swap(Q.extend, Q.extend_next);
        std::swap(Q.regen, Q.regen_next);
        k_gen<<<g_gen, 128, 0, st>>>(sc->dev, L, P, Q, Q.regen, Q.cnt + 3, W.rctr.p);
      };
      // key: everything the captured launches were given by value
      std::vector<unsigned char> key(sizeof(L) + sizeof(P) + sizeof(Q) + sizeof(sc->dev) + 9 * sizeof(int) + sizeof(size_t));
      {
        unsigned char* kp = key.data();
        memcpy(kp, &L, sizeof(L)); kp += sizeof(L);
        memcpy(kp, &P, sizeof(P)); kp += sizeof(P);
        memcpy(kp, &Q, sizeof(Q)); kp += sizeof(Q);
        memcpy(kp, &sc->dev, sizeof(sc->dev)); kp += sizeof(sc->dev);
        int gk[9] = {sc->trace_grid[k_ext], sc->trace_grid[k_any], g_small, shade_i0 * 16 + (int)sc->class_mask, g_gen, scap, k_ext, k_any, gen_i};
        memcpy(kp, gk, sizeof(gk)); kp += sizeof(gk);
        memcpy(kp, &smem, sizeof(size_t));
      }
      if (!W.graph_exec || W.graph_key != key) {
        if (W.graph_exec) { cudaGraphExecDestroy(W.graph_exec); W.graph_exec = nullptr; }
        cudaGraph_t graph = nullptr;
        GP_CUDA(ctx, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int k = 0; k < kGraphIters; k++) enqueue_iteration();  // kGraphIters is even: the queue pointers end where they started
        // the capture is always ended — a launch error inside it must not leave the context's stream in capture mode
        cudaError_t ce = cudaStreamEndCapture(st, &graph);
        cudaError_t ge = ce == cudaSuccess ? cudaGraphInstantiate(&W.graph_exec, graph, 0) : ce;
        if (graph) cudaGraphDestroy(graph);
        if (ge != cudaSuccess) { W.graph_exec = nullptr; W.graph_key.clear(); cudaGetLastError(); }
        GP_CUDA(ctx, ge);
        W.graph_key = key;
      }
      hw_tick("graph lookup/build");
      for (auto& e : W.graph_ev) if (!e) GP_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      for (uint64_t k = 0;; k++) {
        GP_CUDA(ctx, cudaGraphLaunch(W.graph_exec, st));
        GP_CUDA(ctx, cudaEventRecord(W.graph_ev[k & 1], st));
        ctx->launches += (uint64_t)(5 + n_shade) * kGraphIters;
        iterations += kGraphIters; n_extend += kGraphIters; n_shadow += kGraphIters;
        if (k == 0) continue;
        GP_CUDA(ctx, cudaEventSynchronize(W.graph_ev[(k - 1) & 1]));  // graph k-1 is done, graph k is running or queued
        if (*W.remaining_host == 0) break;
        if (sc->cancel.exchange(0)) { rc = GOPBRT_ERR_CANCELLED; break; }
      }
      GP_CUDA(ctx, cudaStreamSynchronize(st));
    }
    for (; !use_graph;) {
      if (iter_log_path) {
        int c[8];
        cudaMemcpyAsync(c, Q.cnt, sizeof(c), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        iter_counts.push_back(c[0]);
      }
      tick(ST_EXTEND);
      sc->trace_k[k_ext]<<<sc->trace_grid[k_ext], kTraceThreads, smem, st>>>(sc->dev, L.ray, nullptr, L.rad, W.codes.p, Q.extend, Q.cnt + 0, 0, scap, Q.cnt + 6, sc->tctr.p, nullptr);
      tick(ST_SHADE);
      k_split_hits<<<g_small, 256, 0, st>>>(L, Q, W.codes.p);
      for (int k = 0; k < n_shade; k++) shade_tab[shade_list[k]]<<<ctx->grid_shade[shade_list[k]], 128, 0, st>>>(sc->dev, L, P, Q, W.rctr.p);
      if (iter_log_path) {
        int c[12];
        cudaMemcpyAsync(c, Q.cnt, sizeof(c), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        iter_shadow.push_back(c[2]);
        iter_hits.push_back(c[8] + c[9] + c[10] + c[11]);
      }
      tick(ST_SHADOW);
      sc->trace_k[k_any]<<<sc->trace_grid[k_any], kTraceThreads, smem, st>>>(sc->dev, nullptr, L.sray, L.rad, L.occl, Q.shadow, Q.cnt + 2, 0, scap, Q.cnt + 7, sc->tctr.p, &W.rctr.p->radiance_gt10);
      tick(ST_RAYGEN);
      k_advance<<<1, 32, 0, st>>>(Q, W.rctr.p, W.remaining_dev);
      std::swap(Q.extend, Q.extend_next);
      std::swap(Q.regen, Q.regen_next);
      k_gen<<<g_gen, 128, 0, st>>>(sc->dev, L, P, Q, Q.regen, Q.cnt + 3, W.rctr.p);
      ctx->launches += 5 + n_shade;
      iterations++;
      n_extend++; n_shadow++;
      // the host looks at the device-written "lanes still in flight" only every few iterations: an iteration over empty
      // queues costs six near-empty launches, a synchronisation costs a full host round trip
      if ((iterations & 3) != 0) continue;
      GP_CUDA(ctx, cudaStreamSynchronize(st));
      if (*W.remaining_host == 0) break;
      if (sc->cancel.exchange(0)) { rc = GOPBRT_ERR_CANCELLED; break; }
    }
    tick(ST_FILM);
    {
      RenderParams PM = P;
      if (P.uniform_fp && (P.groups > 1 || P.last_in_place) && P.lane_base % P.groups == 0 && P.lanes_active % P.groups == 0) {
        const int g_sums = ctx->sm_count * 16;
        const size_t sm_sums = (size_t)4 * 3 * P.groups * sizeof(double);  // <= 24 KB (groups <= 255)
        const bool single = P.s_world * P.groups >= P.spp - 1;  // at most one sample per lane
        if (P.groups <= 8) {  // few groups: one thread per tile (k_group_sums_small)
          if (!P.last_in_place) k_group_sums_small<false, false, 8><<<g_small, 128, 0, st>>>(L, P, W.rctr.p);
          else if (single) k_group_sums_small<true, true, 8><<<g_small, 128, 0, st>>>(L, P, W.rctr.p);
          else k_group_sums_small<true, false, 8><<<g_small, 128, 0, st>>>(L, P, W.rctr.p);
        }
        else if (!P.last_in_place) k_group_sums<false, false><<<g_sums, 128, sm_sums, st>>>(L, P, W.rctr.p);
        else if (single) k_group_sums<true, true><<<g_sums, 128, sm_sums, st>>>(L, P, W.rctr.p);
        else k_group_sums<true, false><<<g_sums, 128, sm_sums, st>>>(L, P, W.rctr.p);
        ctx->launches++;
        PM.groups_merged = 1;
      } else if (P.last_in_place) {
        k_fold_last<<<g_small, 128, 0, st>>>(L, P, W.rctr.p);
        ctx->launches++;
      }
      k_film_merge<<<g_small, 128, 0, st>>>(L, PM, d_film);
      ctx->launches++;
    }
  }
  tick(ST_N);
  hw_tick("wavefront loop");
  GP_CUDA(ctx, cudaEventRecord(ev[1], st));
  GP_CUDA(ctx, cudaStreamSynchronize(st));
  hw_tick("final sync");
  GP_CUDA(ctx, cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, ev[0], ev[1]);
  RenderCounters rcnt;
  TraceCounters tcnt;
  GP_CUDA(ctx, cudaMemcpy(&rcnt, W.rctr.p, sizeof(rcnt), cudaMemcpyDeviceToHost));
  GP_CUDA(ctx, cudaMemcpy(&tcnt, sc->tctr.p, sizeof(tcnt), cudaMemcpyDeviceToHost));
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->camera_rays = rcnt.camera_rays; stats->closest_rays = rcnt.closest_rays + rcnt.root_culled; stats->shadow_rays = rcnt.shadow_rays;
    stats->dead_mis_rays = rcnt.dead_mis_rays; stats->nodes_visited = tcnt.nodes + (count ? rcnt.root_culled : 0); stats->prim_tests = tcnt.prims;
    stats->shadow_nodes_visited = tcnt.snodes; stats->shadow_prim_tests = tcnt.sprims; stats->radiance_gt10 = rcnt.radiance_gt10;
    stats->nan_samples = rcnt.nan_samples; stats->efloat_panics = rcnt.efloat_panics + tcnt.efloat_panics;
    stats->stack_overflows = tcnt.stack_overflows; stats->iterations = iterations; stats->launches = ctx->launches.load() - launches0;
    stats->lanes = (uint64_t)lanes; stats->ms_total = ms; stats->bvh_nodes = sc->bvh_nodes; stats->bvh_depth = sc->bvh_depth;
    stats->tests_triangle = tcnt.t_tri; stats->tests_sphere_fast = tcnt.t_sph; stats->tests_general = tcnt.t_gen;
    stats->extend_launches = n_extend; stats->shadow_launches = n_shadow;
    stats->root_culled_rays = rcnt.root_culled;
    stats->shaded_lanes = rcnt.shaded;
    stats->shadow_tests_triangle = tcnt.st_tri; stats->shadow_tests_sphere_fast = tcnt.st_sph; stats->shadow_tests_general = tcnt.st_gen;
    if (timing && iter_log_path) {
      FILE* fp = fopen(iter_log_path, "w");
      if (fp) {
        size_t it = 0;
        fprintf(fp, "iter,extend_rays,hits,shadow_rays,ms_extend,ms_shade,ms_shadow,ms_raygen\n");
        double row[ST_N + 1] = {0, 0, 0, 0, 0, 0, 0};
        bool open_row = false;
        for (size_t i = 0; i + 1 < ev_used; i++) {
          float t = 0;
          cudaEventElapsedTime(&t, W.events[i], W.events[i + 1]);
          if (ev_stage[i] == ST_EXTEND) {
            if (open_row) { fprintf(fp, "%zu,%d,%d,%d,%.4f,%.4f,%.4f,%.4f\n", it, it < iter_counts.size() ? iter_counts[it] : -1, it < iter_hits.size() ? iter_hits[it] : -1, it < iter_shadow.size() ? iter_shadow[it] : -1, row[ST_EXTEND], row[ST_SHADE], row[ST_SHADOW], row[ST_RAYGEN]); it++; }
            for (int k = 0; k <= ST_N; k++) row[k] = 0;
            open_row = true;
          }
          if (open_row) row[ev_stage[i]] += t;
        }
        if (open_row) fprintf(fp, "%zu,%d,%d,%d,%.4f,%.4f,%.4f,%.4f\n", it, it < iter_counts.size() ? iter_counts[it] : -1, it < iter_hits.size() ? iter_hits[it] : -1, it < iter_shadow.size() ? iter_shadow[it] : -1, row[ST_EXTEND], row[ST_SHADE], row[ST_SHADOW], row[ST_RAYGEN]);
        fclose(fp);
      }
    }
    if (timing) {
      double acc[ST_N + 1] = {0, 0, 0, 0, 0, 0, 0};
      for (size_t i = 0; i + 1 < ev_used; i++) {
        float t = 0;
        cudaEventElapsedTime(&t, W.events[i], W.events[i + 1]);
        acc[ev_stage[i]] += t;
      }
      stats->ms_raygen = acc[ST_RAYGEN]; stats->ms_extend = acc[ST_EXTEND]; stats->ms_shade = acc[ST_SHADE];
      stats->ms_shadow = acc[ST_SHADOW]; stats->ms_film = acc[ST_FILM];
    }
  }
  hw_tick("stats");
  if (rc == GOPBRT_OK && (flags & GOPBRT_FLAG_FAIL_ON_PANIC) && (rcnt.radiance_gt10 || rcnt.efloat_panics || tcnt.efloat_panics || rcnt.unsupported)) {
    ctx->last_error = "a condition on which the reference panics was hit (see gopbrt_stats)";
    return GOPBRT_ERR_REFERENCE_PANIC;
  }
  return rc;
}

// ------------------------------------------------------------------------------------------------ film reduce (NCCL)
// Sums the ranks' device films onto rank 0 with one ncclReduce on the library stream (in place on the root), and waits for
// it.  Every rank of the communicator must arrive: callers reduce even after a cancelled or failed frame (the sum is then
// meaningless but nobody hangs).  Returns the device time of the reduce through ms (CUDA events on the library stream).
static int film_reduce(gopbrt_ctx* ctx, const gopbrt_render_options* opt, double* d_film, size_t n_doubles, double* ms) {
  if (!ctx->comm) { ctx->last_error = "GOPBRT_FLAG_REDUCE_FILM without a communicator (gopbrt_comm_init_rank / gopbrt_multi_init)"; return GOPBRT_ERR_INVALID; }
  int rank = opt ? opt->rank : 0, world = opt ? opt->world : 1;
  if (rank != ctx->comm_rank || world != ctx->comm_world) { ctx->last_error = "render rank/world differ from the communicator's"; return GOPBRT_ERR_INVALID; }
  std::string err;
  gpcomm::Api* a = gpcomm::api(err);
  if (!a) { ctx->last_error = err; return GOPBRT_ERR_CUDA; }
  cudaEvent_t ev[2];
  GP_CUDA(ctx, cudaEventCreate(&ev[0]));
  GP_CUDA(ctx, cudaEventCreate(&ev[1]));
  cudaEventRecord(ev[0], ctx->stream);
  int r = a->Reduce(d_film, d_film, n_doubles, gpcomm::kNcclFloat64, gpcomm::kNcclSum, 0, ctx->comm, ctx->stream);
  cudaEventRecord(ev[1], ctx->stream);
  cudaError_t ce = cudaStreamSynchronize(ctx->stream);
  float t = 0;
  if (ce == cudaSuccess) cudaEventElapsedTime(&t, ev[0], ev[1]);
  cudaEventDestroy(ev[0]);
  cudaEventDestroy(ev[1]);
  if (r != gpcomm::kNcclSuccess) { ctx->last_error = std::string("ncclReduce: ") + a->GetErrorString(r); return GOPBRT_ERR_CUDA; }
  GP_CUDA(ctx, ce);
  ctx->launches += 1;  // NCCL's reduce kernel
  if (ms) *ms = t;
  return GOPBRT_OK;
}

static size_t film_doubles(const gopbrt_film* film) {
  long long cx0 = (long long)ceil((double)film->width * film->crop[0]), cy0 = (long long)ceil((double)film->height * film->crop[1]);
  long long cx1 = (long long)ceil((double)film->width * film->crop[2]), cy1 = (long long)ceil((double)film->height * film->crop[3]);
  if (cx1 <= cx0 || cy1 <= cy0) return 0;
  return (size_t)(cx1 - cx0) * (cy1 - cy0) * 4;
}

extern "C" int gopbrt_render_device(gopbrt_scene* sc, const gopbrt_camera* cam, const gopbrt_sampler* smp, const gopbrt_integrator* ig,
                                    const gopbrt_film* film, const gopbrt_render_options* opt, double* d_film, gopbrt_stats* stats) {
  if (!sc || !cam || !smp || !ig || !film || !d_film) return GOPBRT_ERR_INVALID;
  gopbrt_ctx* ctx = sc->ctx;
  std::lock_guard<std::mutex> g(sc->mu);
  std::lock_guard<std::mutex> grun(sc->ctx->run_mu);
  GP_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = render_impl(sc, cam, smp, ig, film, opt, d_film, stats);
  if (opt && (opt->flags & GOPBRT_FLAG_REDUCE_FILM) && rc != GOPBRT_ERR_CUDA && rc != GOPBRT_ERR_INVALID) {
    double ms = 0;
    int rr = film_reduce(ctx, opt, d_film, film_doubles(film), &ms);
    if (stats) stats->ms_reduce = ms;
    if (rc == GOPBRT_OK) rc = rr;
  }
  return rc;
}

extern "C" int gopbrt_render(gopbrt_scene* sc, const gopbrt_camera* cam, const gopbrt_sampler* smp, const gopbrt_integrator* ig,
                             const gopbrt_film* film, const gopbrt_render_options* opt, double* film_out, gopbrt_stats* stats) {
  if (!sc || !cam || !smp || !ig || !film) return GOPBRT_ERR_INVALID;
  const bool reduce = opt && (opt->flags & GOPBRT_FLAG_REDUCE_FILM);
  // the summed film exists on rank 0 only: the other ranks of a reduced frame pass no host buffer
  if (!film_out && !(reduce && opt->rank != 0)) return GOPBRT_ERR_INVALID;
  gopbrt_ctx* ctx = sc->ctx;
  std::lock_guard<std::mutex> g(sc->mu);
  std::lock_guard<std::mutex> grun(sc->ctx->run_mu);
  GP_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t n = film_doubles(film);
  if (n == 0) { ctx->last_error = "empty film"; return GOPBRT_ERR_INVALID; }
  DevBuf<double>& d_film = sc->ws.film;
  if (d_film.n != n) { d_film.release(); GP_CUDA(ctx, d_film.alloc(n)); }
  int rc = render_impl(sc, cam, smp, ig, film, opt, d_film.p, stats);
  if (reduce && rc != GOPBRT_ERR_CUDA && rc != GOPBRT_ERR_INVALID) {
    double ms = 0;
    int rr = film_reduce(ctx, opt, d_film.p, n, &ms);
    if (stats) stats->ms_reduce = ms;
    if (rc == GOPBRT_OK) rc = rr;
  }
  if (rc != GOPBRT_OK) return rc;
  if (film_out && (!reduce || opt->rank == 0)) {
    auto t0 = std::chrono::steady_clock::now();
    GP_CUDA(ctx, cudaMemcpy(film_out, d_film.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (stats) stats->ms_download = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  return GOPBRT_OK;
}

// ------------------------------------------------------------------------------------------------ communicators
extern "C" int gopbrt_comm_unique_id(unsigned char id[GOPBRT_COMM_ID_BYTES]) {
  if (!id) return GOPBRT_ERR_INVALID;
  std::string err;
  gpcomm::Api* a = gpcomm::api(err);
  if (!a) return GOPBRT_ERR_CUDA;
  gpcomm::ncclUniqueId u;
  static_assert(sizeof(u) == GOPBRT_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  if (a->GetUniqueId(&u) != gpcomm::kNcclSuccess) return GOPBRT_ERR_CUDA;
  memcpy(id, &u, sizeof(u));
  return GOPBRT_OK;
}

extern "C" int gopbrt_comm_init_rank(gopbrt_ctx* ctx, const unsigned char id[GOPBRT_COMM_ID_BYTES], int rank, int world) {
  if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return GOPBRT_ERR_INVALID;
  std::lock_guard<std::mutex> g(ctx->mu);
  std::lock_guard<std::mutex> grun(ctx->run_mu);
  if (ctx->comm) { ctx->last_error = "the context already has a communicator"; return GOPBRT_ERR_INVALID; }
  std::string err;
  gpcomm::Api* a = gpcomm::api(err);
  if (!a) { ctx->last_error = err; return GOPBRT_ERR_CUDA; }
  GP_CUDA(ctx, cudaSetDevice(ctx->device));
  gpcomm::ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  int r = a->CommInitRank(&ctx->comm, world, u, rank);
  if (r != gpcomm::kNcclSuccess) { ctx->comm = nullptr; ctx->last_error = std::string("ncclCommInitRank: ") + a->GetErrorString(r); return GOPBRT_ERR_CUDA; }
  ctx->comm_rank = rank; ctx->comm_world = world;
  return GOPBRT_OK;
}

// ------------------------------------------------------------------------------------------------ one process, N GPUs
// pbrt.Render is ONE call from ONE process (integrator.go:291-350, called at internal/render/server.go:164).  gopbrt_multi
// is that call over N devices: the scene replicated per device, the frame's samples (FAST) or tiles (STRICT) split by rank,
// one host thread per device driving its wavefront, one ncclReduce of the films onto device 0, the host film read from there.
struct gopbrt_multi {
  std::vector<gopbrt_ctx*> ctx;
  std::string last_error;
  std::mutex mu;
};
struct gopbrt_multi_scene {
  gopbrt_multi* m = nullptr;
  std::vector<gopbrt_scene*> sc;
};

extern "C" void gopbrt_multi_shutdown(gopbrt_multi* m) {
  if (!m) return;
  for (gopbrt_ctx* c : m->ctx) gopbrt_shutdown(c);
  delete m;
}

extern "C" int gopbrt_multi_init(int n_gpus, const int* devices, gopbrt_multi** out) {
  if (!out || n_gpus < 1 || n_gpus > 64) return GOPBRT_ERR_INVALID;
  *out = nullptr;
  gopbrt_multi* m = new gopbrt_multi();
  std::vector<int> devs(n_gpus);
  for (int i = 0; i < n_gpus; i++) {
    devs[i] = devices ? devices[i] : i;
    gopbrt_ctx* c = nullptr;
    int rc = gopbrt_init(devs[i], &c);
    if (rc != GOPBRT_OK) { gopbrt_multi_shutdown(m); return rc; }
    m->ctx.push_back(c);
  }
  if (n_gpus > 1) {
    std::string err;
    gpcomm::Api* a = gpcomm::api(err);
    if (!a) { gopbrt_multi_shutdown(m); return GOPBRT_ERR_CUDA; }
    std::vector<gpcomm::ncclComm_t> comms(n_gpus, nullptr);
    if (a->CommInitAll(comms.data(), n_gpus, devs.data()) != gpcomm::kNcclSuccess) { gopbrt_multi_shutdown(m); return GOPBRT_ERR_CUDA; }
    for (int i = 0; i < n_gpus; i++) { m->ctx[i]->comm = comms[i]; m->ctx[i]->comm_rank = i; m->ctx[i]->comm_world = n_gpus; }
  }
  *out = m;
  return GOPBRT_OK;
}

extern "C" int gopbrt_multi_device_count(const gopbrt_multi* m) { return m ? (int)m->ctx.size() : 0; }
extern "C" const char* gopbrt_multi_last_error(const gopbrt_multi* m) { return m ? m->last_error.c_str() : "null multi context"; }
extern "C" uint64_t gopbrt_multi_launch_count(const gopbrt_multi* m) {
  uint64_t n = 0;
  if (m) for (gopbrt_ctx* c : m->ctx) n += c->launches.load();
  return n;
}

extern "C" void gopbrt_multi_scene_destroy(gopbrt_multi_scene* ms) {
  if (!ms) return;
  for (gopbrt_scene* s : ms->sc) gopbrt_scene_destroy(s);
  delete ms;
}

extern "C" int gopbrt_multi_scene_create(gopbrt_multi* m, const gopbrt_scene_desc* d, gopbrt_multi_scene** out) {
  if (!m || !d || !out) return GOPBRT_ERR_INVALID;
  *out = nullptr;
  HostScene H;  // validated, bounded and BVH-built once on the host; every device receives the same bytes
  std::string err;
  int rc = build_host_scene(d, H, err);
  if (rc != GOPBRT_OK) { std::lock_guard<std::mutex> g(m->mu); m->last_error = err; return rc; }
  const int n = (int)m->ctx.size();
  gopbrt_multi_scene* ms = new gopbrt_multi_scene();
  ms->m = m;
  ms->sc.assign(n, nullptr);
  std::vector<int> rcs(n, GOPBRT_OK);
  std::vector<std::thread> th;
  for (int i = 0; i < n; i++) th.emplace_back([&, i]() { rcs[i] = upload_scene(m->ctx[i], H, &ms->sc[i]); });
  for (auto& t : th) t.join();
  for (int i = 0; i < n; i++)
    if (rcs[i] != GOPBRT_OK) {
      { std::lock_guard<std::mutex> g(m->mu); m->last_error = "device " + std::to_string(m->ctx[i]->device) + ": " + m->ctx[i]->last_error; }
      rc = rcs[i];
      gopbrt_multi_scene_destroy(ms);
      return rc;
    }
  *out = ms;
  return GOPBRT_OK;
}

extern "C" int gopbrt_multi_cancel(gopbrt_multi_scene* ms) {
  if (!ms) return GOPBRT_ERR_INVALID;
  for (gopbrt_scene* s : ms->sc) gopbrt_cancel(s);
  return GOPBRT_OK;
}

extern "C" int gopbrt_multi_render(gopbrt_multi_scene* ms, const gopbrt_camera* cam, const gopbrt_sampler* smp, const gopbrt_integrator* ig,
                                   const gopbrt_film* film, int flags, double* film_out, gopbrt_stats* stats_out) {
  if (!ms || !cam || !smp || !ig || !film || !film_out) return GOPBRT_ERR_INVALID;
  const int n = (int)ms->sc.size();
  std::vector<int> rcs(n, GOPBRT_OK);
  std::vector<gopbrt_stats> st(n);
  auto one = [&](int i) {
    gopbrt_render_options o;
    o.rank = i; o.world = n; o.max_lanes = 0;
    o.flags = (flags & ~GOPBRT_FLAG_REDUCE_FILM) | (n > 1 ? GOPBRT_FLAG_REDUCE_FILM : 0);
    rcs[i] = gopbrt_render(ms->sc[i], cam, smp, ig, film, &o, i == 0 ? film_out : nullptr, &st[i]);
  };
  std::vector<std::thread> th;
  for (int i = 1; i < n; i++) th.emplace_back(one, i);
  one(0);
  for (auto& t : th) t.join();
  int rc = GOPBRT_OK;
  for (int i = 0; i < n; i++)
    if (rcs[i] != GOPBRT_OK) {
      std::lock_guard<std::mutex> g(ms->m->mu);
      ms->m->last_error = "device " + std::to_string(ms->m->ctx[i]->device) + ": " + ms->m->ctx[i]->last_error;
      if (rc == GOPBRT_OK) rc = rcs[i];
    }
  if (stats_out) {  // counters summed over the devices, times = the slowest device (the frame's critical path)
    gopbrt_stats t = st[0];
    for (int i = 1; i < n; i++) {
      const gopbrt_stats& s = st[i];
      t.camera_rays += s.camera_rays; t.closest_rays += s.closest_rays; t.shadow_rays += s.shadow_rays; t.dead_mis_rays += s.dead_mis_rays;
      t.nodes_visited += s.nodes_visited; t.prim_tests += s.prim_tests; t.shadow_nodes_visited += s.shadow_nodes_visited;
      t.shadow_prim_tests += s.shadow_prim_tests; t.radiance_gt10 += s.radiance_gt10; t.nan_samples += s.nan_samples;
      t.efloat_panics += s.efloat_panics; t.stack_overflows += s.stack_overflows; t.launches += s.launches; t.lanes += s.lanes;
      t.tests_triangle += s.tests_triangle; t.tests_sphere_fast += s.tests_sphere_fast; t.tests_general += s.tests_general;
      t.extend_launches += s.extend_launches; t.shadow_launches += s.shadow_launches; t.root_culled_rays += s.root_culled_rays; t.shaded_lanes += s.shaded_lanes;
      t.shadow_tests_triangle += s.shadow_tests_triangle; t.shadow_tests_sphere_fast += s.shadow_tests_sphere_fast;
      t.shadow_tests_general += s.shadow_tests_general;
      t.iterations = std::max(t.iterations, s.iterations);
      t.ms_total = std::max(t.ms_total, s.ms_total); t.ms_raygen = std::max(t.ms_raygen, s.ms_raygen); t.ms_extend = std::max(t.ms_extend, s.ms_extend);
      t.ms_shade = std::max(t.ms_shade, s.ms_shade); t.ms_shadow = std::max(t.ms_shadow, s.ms_shadow); t.ms_film = std::max(t.ms_film, s.ms_film);
      t.ms_reduce = std::max(t.ms_reduce, s.ms_reduce);
    }
    *stats_out = t;
  }
  return rc;
}

// ------------------------------------------------------------------------------------------------ self-test hook
// Evaluates one device function of the raygen / shade / film stages on the GPU (gp_kat.cuh).  Returns the number of
// doubles written to `out`, or a negative GOPBRT_ERR_* code.
extern "C" int gopbrt_kat_eval(gopbrt_scene* sc, int fn, const double* in, int n_in, double* out, int n_out) {
  if (!sc || !in || !out || n_in < 0 || n_in > 64 || n_out < 1 || n_out > 4096 || fn < 0 || fn >= KAT_N) return -GOPBRT_ERR_INVALID;
  gopbrt_ctx* ctx = sc->ctx;
  std::lock_guard<std::mutex> g(sc->mu);
  std::lock_guard<std::mutex> grun(ctx->run_mu);
  auto cuda_fail = [&](cudaError_t e) { ctx->last_error = std::string("gopbrt_kat_eval: ") + cudaGetErrorString(e); return -GOPBRT_ERR_CUDA; };
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess) return cuda_fail(e);
  RenderParams P;
  memset(&P, 0, sizeof(P));
  std::vector<double> scratch(1, 0.0);
  if (fn == KAT_SAMPLE_DISCRETE_UNIFORM && n_in == 2) {
    int nl = (int)in[0];
    if (nl < 1 || nl > 1024) return -GOPBRT_ERR_INVALID;
    std::vector<double> cdf;
    double fi = uniform_light_cdf(nl, cdf);
    scratch.assign(1, fi);
    scratch.insert(scratch.end(), cdf.begin(), cdf.end());
  } else if (fn == KAT_FILM_ADD_SAMPLE && n_in == 11) {
    gopbrt_film f;
    f.width = (int)in[0]; f.height = (int)in[1];
    f.crop[0] = 0; f.crop[1] = 0; f.crop[2] = 1; f.crop[3] = 1;
    f.filter_radius[0] = in[3]; f.filter_radius[1] = in[4];
    if (const char* msg = film_params(&f, (int64_t)in[2], P)) { ctx->last_error = msg; return -GOPBRT_ERR_INVALID; }
    P.world = 1; P.groups = 1; P.s_world = 1;
    if (in[5] < 0 || in[5] >= (double)P.ntiles) return -GOPBRT_ERR_INVALID;
    scratch.assign((size_t)P.tpw * P.tph * 4, 0.0);
  } else if (fn == KAT_STRATIFIED_START_PIXEL && n_in == 5) {
    P.sampler_kind = GOPBRT_SAMPLER_STRATIFIED; P.mode = GOPBRT_MODE_STRICT;
    P.xs = (int)in[1]; P.ys = (int)in[2]; P.jitter = (int)in[3]; P.ndims = (int)in[4];
    P.spp = P.xs * P.ys;
    if (P.spp < 1 || P.spp > 4096 || P.ndims < 0 || P.ndims > 16) return -GOPBRT_ERR_INVALID;
    scratch.assign((size_t)std::max(1, P.ndims * P.spp), 0.0);
  } else if (fn == KAT_CAMERA_RAY && n_in == 38) {
    P.raster_to_camera = m4_from(in); P.camera_to_world = m4_from(in + 16);
    P.lens_radius = in[32]; P.focal_distance = in[33];
  }
  DevBuf<double> d_in, d_out, d_scratch;
  DevBuf<int> d_n;
  cudaStream_t st = ctx->stream;
  if ((e = d_in.alloc(std::max(1, n_in))) != cudaSuccess || (e = d_out.alloc(n_out)) != cudaSuccess || (e = d_scratch.alloc(scratch.size())) != cudaSuccess ||
      (e = d_n.alloc(1)) != cudaSuccess)
    return cuda_fail(e);
  if (n_in && (e = cudaMemcpyAsync(d_in.p, in, n_in * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(e);
  if ((e = cudaMemcpyAsync(d_scratch.p, scratch.data(), scratch.size() * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(e);
  if ((e = cudaMemsetAsync(d_out.p, 0, n_out * sizeof(double), st)) != cudaSuccess) return cuda_fail(e);
  k_kat<<<1, 32, 0, st>>>(sc->dev, fn, d_in.p, n_in, d_out.p, n_out, d_scratch.p, P, d_n.p);
  ctx->launches++;
  int n = -1;
  if ((e = cudaMemcpyAsync(&n, d_n.p, sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(e);
  if ((e = cudaMemcpyAsync(out, d_out.p, n_out * sizeof(double), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(e);
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail(e);
  if (n < 0) { ctx->last_error = "gopbrt_kat_eval: bad argument count for this function"; return -GOPBRT_ERR_INVALID; }
  return n;
}
