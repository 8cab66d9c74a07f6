/*
 * gopbrt_cuda.h — C ABI of libgopbrt_cuda.so, the B200 (sm_100a) backend for the
 * ray-scene intersection + path-integrator hot path of ssttuu/go-pbrt.
 *
 * The reference has no FFI today; every entry point below names the Go
 * interface/function it stands in for (paths relative to the reference repo root).
 * All structs are plain-old-data, all arrays are caller-owned and copied inside the call
 * (no pointer is retained after a call returns — the cgo rule).  All arithmetic types are
 * float64 ("double") because the reference is float64 end to end (pkg/pbrt/geometry.go:9-22).
 *
 * There is NO CPU fallback: every entry point fails with GOPBRT_ERR_CUDA if no sm_100
 * device is usable.
 */
#ifndef GOPBRT_CUDA_H
#define GOPBRT_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GOPBRT_ABI_VERSION 1

/* ---- status codes (Go side wraps them with errors.Wrap like internal/render/server.go:165-167) ---- */
enum {
  GOPBRT_OK = 0,
  GOPBRT_ERR_INVALID = 1,    /* bad argument / inconsistent scene description            */
  GOPBRT_ERR_CUDA = 2,       /* CUDA failure or no sm_100 device: hard error, no fallback */
  GOPBRT_ERR_CANCELLED = 3,  /* gopbrt_cancel() observed (ctx.Done(), integrator.go:332-336) */
  GOPBRT_ERR_REFERENCE_PANIC = 4, /* a condition on which the reference would panic was hit and
                                     GOPBRT_FLAG_FAIL_ON_PANIC was set (see gopbrt_stats)     */
  GOPBRT_ERR_UNSUPPORTED = 5
};

typedef struct gopbrt_ctx gopbrt_ctx;     /* one per process+device                         */
typedef struct gopbrt_scene gopbrt_scene; /* immutable after create; shareable across threads */
/* Threading: every entry point may be called from any host thread (each gRPC request of the reference renders on its own
 * goroutine with its own scene, internal/render/server.go).  Device work of the scene handles of ONE gopbrt_ctx is
 * serialised inside the library (one stream, CUDA-graph capture); gopbrt_cancel never blocks. */

/* pbrt.Transform {Matrix, MatrixInverse} (pkg/pbrt/transform.go:144-146), row-major, passed
 * verbatim: the library never recomputes an inverse (the host's may be "wrong"). */
typedef struct {
  double m[16];
  double minv[16];
} gopbrt_transform;

enum { GOPBRT_SHAPE_SPHERE = 0, GOPBRT_SHAPE_DISK = 1, GOPBRT_SHAPE_TRIANGLE = 2 };

/* pbrt.NewSphere arguments (pkg/pbrt/sphere.go:19-32); worldToObject = objectToWorld.Inverse()
 * i.e. the swapped pair (sphere.go:34-36, transform.go:175-177). */
typedef struct {
  int32_t object_to_world; /* index into transforms */
  int32_t reverse_orientation;
  double radius, z_min, z_max, phi_max_deg;
} gopbrt_sphere;

/* shapes.NewDisk arguments (pkg/shapes/disk.go:22-35); reverseOrientation is hard-wired false there. */
typedef struct {
  int32_t object_to_world;
  int32_t reverse_orientation;
  double height, radius, inner_radius, phi_max_deg;
} gopbrt_disk;

/* Triangle: NOT in the reference — a new pbrt.Shape (pkg/pbrt/shape.go:9-22) defined
 * by this backend and its oracle.  Vertices are world-space float64, indices into `vertices`. */
typedef struct {
  int32_t v[3];
  int32_t reverse_orientation;
} gopbrt_triangle;

/* pbrt.GeometricPrimitive (pkg/pbrt/primitive.go:22-36) optionally wrapped in a
 * pbrt.TransformedPrimitive with a static AnimatedTransform (primitive.go:82-115). */
typedef struct {
  int32_t shape_kind;
  int32_t shape_index;
  int32_t material;         /* index into materials; -1 = nil material (reference panics when shaded) */
  int32_t prim_to_world;    /* index into transforms, or -1 for a bare GeometricPrimitive             */
} gopbrt_primitive;

enum { GOPBRT_MAT_MATTE = 0, GOPBRT_MAT_MIRROR = 1, GOPBRT_MAT_GLASS = 2 };

/* materials.MatteMaterial (matte.go:8-37), materials.Mirror (mirror.go:9-32), materials.Glass (glass.go:7-75).
 * Float textures are constants (the reference has only ConstantFloatTexture, texture.go:70-82). */
typedef struct {
  int32_t kind;
  int32_t tex_a;  /* matte: Kd; mirror: Kr; glass: Kr */
  int32_t tex_b;  /* glass: Kt; else -1 */
  int32_t pad;
  double sigma;   /* matte */
  double eta;     /* glass index */
  double u_rough, v_rough; /* glass; non-zero => microfacet branch, unsupported (reference panics) */
} gopbrt_material;

enum { GOPBRT_TEX_CONSTANT = 0, GOPBRT_TEX_CHECKERBOARD = 1 };
enum { GOPBRT_MAP_UV = 0, GOPBRT_MAP_PLANAR = 1 };

/* pbrt.ConstantSpectrumTexture (texture.go:56-68) / textures.Checkerboard2D (checkerboard.go:14-40) with
 * pbrt.UVMapping2D (texture.go:9-26) or pbrt.PlanarMapping2D (texture.go:28-46). */
typedef struct {
  int32_t kind;
  int32_t mapping;
  int32_t tex1, tex2;   /* checkerboard children */
  double rgb[3];        /* constant value */
  double vs[3], vt[3];  /* planar mapping */
  double ds, dt;        /* planar mapping */
  double su, sv, du, dv;/* uv mapping */
} gopbrt_texture;

enum { GOPBRT_LIGHT_DISTANT = 0, GOPBRT_LIGHT_POINT = 1, GOPBRT_LIGHT_DIFFUSE_AREA = 2 };

/* lights.Distant (distant.go:8-31): v = wLight as held by the struct (already transformed+normalised);
 * lights.Point (point.go:8-31): v = pLight;  lights.DiffuseAreaLight (diffuse.go:8-24): shape + LEmit. */
typedef struct {
  int32_t kind;
  int32_t shape_kind;   /* area light: GOPBRT_SHAPE_SPHERE or GOPBRT_SHAPE_DISK */
  int32_t shape_index;  /* index into spheres/disks (the shape need not be referenced by a primitive) */
  int32_t two_sided;
  double rgb[3];        /* L / I / LEmit */
  double v[3];
} gopbrt_light;

/* Everything accelerator.NewBVH + pbrt.NewScene receive (internal/render/server.go:104-132). */
typedef struct {
  int32_t n_transforms;  const gopbrt_transform* transforms;
  int32_t n_spheres;     const gopbrt_sphere* spheres;
  int32_t n_disks;       const gopbrt_disk* disks;
  int64_t n_vertices;    const double* vertices;      /* xyz triples */
  int64_t n_triangles;   const gopbrt_triangle* triangles;
  int64_t n_primitives;  const gopbrt_primitive* primitives;
  int32_t n_materials;   const gopbrt_material* materials;
  int32_t n_textures;    const gopbrt_texture* textures;
  int32_t n_lights;      const gopbrt_light* lights;
  int32_t max_prims_in_node; /* NewBVH's maxPrimsInNode (bvh.go:223), an upper bound on a leaf: the library's own tree
                                uses min(this, 2) primitives per leaf (results do not depend on the tree); 0 = default */
  int32_t flags;
} gopbrt_scene_desc;

/* pbrt.PerspectiveCamera as constructed (camera.go:106-165): the two matrices the hot path reads
 * (RasterToCamera.Matrix, cameraToWorld.startTransform.Matrix) plus lens and shutter as the struct
 * holds them (note camera.go:116 stores shutterClose = shutterOpen). */
typedef struct {
  double raster_to_camera[16];
  double camera_to_world[16];
  double lens_radius, focal_distance;
  double shutter_open, shutter_close;
} gopbrt_camera;

enum { GOPBRT_SAMPLER_STRATIFIED = 0, GOPBRT_SAMPLER_RANDOM = 1 };
/* STRICT: the reference's RNG stream per tile, Clone(seed = tileY*nTilesX + tileX)
 * (pkg/pbrt/integrator.go:318-328, pkg/sampler/pixel.go:34-42), sequential over the tile's pixels and
 * samples: identical per-pixel sample sequences to pbrt.Render(…, tileSize).
 * FAST: counter-based stream per (pixel, sample) — needs the matching Go Sampler on the host. */
enum { GOPBRT_MODE_STRICT = 0, GOPBRT_MODE_FAST = 1 };

/* sampler.NewStratified(xSamples, ySamples, jitter, nSampledDimensions) (stratified.go:13-20) or
 * sampler.NewRandomSampler(ns, seed) (random.go:12-19; x_samples = ns, y_samples = 1). */
typedef struct {
  int32_t kind;
  int32_t x_samples, y_samples;
  int32_t jitter;
  int32_t n_sampled_dimensions;
  int32_t mode;
} gopbrt_sampler;

enum {
  GOPBRT_INTEGRATOR_PATH = 0,            /* integrator.Path (pkg/integrator/path.go:32-157)                              */
  GOPBRT_INTEGRATOR_DIRECT_LIGHTING = 1  /* integrator.DirectLighting (pkg/integrator/directlighting.go:62-104)          */
};
/* light_strategy: Path -> the LightSampleStrategy of lightdistribution.go:5-9 (Uniform = 1, Power = 2, Spatial = 4).
 *   Power is reproduced bugs included (lightdistribution.go:57-68 appends to a zero-filled slice and Spectrum.Y() is 0, so
 *   no light is ever sampled); Spatial has no distribution in the reference (nil, path.go:80 panics) -> GOPBRT_ERR_UNSUPPORTED.
 * DirectLighting -> its LightStrategy (directlighting.go:12-15): 1 = UniformSampleAll, 2 = UniformSampleOne */
enum { GOPBRT_LIGHTS_UNIFORM = 1, GOPBRT_LIGHTS_POWER = 2, GOPBRT_LIGHTS_SPATIAL = 4 };
enum { GOPBRT_DL_SAMPLE_ALL = 1, GOPBRT_DL_SAMPLE_ONE = 2 };

/* integrator.NewPath(maxDepth, camera, sampler, pixelBounds, rrThreshold, strategy) (path.go:10-18) or
 * integrator.NewDirectLighting(strategy, maxDepth, camera, sampler, pixelBounds) (directlighting.go:27-35; rr_threshold
 * unused), and the tileSize argument of pbrt.Render (integrator.go:291). */
typedef struct {
  int32_t kind;
  int32_t max_depth;
  double rr_threshold;
  int32_t light_strategy;
  int32_t pad;
  int64_t tile_size;
} gopbrt_integrator;

/* pbrt.NewFilm(resolution, cropWindow, BoxFilter(radius), …) (film.go:43-76). */
typedef struct {
  int32_t width, height;
  double crop[4];          /* min.x, min.y, max.x, max.y */
  double filter_radius[2]; /* BoxFilter radius (filter.go:20-32) */
} gopbrt_film;

enum {
  GOPBRT_FLAG_COUNT_TRAVERSAL = 1, /* instrumented kernels: count BVH nodes visited / primitive tests */
  GOPBRT_FLAG_FAIL_ON_PANIC = 2,   /* return GOPBRT_ERR_REFERENCE_PANIC instead of counting           */
  GOPBRT_FLAG_TIME_KERNELS = 4,    /* CUDA-event time every stage launch (fills ms_raygen … ms_film)   */
  /* 8: reserved (was GOPBRT_FLAG_TAIL in ABI 1 drafts; ignored) */
  GOPBRT_FLAG_REDUCE_FILM = 16,    /* world > 1: after the last wavefront the ranks' films are summed onto rank 0 with ONE
                                      ncclReduce (float64, sum) on the library stream — the GPU form of the reference merging
                                      every worker's FilmTile into one Film (film.go:115-132).  Needs a communicator on the
                                      context (gopbrt_comm_init_rank, or gopbrt_multi_init) whose rank/world equal the
                                      render options'.  Every rank must make the call. */
  /* bits 8..15: GOPBRT_MODE_FAST only — lane groups per pixel tile (each group renders every n-th sample of the
   * rank's share, into its own FilmTile; the groups are merged in ascending order).  0 = automatic. */
  GOPBRT_FLAG_GROUPS_SHIFT = 8,
  GOPBRT_FLAG_GROUPS_MASK = 0xff00
};

/* work partition for one process per GPU: this rank renders tiles t with t % world == rank
 * (strict) or samples s with s % world == rank (fast).  world = 1 for a single GPU. */
typedef struct {
  int32_t rank, world;
  int32_t flags;
  int32_t max_lanes;  /* 0 = default; upper bound on paths in flight */
} gopbrt_render_options;

typedef struct {
  uint64_t camera_rays;        /* paths started = W*H*(spp-1) (sampler.go:29-34)               */
  uint64_t closest_rays;       /* scene.Intersect queries (path.go:45)                         */
  uint64_t shadow_rays;        /* scene.IntersectP queries (light.go:46-48)                    */
  uint64_t dead_mis_rays;      /* EstimateDirect's always-discarded ray (integrator.go:165-183): counted, not traced */
  uint64_t nodes_visited, prim_tests;               /* closest-hit, with COUNT_TRAVERSAL (prim_tests = shape tests run) */
  uint64_t shadow_nodes_visited, shadow_prim_tests; /* any-hit, with COUNT_TRAVERSAL           */
  uint64_t radiance_gt10;      /* UniformSampleOneLight panic condition (integrator.go:73-75)  */
  uint64_t nan_samples;        /* L.HasNaNs() → 0.1 grey (integrator.go:256-257)               */
  uint64_t efloat_panics;      /* efloat.Check failures (efloat.go:102-111)                    */
  uint64_t stack_overflows;    /* traversal stack exhausted (never with the library's own BVH) */
  uint64_t iterations, launches;
  uint64_t lanes;
  double ms_total, ms_raygen, ms_extend, ms_shade, ms_shadow, ms_film, ms_download;
  uint64_t bvh_nodes, bvh_depth;
  uint64_t tests_triangle, tests_sphere_fast, tests_general; /* closest-hit shape tests by record kind (COUNT_TRAVERSAL) */
  uint64_t extend_launches, shadow_launches;
  uint64_t shadow_tests_triangle, shadow_tests_sphere_fast, shadow_tests_general; /* any-hit, same split */
  uint64_t shaded_lanes;       /* lanes the shade stage worked on (closest hits that reached a Path.Li / DirectLighting.Li body) */
  double ms_reduce;            /* GOPBRT_FLAG_REDUCE_FILM: device time of the NCCL film reduce (CUDA events on the library stream) */
  uint64_t root_culled_rays;   /* closest_rays answered by the BVH-root slab test inside raygen (never reach the extend kernel) */
} gopbrt_stats;

/* ---- lifecycle ---- */
int gopbrt_abi_version(void);
/* device = CUDA ordinal (LOCAL_RANK).  Fails if the device is not sm_100. */
int gopbrt_init(int device, gopbrt_ctx** out);
void gopbrt_shutdown(gopbrt_ctx*);
const char* gopbrt_last_error(const gopbrt_ctx*);

/* == accelerator.NewBVH(prims, maxPrims, split) + pbrt.NewScene(agg, lights) (bvh.go:223, scene.go:16-36):
 * copies everything, builds + flattens the BVH, uploads once. */
int gopbrt_scene_create(gopbrt_ctx*, const gopbrt_scene_desc*, gopbrt_scene** out);
void gopbrt_scene_destroy(gopbrt_scene*);
/* root bound of the aggregate == Aggregate.WorldBound() (bvh.go:653-658): min xyz, max xyz */
int gopbrt_scene_world_bound(const gopbrt_scene*, double out6[6]);

/* == Aggregate.Intersect (bvh.go:659-712), batched over n rays, SoA float64 host arrays.
 * prim: index into desc.primitives or -1; t: tHit (r.TMax after the call); p, n: world-space hit
 * point and geometric normal, xyz triples (may be NULL). */
int gopbrt_trace_closest(gopbrt_scene*, int64_t n,
                         const double* ox, const double* oy, const double* oz,
                         const double* dx, const double* dy, const double* dz,
                         const double* tmax,
                         int32_t* prim, double* t, double* p, double* nrm);
/* == Aggregate.IntersectP (bvh.go:713-765) */
int gopbrt_trace_any(gopbrt_scene*, int64_t n,
                     const double* ox, const double* oy, const double* oz,
                     const double* dx, const double* dy, const double* dz,
                     const double* tmax, uint8_t* hit);

/* device-resident variants (all pointers are device pointers on the ctx's device, work is enqueued on
 * `stream` (a cudaStream_t passed as void*) and NOT synchronised) */
int gopbrt_trace_closest_device(gopbrt_scene*, int64_t n, const double* rays_soa7 /* ox,oy,oz,dx,dy,dz,tmax planes of n */,
                                int32_t* prim, double* t, void* stream);
int gopbrt_trace_any_device(gopbrt_scene*, int64_t n, const double* rays_soa7, uint8_t* hit, void* stream);

/* == pbrt.Render(ctx, integrator, scene, tileSize) (integrator.go:291-350) up to, not including,
 * Film.WriteImage: film_out receives W'*H'*4 float64 = Pixel{value[3] (XYZ sums), filterWeightSum}
 * (film.go:20-25) row-major over CroppedPixelBounds. */
int gopbrt_render(gopbrt_scene*, const gopbrt_camera*, const gopbrt_sampler*, const gopbrt_integrator*,
                  const gopbrt_film*, const gopbrt_render_options*, double* film_out, gopbrt_stats* stats_out);
/* same, film left in device memory (d_film: device pointer to W'*H'*4 doubles, zeroed by the call) */
int gopbrt_render_device(gopbrt_scene*, const gopbrt_camera*, const gopbrt_sampler*, const gopbrt_integrator*,
                         const gopbrt_film*, const gopbrt_render_options*, double* d_film, gopbrt_stats* stats_out);
/* thread-safe; makes a running gopbrt_render return GOPBRT_ERR_CANCELLED between wavefront iterations */
int gopbrt_cancel(gopbrt_scene*);

/* ---- multi-GPU: the frame's samples (FAST) or tiles (STRICT) split by rank, films summed with one NCCL reduce ----
 * The reference renders a frame with ONE call from ONE process (pbrt.Render, integrator.go:291-350, called at
 * internal/render/server.go:164; its workers are goroutines, integrator.go:304-312).  Two bindings of that shape:
 *
 * (1) one process per GPU (torchrun / MPI style): rank 0 calls gopbrt_comm_unique_id and hands the 128 bytes to every rank by
 *     any host channel; every rank calls gopbrt_comm_init_rank on its own context; gopbrt_render / gopbrt_render_device with
 *     {rank, world, GOPBRT_FLAG_REDUCE_FILM} then leave the summed film on rank 0 (host film_out may be NULL on ranks != 0).
 * (2) one process, N GPUs (the Go daemon): gopbrt_multi_init creates one context per device and their communicators
 *     (ncclCommInitAll); gopbrt_multi_scene_create builds the BVH once on the host and uploads it to every device;
 *     gopbrt_multi_render is pbrt.Render: one host thread per device drives that device's wavefront, the films meet in one
 *     ncclReduce on device 0 and film_out is read from there.  stats_out: counters summed over devices, times = max.
 * libnccl.so.2 is bound at run time (dlopen; GOPBRT_NCCL_LIB overrides) — single-GPU hosts do not need it. */
#define GOPBRT_COMM_ID_BYTES 128
int gopbrt_comm_unique_id(unsigned char id[GOPBRT_COMM_ID_BYTES]);
int gopbrt_comm_init_rank(gopbrt_ctx*, const unsigned char id[GOPBRT_COMM_ID_BYTES], int rank, int world);

typedef struct gopbrt_multi gopbrt_multi;             /* N contexts + their communicators, one process */
typedef struct gopbrt_multi_scene gopbrt_multi_scene; /* one scene replicated on the N devices         */
/* devices: n_gpus CUDA ordinals, or NULL for 0..n_gpus-1 */
int gopbrt_multi_init(int n_gpus, const int* devices, gopbrt_multi** out);
void gopbrt_multi_shutdown(gopbrt_multi*);
int gopbrt_multi_device_count(const gopbrt_multi*);
const char* gopbrt_multi_last_error(const gopbrt_multi*);
uint64_t gopbrt_multi_launch_count(const gopbrt_multi*);
int gopbrt_multi_scene_create(gopbrt_multi*, const gopbrt_scene_desc*, gopbrt_multi_scene** out);
void gopbrt_multi_scene_destroy(gopbrt_multi_scene*);
/* flags: GOPBRT_FLAG_* of gopbrt_render_options (the reduce flag is implied) */
int gopbrt_multi_render(gopbrt_multi_scene*, const gopbrt_camera*, const gopbrt_sampler*, const gopbrt_integrator*,
                        const gopbrt_film*, int flags, double* film_out, gopbrt_stats* stats_out);
int gopbrt_multi_cancel(gopbrt_multi_scene*);

/* Self-test hook (tests/test_shading_kats.py): evaluates ONE device function of the raygen / shade / film stages on the GPU,
 * single thread, on flat float64 arguments, so the functions the kernels call can be pinned against known answers derived
 * independently of the oracle (FrDielectric reflection.go:21-42, OrenNayar.F :616-652, FresnelSpecular.SampleF :482-523,
 * BSDF.SampleF :183-253, the sampling warps sampling.go:173-198, Distribution1D.SampleDiscrete :42-55, the lights' SampleLi,
 * FilmTile.AddSample film.go:211-248, RGBToXYZ spectrum.go:35-41, the RNG rng.go:28-57, Stratified.StartPixel
 * stratified.go:21-48, GenerateRayDifferential camera.go:192-242, SpawnRayToInteraction interaction.go:91-102).
 * fn: index into the KAT_* list of csrc/gp_kat.cuh.  Returns the number of doubles written, or -GOPBRT_ERR_*. */
int gopbrt_kat_eval(gopbrt_scene*, int fn, const double* in, int n_in, double* out, int n_out);

/* number of kernel launches this library has issued on the ctx since init (bench.py "gpu_launches") */
uint64_t gopbrt_launch_count(const gopbrt_ctx*);

#ifdef __cplusplus
}
#endif
#endif /* GOPBRT_CUDA_H */
