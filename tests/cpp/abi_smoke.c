/* abi_smoke.c — the drop-in boundary driven from plain C99, with no Python and no C++ in the caller: what a cgo / JNI / FFI
 * binding of include/gopbrt_cuda.h does, call for call:  gopbrt_init -> gopbrt_scene_create -> gopbrt_trace_closest /
 * gopbrt_trace_any -> gopbrt_render -> gopbrt_multi_* (one device) -> teardown.
 * Built by tests/test_abi.py with `gcc -std=c99 -pedantic` against libgopbrt_cuda.so (link check on the CPU box) and run by
 * tests/test_gpu_multi.py on the GPU box.  Exit code 0 and a final "abi_smoke: ok" line mean every check held. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gopbrt_cuda.h"

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { failures++; printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } while (0)

static gopbrt_transform translate(double x, double y, double z) { /* pbrt.Translate (transform.go:347-362) */
  gopbrt_transform t;
  int i;
  memset(&t, 0, sizeof t);
  for (i = 0; i < 4; i++) { t.m[5 * i] = 1.0; t.minv[5 * i] = 1.0; }
  t.m[3] = x; t.m[7] = y; t.m[11] = z;
  t.minv[3] = -x; t.minv[7] = -y; t.minv[11] = -z;
  return t;
}

int main(void) {
  enum { W = 48, H = 27, XS = 3, YS = 3 };
  gopbrt_ctx* ctx = NULL;
  gopbrt_scene* scene = NULL;
  gopbrt_scene_desc d;
  gopbrt_transform xf[2];
  gopbrt_sphere spheres[2];
  gopbrt_primitive prims[2];
  gopbrt_material mat;
  gopbrt_texture tex;
  gopbrt_light light;
  gopbrt_camera cam;
  gopbrt_sampler smp;
  gopbrt_integrator integ;
  gopbrt_film film;
  gopbrt_render_options opt;
  gopbrt_stats st, mst;
  double* px = (double*)malloc(sizeof(double) * W * H * 4);
  double* mpx = (double*)malloc(sizeof(double) * W * H * 4);
  double bound[6];
  int rc, i;

  CHECK(gopbrt_abi_version() == GOPBRT_ABI_VERSION, "ABI version");
  rc = gopbrt_init(0, &ctx);
  if (rc != GOPBRT_OK) { printf("gopbrt_init failed (%d): a B200 is required, there is no CPU fallback\n", rc); return 2; }

  /* a unit sphere at the origin on a large ground sphere, one point light, one matte material */
  xf[0] = translate(0, 0, 0);
  xf[1] = translate(0, -101, 0);
  memset(spheres, 0, sizeof spheres);
  spheres[0].object_to_world = 0; spheres[0].radius = 1; spheres[0].z_min = -1; spheres[0].z_max = 1; spheres[0].phi_max_deg = 360;
  spheres[1].object_to_world = 1; spheres[1].radius = 100; spheres[1].z_min = -100; spheres[1].z_max = 100; spheres[1].phi_max_deg = 360;
  for (i = 0; i < 2; i++) { prims[i].shape_kind = GOPBRT_SHAPE_SPHERE; prims[i].shape_index = i; prims[i].material = 0; prims[i].prim_to_world = -1; }
  memset(&tex, 0, sizeof tex);
  tex.kind = GOPBRT_TEX_CONSTANT; tex.tex1 = tex.tex2 = -1; tex.rgb[0] = tex.rgb[1] = tex.rgb[2] = 0.5;
  memset(&mat, 0, sizeof mat);
  mat.kind = GOPBRT_MAT_MATTE; mat.tex_a = 0; mat.tex_b = -1;
  memset(&light, 0, sizeof light);
  light.kind = GOPBRT_LIGHT_POINT; light.rgb[0] = light.rgb[1] = light.rgb[2] = 40; light.v[0] = 3; light.v[1] = 4; light.v[2] = -3;
  memset(&d, 0, sizeof d);
  d.n_transforms = 2; d.transforms = xf;
  d.n_spheres = 2; d.spheres = spheres;
  d.n_primitives = 2; d.primitives = prims;
  d.n_materials = 1; d.materials = &mat;
  d.n_textures = 1; d.textures = &tex;
  d.n_lights = 1; d.lights = &light;
  d.max_prims_in_node = 2;
  rc = gopbrt_scene_create(ctx, &d, &scene);
  CHECK(rc == GOPBRT_OK, "gopbrt_scene_create: %d %s", rc, gopbrt_last_error(ctx));
  if (rc != GOPBRT_OK) return 1;
  CHECK(gopbrt_scene_world_bound(scene, bound) == GOPBRT_OK && bound[0] == -100 && bound[3] == 100 && bound[4] == 1 && bound[1] == -201, "world bound");

  { /* Aggregate.Intersect / IntersectP over a batch of three rays */
    double ox[3] = {0, 0, 0}, oy[3] = {0, 2, 50}, oz[3] = {-5, 0, 0};
    double dx[3] = {0, 0, 0}, dy[3] = {0, 1, -1}, dz[3] = {1, 0, 0};
    double tmax[3], t[3], p[9], n[9];
    int32_t prim[3];
    uint8_t hit[3];
    tmax[0] = tmax[1] = tmax[2] = HUGE_VAL;
    rc = gopbrt_trace_closest(scene, 3, ox, oy, oz, dx, dy, dz, tmax, prim, t, p, n);
    CHECK(rc == GOPBRT_OK, "gopbrt_trace_closest: %d", rc);
    CHECK(prim[0] == 0 && fabs(t[0] - 4.0) < 1e-12 && fabs(p[2] + 1.0) < 1e-12 && fabs(n[2] + 1.0) < 1e-12, "ray 0: prim %d t %.17g", prim[0], t[0]);
    CHECK(prim[1] == -1 && t[1] == HUGE_VAL, "ray 1 should escape: prim %d", prim[1]);
    CHECK(prim[2] == 0 && fabs(t[2] - 49.0) < 1e-12, "ray 2: prim %d t %.17g", prim[2], t[2]);
    rc = gopbrt_trace_any(scene, 3, ox, oy, oz, dx, dy, dz, tmax, hit);
    CHECK(rc == GOPBRT_OK && hit[0] == 1 && hit[1] == 0 && hit[2] == 1, "gopbrt_trace_any: %d (%d %d %d)", rc, hit[0], hit[1], hit[2]);
    CHECK(gopbrt_trace_closest(scene, 0, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL) == GOPBRT_OK, "empty batch");
    CHECK(gopbrt_trace_closest(scene, 3, NULL, oy, oz, dx, dy, dz, tmax, prim, t, NULL, NULL) == GOPBRT_ERR_INVALID, "NULL array must be rejected");
  }

  /* pbrt.Render: a pinhole camera 5 units in front of the sphere, 90 degree field of view, Stratified 3x3, Path maxDepth 5 */
  memset(&cam, 0, sizeof cam);
  {
    double ar = (double)W / H;
    double r2c[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 1};
    r2c[0] = 2 * ar / W; r2c[3] = -ar; r2c[5] = -2.0 / H; r2c[7] = 1;
    memcpy(cam.raster_to_camera, r2c, sizeof r2c);
    memcpy(cam.camera_to_world, translate(0, 0, -5).m, sizeof cam.camera_to_world);
  }
  cam.focal_distance = 1e6;
  memset(&smp, 0, sizeof smp);
  smp.kind = GOPBRT_SAMPLER_STRATIFIED; smp.x_samples = XS; smp.y_samples = YS; smp.n_sampled_dimensions = 4; smp.mode = GOPBRT_MODE_STRICT;
  memset(&integ, 0, sizeof integ);
  integ.kind = GOPBRT_INTEGRATOR_PATH; integ.max_depth = 5; integ.rr_threshold = 1; integ.light_strategy = GOPBRT_LIGHTS_UNIFORM; integ.tile_size = 16;
  memset(&film, 0, sizeof film);
  film.width = W; film.height = H; film.crop[2] = film.crop[3] = 1; film.filter_radius[0] = film.filter_radius[1] = 0.5;
  memset(&opt, 0, sizeof opt);
  opt.world = 1;
  rc = gopbrt_render(scene, &cam, &smp, &integ, &film, &opt, px, &st);
  CHECK(rc == GOPBRT_OK, "gopbrt_render: %d %s", rc, gopbrt_last_error(ctx));
  CHECK(st.camera_rays == (uint64_t)W * H * (XS * YS - 1), "camera rays %llu (samples 1..spp-1 run, sampler.go:29-34)", (unsigned long long)st.camera_rays);
  CHECK(st.closest_rays >= st.camera_rays && st.shadow_rays > 0 && st.launches > 0, "ray counters");
  {
    int bad_w = 0, lit = 0;
    /* unjittered Stratified leaves every pFilm on its pixel's integer corner (the 2-D tables are never written, sampling.go:112-127),
       so with the box filter of radius 0.5 a sample weighs 1 in the 2 x 2 pixels around that corner (film.go:211-248): an interior
       pixel collects 4 * (spp - 1) — its own corner's samples and its right / lower neighbours' —, a pixel of the last column or
       row 2 *, the last pixel 1 * */
    for (i = 0; i < W * H; i++) {
      int x = i % W, y = i / W;
      double want = (double)(XS * YS - 1) * (x == W - 1 ? 1 : 2) * (y == H - 1 ? 1 : 2);
      if (px[4 * i + 3] != want) bad_w++;
      if (px[4 * i + 1] > 0) lit++;
      if (!(px[4 * i] >= 0 && px[4 * i + 1] >= 0 && px[4 * i + 2] >= 0)) bad_w++;
    }
    CHECK(bad_w == 0, "%d pixels with a wrong filterWeightSum or a negative / NaN value", bad_w);
    CHECK(lit > W * H / 4, "only %d lit pixels", lit);
    CHECK(px[4 * ((H / 2) * W + W / 2) + 1] > 0, "the centre pixel sees the lit sphere");
  }
  /* the same frame again is bit-identical (STRICT mode is deterministic) */
  rc = gopbrt_render(scene, &cam, &smp, &integ, &film, &opt, mpx, NULL);
  CHECK(rc == GOPBRT_OK && memcmp(px, mpx, sizeof(double) * W * H * 4) == 0, "second render differs");
  /* a cancel request issued before the call is consumed by it */
  gopbrt_cancel(scene);
  CHECK(gopbrt_render(scene, &cam, &smp, &integ, &film, &opt, mpx, NULL) == GOPBRT_ERR_CANCELLED, "pending cancel must cancel the next frame");
  CHECK(gopbrt_render(scene, &cam, &smp, &integ, &film, &opt, mpx, NULL) == GOPBRT_OK, "... and only that one");
  /* bad arguments */
  integ.light_strategy = GOPBRT_LIGHTS_SPATIAL;
  CHECK(gopbrt_render(scene, &cam, &smp, &integ, &film, &opt, mpx, NULL) == GOPBRT_ERR_UNSUPPORTED, "Spatial strategy");
  integ.light_strategy = GOPBRT_LIGHTS_UNIFORM;
  opt.flags = GOPBRT_FLAG_REDUCE_FILM;
  CHECK(gopbrt_render(scene, &cam, &smp, &integ, &film, &opt, mpx, NULL) == GOPBRT_ERR_INVALID, "reduce without a communicator");
  opt.flags = 0;
  gopbrt_scene_destroy(scene);
  gopbrt_shutdown(ctx);

  { /* one process, N devices — here N = 1: the same frame through gopbrt_multi_render */
    gopbrt_multi* m = NULL;
    gopbrt_multi_scene* ms = NULL;
    rc = gopbrt_multi_init(1, NULL, &m);
    CHECK(rc == GOPBRT_OK && gopbrt_multi_device_count(m) == 1, "gopbrt_multi_init: %d", rc);
    rc = gopbrt_multi_scene_create(m, &d, &ms);
    CHECK(rc == GOPBRT_OK, "gopbrt_multi_scene_create: %d %s", rc, gopbrt_multi_last_error(m));
    if (rc == GOPBRT_OK) {
      rc = gopbrt_multi_render(ms, &cam, &smp, &integ, &film, 0, mpx, &mst);
      CHECK(rc == GOPBRT_OK && memcmp(px, mpx, sizeof(double) * W * H * 4) == 0, "gopbrt_multi_render: %d", rc);
      CHECK(mst.camera_rays == st.camera_rays && mst.closest_rays == st.closest_rays && mst.shadow_rays == st.shadow_rays, "multi stats");
      CHECK(gopbrt_multi_launch_count(m) > 0, "launch count");
      gopbrt_multi_scene_destroy(ms);
    }
    gopbrt_multi_shutdown(m);
  }
  free(px);
  free(mpx);
  if (failures) { printf("abi_smoke: %d FAILURES\n", failures); return 1; }
  printf("abi_smoke: ok (%llu rays, %llu kernel launches in the first frame)\n", (unsigned long long)(st.closest_rays + st.shadow_rays), (unsigned long long)st.launches);
  return 0;
}
