// gp_kat.cuh — self-test hook of the library: evaluates ONE device function of the raygen / shade / film stages on the
// GPU (a single thread) on flat float64 arguments, so that tests can pin the very functions the kernels call against
// known answers computed independently of the oracle (tests/golden/make_shading_kats.py, tests/test_shading_kats.py).
// Nothing here is on the render path.
#pragma once
#include "gp_render.cuh"

namespace gp {

enum {
  KAT_FR_DIELECTRIC = 0, KAT_OREN_NAYAR_F, KAT_FRESNEL_SPECULAR_SAMPLE_F, KAT_CONCENTRIC_SAMPLE_DISK, KAT_COSINE_SAMPLE_HEMISPHERE,
  KAT_LAMBERT_SAMPLE_F, KAT_OFFSET_RAY_ORIGIN, KAT_COORDINATE_SYSTEM, KAT_SAMPLE_DISCRETE_UNIFORM, KAT_RGB_TO_XYZ, KAT_FILM_ADD_SAMPLE,
  KAT_RNG_U32, KAT_RNG_UNIFORM, KAT_RNG_U32B, KAT_STRATIFIED_START_PIXEL, KAT_LIGHT_SAMPLE_LI, KAT_SPAWN_RAY_TO, KAT_CAMERA_RAY, KAT_GO_MATH, KAT_N
};

GP_D void kat_identity_frame(BSDF& b) {
  b.ns = mk3(0, 0, 1); b.ng = mk3(0, 0, 1); b.ss = mk3(1, 0, 0); b.ts = mk3(0, 1, 0);
  b.eta = 1.0; b.a = 0; b.b = 0; b.etaB = 1.0; b.kind = BX_NONE; b.type = 0;
  b.r = rgb(0, 0, 0); b.t = rgb(0, 0, 0);
}

// scratch: device doubles the function may use (film tile / sampler tables / light cdf); P: film + sampler parameters
__global__ void k_kat(DevScene sc, int fn, const double* __restrict__ in, int n_in, double* __restrict__ out, int n_out, double* scratch,
                      RenderParams P, int* n_written) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  auto v3 = [&](int i) { return mk3(in[i], in[i + 1], in[i + 2]); };
  auto put3 = [&](int i, V3 v) { out[i] = v.x; out[i + 1] = v.y; out[i + 2] = v.z; };
  auto putc = [&](int i, RGB c) { out[i] = c.r; out[i + 1] = c.g; out[i + 2] = c.b; };
  int n = -1;
  switch (fn) {
    case KAT_FR_DIELECTRIC: if (n_in == 3) { out[0] = fr_dielectric(in[0], in[1], in[2]); n = 1; } break;
    case KAT_OREN_NAYAR_F: if (n_in == 10) {
      BSDF b; kat_identity_frame(b);
      b.r = rgb(in[1], in[2], in[3]); b.type = BSDF_REFLECTION | BSDF_DIFFUSE;
      oren_nayar_init(&b, in[0]);
      putc(0, bxdf_f(b, v3(4), v3(7))); n = 3;
    } break;
    case KAT_FRESNEL_SPECULAR_SAMPLE_F: if (n_in == 12) {
      BSDF b; kat_identity_frame(b);
      b.kind = BX_FRESNEL_SPECULAR; b.type = BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_SPECULAR;
      b.r = rgb(in[0], in[1], in[2]); b.t = rgb(in[3], in[4], in[5]); b.etaB = in[6];
      RGB f; V3 wi; double pdf; int st;
      bsdf_sample_f(b, v3(7), in[10], in[11], BSDF_ALL, &f, &wi, &pdf, &st);
      putc(0, f); put3(3, wi); out[6] = pdf; out[7] = (double)st; n = 8;
    } break;
    case KAT_CONCENTRIC_SAMPLE_DISK: if (n_in == 2) { concentric_sample_disk(in[0], in[1], &out[0], &out[1]); n = 2; } break;
    case KAT_COSINE_SAMPLE_HEMISPHERE: if (n_in == 2) { put3(0, cosine_sample_hemisphere(in[0], in[1])); n = 3; } break;
    case KAT_LAMBERT_SAMPLE_F: if (n_in == 8) {
      BSDF b; kat_identity_frame(b);
      b.kind = BX_LAMBERT; b.type = BSDF_REFLECTION | BSDF_DIFFUSE; b.r = rgb(in[0], in[1], in[2]);
      RGB f; V3 wi; double pdf; int st;
      bsdf_sample_f(b, v3(3), in[6], in[7], BSDF_ALL, &f, &wi, &pdf, &st);
      putc(0, f); put3(3, wi); out[6] = pdf; n = 7;
    } break;
    case KAT_OFFSET_RAY_ORIGIN: if (n_in == 12) { put3(0, offset_ray_origin(v3(0), v3(3), v3(6), v3(9))); n = 3; } break;
    case KAT_COORDINATE_SYSTEM: if (n_in == 3) { V3 a, b; coordinate_system(v3(0), &a, &b); put3(0, a); put3(3, b); n = 6; } break;
    case KAT_SAMPLE_DISCRETE_UNIFORM: if (n_in == 2) {  // scratch = {func_int, cdf[0..nl]} built by the host's scene code
      int off; double pdf;
      sample_discrete(scratch + 1, (int)in[0], scratch[0], in[1], &off, &pdf);
      out[0] = off; out[1] = pdf; n = 2;
    } break;
    case KAT_RGB_TO_XYZ: if (n_in == 3) { rgb_to_xyz(in[0], in[1], in[2], &out[0], &out[1], &out[2]); n = 3; } break;
    case KAT_FILM_ADD_SAMPLE: if (n_in == 11) {  // P: film geometry; scratch: the lane's FilmTile (tpw * tph * 4, zeroed)
      Lanes L;
      L.n = 1; L.tilepix = scratch; L.tile_stride = (long long)P.tpw * P.tph * 4;
      long long tile = (long long)in[5];
      film_add_sample(L, P, 0, tile, in[6], in[7], rgb(in[8], in[9], in[10]));
      long long x0, y0, x1, y1, bx0, by0, bx1, by1;
      tile_bounds(P, tile, &x0, &y0, &x1, &y1);
      tile_pixel_bounds(P, x0, y0, x1, y1, &bx0, &by0, &bx1, &by1);
      int k = 0;
      out[k++] = (double)bx0; out[k++] = (double)by0; out[k++] = (double)bx1; out[k++] = (double)by1;
      for (int y = 0; y < P.tph; y++)
        for (int x = 0; x < P.tpw; x++) {
          const double* p = scratch + ((size_t)y * P.tpw + x) * 4;
          if (p[3] == 0) continue;
          if (k + 6 > n_out) { k = -1; y = P.tph; break; }
          out[k++] = (double)(bx0 + x); out[k++] = (double)(by0 + y); out[k++] = p[0]; out[k++] = p[1]; out[k++] = p[2]; out[k++] = p[3];
        }
      n = k;
    } break;
    case KAT_RNG_U32: if (n_in == 3 && (int)in[2] <= n_out) {
      Smp s; s.state = 0x853c49e6748fea9bULL; s.inc = 0xda3e39cb94b95bdbULL;
      if (in[0] != 0) rng_set_sequence(s, (unsigned long long)in[1]);
      n = (int)in[2];
      for (int i = 0; i < n; i++) out[i] = (double)rng_u32(s);
    } break;
    case KAT_RNG_UNIFORM: if (n_in == 2 && (int)in[1] <= n_out) {
      Smp s; rng_set_sequence(s, (unsigned long long)in[0]);
      n = (int)in[1];
      for (int i = 0; i < n; i++) out[i] = rng_uniform(s);
    } break;
    case KAT_RNG_U32B: if (n_in == 3 && (int)in[2] <= n_out) {
      Smp s; rng_set_sequence(s, (unsigned long long)in[0]);
      n = (int)in[2];
      for (int i = 0; i < n; i++) out[i] = (double)rng_u32b(s, (uint32_t)((int)in[1] - i));
    } break;
    case KAT_STRATIFIED_START_PIXEL: if (n_in == 5) {  // P: sampler parameters; scratch: tables[dim][k] of ONE lane
      Lanes L;
      L.n = 1; L.tables = scratch;
      Smp s; s.lane = 0; s.cur1 = 0; s.cur2 = 0; s.sidx = 0;
      rng_set_sequence(s, (unsigned long long)in[0]);
      start_pixel(s, L, P);
      int k = P.ndims * P.spp;
      if (k + 2 <= n_out) {
        for (int i = 0; i < k; i++) out[i] = scratch[i];
        out[k] = rng_uniform(s); out[k + 1] = rng_uniform(s);
        n = k + 2;
      }
    } break;
    case KAT_LIGHT_SAMPLE_LI: if (n_in == 12 && n_out >= 17 && (int)in[0] >= 0 && (int)in[0] < sc.n_lights) {
      Intr ref; ref.p = v3(1); ref.perr = v3(4); ref.n = v3(7);
      LightSample ls;
      ls.p1.p = mk3(0, 0, 0); ls.p1.perr = mk3(0, 0, 0); ls.p1.n = mk3(0, 0, 0);
      light_sample_li(sc, sc.lights[(int)in[0]], ref, in[10], in[11], &ls);
      putc(0, ls.Li); put3(3, ls.wi); out[6] = ls.pdf; put3(7, ls.p1.p); put3(10, ls.p1.perr); put3(13, ls.p1.n); out[16] = ls.delta ? 1.0 : 0.0;
      n = 17;
    } break;
    case KAT_SPAWN_RAY_TO: if (n_in == 18) {
      Intr a, b; a.p = v3(0); a.perr = v3(3); a.n = v3(6); b.p = v3(9); b.perr = v3(12); b.n = v3(15);
      put3(0, a.p); put3(3, spawn_ray_to(a, b)); out[6] = 1 - 0.0001; n = 7;
    } break;
    case KAT_GO_MATH: if (n_in == 2) {  // math.Max, math.Min, math.Sin, math.Cos, and the fused sin + cos the kernels use
      const SinCos sc2 = go_sincos(in[0]);
      out[0] = go_max(in[0], in[1]); out[1] = go_min(in[0], in[1]); out[2] = go_sin(in[0]); out[3] = go_cos(in[0]);
      out[4] = sc2.sn; out[5] = sc2.cs; n = 6;
    } break;
    case KAT_CAMERA_RAY: if (n_in == 38) {
      Ray r = camera_ray(P, in[34], in[35], in[36], in[37]);
      put3(0, r.o); put3(3, r.d); n = 6;
    } break;
  }
  *n_written = n;
}

}  // namespace gp
