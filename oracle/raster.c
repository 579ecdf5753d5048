/* oracle/raster.c -- CPU restatement of the canonical+delta Gaussian-splat tile rasteriser.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for the sm_100a
 * rasteriser and the "port" CPU baseline of bench.py.  Never linked into the product.
 *
 * PARITY UNPINNED.  The algorithm restated here lives in a third-party dependency that is
 * absent from /root/reference: `diff_gaussian_rasterization` from
 * autonomousvision/mip-splatting (submodules/diff-gaussian-rasterization), installed by
 * the reference's setup.sh:220-224 at an UN-PINNED git HEAD.  It is restated from the
 * published 3D Gaussian Splatting / Mip-Splatting algorithm and anchored on the
 * reference's own call site and conventions:
 *   - call signature + settings            renderers/gaussian_render.py:105-125,198-206
 *   - camera matrices (transposed view / full projection, tanfov = 0.5/f)
 *                                          renderers/gaussian_render.py:57-82,302-321
 *   - activations + delta application      representations/gaussian/gaussian_model.py:84-114
 *   - SH degree 0 colour                   renderers/sh_utils.py (C0 = 0.28209479177387814)
 * No test or fixture in the reference pins rasteriser outputs.
 *
 * Stage structure follows upstream so the integer by-products are comparable:
 *   preprocess (cull z<=0.2, EWA cov2D, mip 2-D filter + opacity compensation,
 *   radius = ceil(3 sqrt(lambda_max)), 16x16 tile rectangle) -> inclusive scan of
 *   tiles_touched -> (tile<<32 | depth bits, gaussian id) pairs -> stable LSD radix sort
 *   -> per-tile ranges -> front-to-back alpha blend (alpha<1/255 skipped, min(.99,.),
 *   stop at T<1e-4, out = C + T*bg).
 *
 * Arithmetic that feeds integers goes through include/gvf_math.h and is written without
 * FMA contraction (compile with -ffp-contract=off) so that indices are bit-exact vs CUDA.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "../include/gvf_math.h"

#define TILE 16
#define SH_C0 0.28209479177387814f

/* float -> int truncation with a defined result for far-off-screen values (C leaves the
 * out-of-range conversion undefined; CUDA saturates).  Both sides clamp first. */
static inline int f2i(float v) {
  v = fminf(fmaxf(v, -1.0e6f), 1.0e6f);
  return (int)v;
}

typedef struct {
  int32_t H, W;
  float tanfovx, tanfovy;
  float kernel_size;      /* mip 2-D filter variance added to cov2D (0.1) */
  float scale_modifier;
  float bg[3];
  /* GaussianModel constants (representations/gaussian/gaussian_model.py:17-90) */
  float aabb[6];          /* xyz = _xyz * aabb[3:6] + aabb[0:3] */
  float scale_bias;       /* inverse_softplus(scaling_bias) */
  float min_kernel;       /* 3-D filter size (mininum_kernel_size) */
  float opacity_bias;     /* logit(opacity_bias) */
  int32_t softplus;       /* 1: softplus scaling activation, 0: exp */
  int32_t mip_filter;     /* 1: mip 2-D filter + opacity compensation; 0: plain 3DGS dilation by kernel_size */
} gvf_oracle_params;

/* ---- GaussianModel.get_*_with_delta: raw canonical params (+ delta) -> activated ---- */
/* delta layout (renderers/gaussian_render.py:155-160): [xyz3 | scale3 | rot4 | rgb3 | opacity1] */
void gvf_oracle_activate(const gvf_oracle_params* prm, int P,
                         const float* xyz_raw, const float* dc_raw, const float* scaling_raw,
                         const float* rotation_raw, const float* opacity_raw,
                         const float* delta /* [P,14] or NULL */,
                         float* means3D, float* scales, float* rots, float* shs, float* opac) {
  const float k2 = prm->min_kernel * prm->min_kernel;
  for (int i = 0; i < P; ++i) {
    const float* d = delta ? delta + (size_t)i * 14 : NULL;
    for (int c = 0; c < 3; ++c) {
      float v = xyz_raw[i * 3 + c] * prm->aabb[3 + c] + prm->aabb[c];
      means3D[i * 3 + c] = d ? v + d[c] : v;
    }
    for (int c = 0; c < 3; ++c) {
      float s = scaling_raw[i * 3 + c] + prm->scale_bias;
      if (d) s = s + d[3 + c];
      s = prm->softplus ? gvf_softplusf(s) : gvf_expf(s);
      scales[i * 3 + c] = sqrtf(s * s + k2);
    }
    float q[4];
    for (int c = 0; c < 4; ++c) {
      float r = rotation_raw[i * 4 + c] + (c == 0 ? 1.0f : 0.0f);
      q[c] = d ? r + d[6 + c] : r;
    }
    /* F.normalize: x / max(||x||, 1e-12) */
    float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    n = fmaxf(n, 1e-12f);
    for (int c = 0; c < 4; ++c) rots[i * 4 + c] = q[c] / n;
    for (int c = 0; c < 3; ++c) shs[i * 3 + c] = d ? dc_raw[i * 3 + c] + d[10 + c] : dc_raw[i * 3 + c];
    float o = opacity_raw[i] + prm->opacity_bias;
    if (d) o = o + d[13];
    opac[i] = gvf_sigmoidf(o);
  }
}

/* ---- per-Gaussian projection ---- */
typedef struct {
  float depth, px, py, ca, cb, cc, op, r, g, b;
  int32_t radius;
  int32_t x0, y0, x1, y1;   /* tile rect [x0,x1) x [y0,y1) */
} splat_t;

static int preprocess_one(const gvf_oracle_params* prm, const float* view, const float* proj,
                          const float* m3, const float* sc, const float* q, const float* sh,
                          float opacity, splat_t* out) {
  const int W = prm->W, H = prm->H;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const float px = m3[0], py = m3[1], pz = m3[2];
  /* view-space point (matrix memory = transposed view, i.e. column-major view) */
  const float tx = view[0] * px + view[4] * py + view[8] * pz + view[12];
  const float ty = view[1] * px + view[5] * py + view[9] * pz + view[13];
  const float tz = view[2] * px + view[6] * py + view[10] * pz + view[14];
  out->radius = 0;
  if (tz <= 0.2f) return 0;
  const float hx = proj[0] * px + proj[4] * py + proj[8] * pz + proj[12];
  const float hy = proj[1] * px + proj[5] * py + proj[9] * pz + proj[13];
  const float hw = proj[3] * px + proj[7] * py + proj[11] * pz + proj[15];
  const float pw = 1.0f / (hw + 0.0000001f);
  const float ndcx = hx * pw, ndcy = hy * pw;

  /* 3-D covariance from scale / rotation: Sigma = (S R)^T (S R), R from quaternion (w,x,y,z) */
  const float s0 = prm->scale_modifier * sc[0], s1 = prm->scale_modifier * sc[1],
              s2 = prm->scale_modifier * sc[2];
  const float r = q[0], x = q[1], y = q[2], z = q[3];
  const float R00 = 1.f - 2.f * (y * y + z * z), R01 = 2.f * (x * y - r * z), R02 = 2.f * (x * z + r * y);
  const float R10 = 2.f * (x * y + r * z), R11 = 1.f - 2.f * (x * x + z * z), R12 = 2.f * (y * z - r * x);
  const float R20 = 2.f * (x * z - r * y), R21 = 2.f * (y * z + r * x), R22 = 1.f - 2.f * (x * x + y * y);
  /* L = R * diag(s); Sigma = L L^T */
  const float L00 = R00 * s0, L01 = R01 * s1, L02 = R02 * s2;
  const float L10 = R10 * s0, L11 = R11 * s1, L12 = R12 * s2;
  const float L20 = R20 * s0, L21 = R21 * s1, L22 = R22 * s2;
  const float S00 = L00 * L00 + L01 * L01 + L02 * L02;
  const float S01 = L00 * L10 + L01 * L11 + L02 * L12;
  const float S02 = L00 * L20 + L01 * L21 + L02 * L22;
  const float S11 = L10 * L10 + L11 * L11 + L12 * L12;
  const float S12 = L10 * L20 + L11 * L21 + L12 * L22;
  const float S22 = L20 * L20 + L21 * L21 + L22 * L22;

  /* EWA projection: cov2D = J W Sigma W^T J^T */
  const float fx = (float)W / (2.0f * prm->tanfovx), fy = (float)H / (2.0f * prm->tanfovy);
  const float limx = 1.3f * prm->tanfovx, limy = 1.3f * prm->tanfovy;
  const float txtz = tx / tz, tytz = ty / tz;
  const float cx = fminf(limx, fmaxf(-limx, txtz)) * tz;
  const float cy = fminf(limy, fmaxf(-limy, tytz)) * tz;
  const float J00 = fx / tz, J02 = -(fx * cx) / (tz * tz);
  const float J11 = fy / tz, J12 = -(fy * cy) / (tz * tz);
  /* rows of the world->view rotation: Wr[i][j] = view[j*4+i] */
  const float W00 = view[0], W01 = view[4], W02 = view[8];
  const float W10 = view[1], W11 = view[5], W12 = view[9];
  const float W20 = view[2], W21 = view[6], W22 = view[10];
  /* T = J * Wr (2x3) */
  const float T00 = J00 * W00 + J02 * W20, T01 = J00 * W01 + J02 * W21, T02 = J00 * W02 + J02 * W22;
  const float T10 = J11 * W10 + J12 * W20, T11 = J11 * W11 + J12 * W21, T12 = J11 * W12 + J12 * W22;
  /* V = T * Sigma (2x3) */
  const float V00 = T00 * S00 + T01 * S01 + T02 * S02;
  const float V01 = T00 * S01 + T01 * S11 + T02 * S12;
  const float V02 = T00 * S02 + T01 * S12 + T02 * S22;
  const float V10 = T10 * S00 + T11 * S01 + T12 * S02;
  const float V11 = T10 * S01 + T11 * S11 + T12 * S12;
  const float V12 = T10 * S02 + T11 * S12 + T12 * S22;
  float ca = V00 * T00 + V01 * T01 + V02 * T02;
  const float cb = V00 * T10 + V01 * T11 + V02 * T12;
  float cc = V10 * T10 + V11 * T11 + V12 * T12;

  /* mip-splatting 2-D filter and opacity compensation */
  const float ks = prm->kernel_size;
  const float det0 = fmaxf(1e-6f, ca * cc - cb * cb);
  const float det1 = fmaxf(1e-6f, (ca + ks) * (cc + ks) - cb * cb);
  float coef = sqrtf(det0 / (det1 + 1e-6f) + 1e-6f);
  if (det0 <= 1e-6f || det1 <= 1e-6f) coef = 0.0f;
  if (!prm->mip_filter) coef = 1.0f;
  ca = ca + ks;
  cc = cc + ks;
  const float det = ca * cc - cb * cb;
  if (det == 0.0f) return 0;
  const float det_inv = 1.0f / det;
  const float mid = 0.5f * (ca + cc);
  const float disc = sqrtf(fmaxf(0.1f, mid * mid - det));
  const float lambda1 = mid + disc, lambda2 = mid - disc;
  const float rad = ceilf(3.0f * sqrtf(fmaxf(lambda1, lambda2)));
  const float pix_x = ((ndcx + 1.0f) * (float)W - 1.0f) * 0.5f;
  const float pix_y = ((ndcy + 1.0f) * (float)H - 1.0f) * 0.5f;
  const int irad = f2i(rad);
  int x0 = f2i((pix_x - (float)irad) / (float)TILE);
  int y0 = f2i((pix_y - (float)irad) / (float)TILE);
  int x1 = f2i((pix_x + (float)irad + (float)(TILE - 1)) / (float)TILE);
  int y1 = f2i((pix_y + (float)irad + (float)(TILE - 1)) / (float)TILE);
  x0 = x0 < 0 ? 0 : (x0 > gx ? gx : x0);
  y0 = y0 < 0 ? 0 : (y0 > gy ? gy : y0);
  x1 = x1 < 0 ? 0 : (x1 > gx ? gx : x1);
  y1 = y1 < 0 ? 0 : (y1 > gy ? gy : y1);
  if ((x1 - x0) * (y1 - y0) == 0) return 0;

  out->depth = tz;
  out->px = pix_x;
  out->py = pix_y;
  out->ca = cc * det_inv;
  out->cb = -cb * det_inv;
  out->cc = ca * det_inv;
  out->op = opacity * coef;
  out->r = fmaxf(SH_C0 * sh[0] + 0.5f, 0.0f);
  out->g = fmaxf(SH_C0 * sh[1] + 0.5f, 0.0f);
  out->b = fmaxf(SH_C0 * sh[2] + 0.5f, 0.0f);
  out->radius = irad;
  out->x0 = x0; out->y0 = y0; out->x1 = x1; out->y1 = y1;
  return (x1 - x0) * (y1 - y0);
}

/* stable LSD radix sort of (key,val) pairs on key bits [0,nbits) */
static void radix_sort_pairs(uint64_t* keys, uint32_t* vals, size_t n, int nbits) {
  uint64_t* k2 = (uint64_t*)malloc(n * sizeof(uint64_t));
  uint32_t* v2 = (uint32_t*)malloc(n * sizeof(uint32_t));
  for (int shift = 0; shift < nbits; shift += 8) {
    size_t cnt[257];
    memset(cnt, 0, sizeof(cnt));
    for (size_t i = 0; i < n; ++i) cnt[((keys[i] >> shift) & 0xff) + 1]++;
    for (int b = 0; b < 256; ++b) cnt[b + 1] += cnt[b];
    for (size_t i = 0; i < n; ++i) {
      size_t p = cnt[(keys[i] >> shift) & 0xff]++;
      k2[p] = keys[i];
      v2[p] = vals[i];
    }
    memcpy(keys, k2, n * sizeof(uint64_t));
    memcpy(vals, v2, n * sizeof(uint32_t));
  }
  free(k2);
  free(v2);
}

/* One frame, activated inputs.  Outputs may be NULL when not wanted (except out_rgba).
 * out_rgba: [4,H,W] (RGB composited over bg, A = 1 - T_final).
 * splat: [P,10] (depth,px,py,conic a,b,c,opacity',r,g,b) ; rects: [P,4] tile rect.
 * ranges: [tiles,2] ; point_list: capacity `cap` ids in (tile, depth, id) order.
 * returns num_rendered, or -1 if cap is too small. */
int64_t gvf_oracle_forward(const gvf_oracle_params* prm, int P, const float* means3D,
                           const float* scales, const float* rots, const float* shs,
                           const float* opac, const float* view, const float* proj,
                           const float* subpixel_offset /* [H,W,2] or NULL */,
                           float* out_rgba, int32_t* radii, uint32_t* tiles_touched,
                           float* splat, int32_t* rects, uint32_t* ranges,
                           uint32_t* point_list, uint64_t* key_list, int64_t cap,
                           uint32_t* n_contrib, float* final_T) {
  const int W = prm->W, H = prm->H;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int ntiles = gx * gy;
  splat_t* sp = (splat_t*)calloc((size_t)P, sizeof(splat_t));
  uint32_t* offs = (uint32_t*)malloc((size_t)P * sizeof(uint32_t));
  uint32_t run = 0;
  for (int i = 0; i < P; ++i) {
    int t = preprocess_one(prm, view, proj, means3D + 3 * i, scales + 3 * i, rots + 4 * i,
                           shs + 3 * i, opac[i], &sp[i]);
    run += (uint32_t)t;
    offs[i] = run; /* inclusive scan */
    if (radii) radii[i] = sp[i].radius;
    if (tiles_touched) tiles_touched[i] = (uint32_t)t;
    if (splat) {
      float* s = splat + (size_t)i * 10;
      s[0] = sp[i].depth; s[1] = sp[i].px; s[2] = sp[i].py; s[3] = sp[i].ca; s[4] = sp[i].cb;
      s[5] = sp[i].cc; s[6] = sp[i].op; s[7] = sp[i].r; s[8] = sp[i].g; s[9] = sp[i].b;
    }
    if (rects) {
      rects[i * 4 + 0] = sp[i].x0; rects[i * 4 + 1] = sp[i].y0;
      rects[i * 4 + 2] = sp[i].x1; rects[i * 4 + 3] = sp[i].y1;
    }
  }
  const size_t R = run;
  uint64_t* keys = (uint64_t*)malloc((R ? R : 1) * sizeof(uint64_t));
  uint32_t* vals = (uint32_t*)malloc((R ? R : 1) * sizeof(uint32_t));
  for (int i = 0; i < P; ++i) {
    if (sp[i].radius <= 0) continue;
    size_t off = (i == 0) ? 0 : offs[i - 1];
    for (int y = sp[i].y0; y < sp[i].y1; ++y)
      for (int x = sp[i].x0; x < sp[i].x1; ++x) {
        keys[off] = ((uint64_t)(uint32_t)(y * gx + x) << 32) | gvf_f2u(sp[i].depth);
        vals[off] = (uint32_t)i;
        ++off;
      }
  }
  int tbits = 0;
  while ((1 << tbits) < ntiles) ++tbits;
  radix_sort_pairs(keys, vals, R, 32 + tbits);
  uint32_t* rng = (uint32_t*)calloc((size_t)ntiles * 2, sizeof(uint32_t));
  for (size_t i = 0; i < R; ++i) {
    uint32_t t = (uint32_t)(keys[i] >> 32);
    if (i == 0 || t != (uint32_t)(keys[i - 1] >> 32)) rng[t * 2] = (uint32_t)i;
    if (i + 1 == R || t != (uint32_t)(keys[i + 1] >> 32)) rng[t * 2 + 1] = (uint32_t)(i + 1);
  }
  if (ranges) memcpy(ranges, rng, (size_t)ntiles * 2 * sizeof(uint32_t));
  int64_t ret = (int64_t)R;
  if (point_list || key_list) {
    if ((int64_t)R > cap) ret = -1;
    else {
      if (point_list) memcpy(point_list, vals, R * sizeof(uint32_t));
      if (key_list) memcpy(key_list, keys, R * sizeof(uint64_t));
    }
  }

  /* blend */
  for (int ty = 0; ty < gy; ++ty)
    for (int txx = 0; txx < gx; ++txx) {
      const uint32_t s = rng[(ty * gx + txx) * 2], e = rng[(ty * gx + txx) * 2 + 1];
      for (int ly = 0; ly < TILE; ++ly)
        for (int lx = 0; lx < TILE; ++lx) {
          const int pxi = txx * TILE + lx, pyi = ty * TILE + ly;
          if (pxi >= W || pyi >= H) continue;
          const size_t pid = (size_t)pyi * W + pxi;
          float pfx = (float)pxi, pfy = (float)pyi;
          if (subpixel_offset) {
            pfx += subpixel_offset[pid * 2];
            pfy += subpixel_offset[pid * 2 + 1];
          }
          float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
          uint32_t contributor = 0, last = 0;
          for (uint32_t j = s; j < e; ++j) {
            ++contributor;
            const splat_t* g = &sp[vals[j]];
            const float dx = g->px - pfx, dy = g->py - pfy;
            const float power = -0.5f * (g->ca * dx * dx + g->cc * dy * dy) - g->cb * dx * dy;
            if (power > 0.0f) continue;
            const float alpha = fminf(0.99f, g->op * expf(power));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = T * (1.0f - alpha);
            if (test_T < 0.0001f) break;
            C0 += g->r * alpha * T;
            C1 += g->g * alpha * T;
            C2 += g->b * alpha * T;
            T = test_T;
            last = contributor;
          }
          out_rgba[0 * (size_t)H * W + pid] = C0 + T * prm->bg[0];
          out_rgba[1 * (size_t)H * W + pid] = C1 + T * prm->bg[1];
          out_rgba[2 * (size_t)H * W + pid] = C2 + T * prm->bg[2];
          out_rgba[3 * (size_t)H * W + pid] = 1.0f - T;
          if (n_contrib) n_contrib[pid] = last;
          if (final_T) final_T[pid] = T;
        }
    }
  free(rng);
  free(keys);
  free(vals);
  free(offs);
  free(sp);
  return ret;
}

/* F frames of canonical + per-frame delta, one camera per frame (the path
 * render_and_save_images walks, utils/inference_utils.py:256-269).  OpenMP over frames.
 * canonical raw params are shared; delta [F,P,14] or NULL; views/projs [F,16].
 * out_rgba [F,4,H,W]; radii [F,P] or NULL; num_rendered [F] or NULL. */
int gvf_oracle_render_frames(const gvf_oracle_params* prm, int F, int P, const float* xyz_raw,
                             const float* dc_raw, const float* scaling_raw,
                             const float* rotation_raw, const float* opacity_raw,
                             const float* delta, const float* views, const float* projs,
                             float* out_rgba, int32_t* radii, int64_t* num_rendered) {
  const size_t HW = (size_t)prm->H * prm->W;
  int err = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int f = 0; f < F; ++f) {
    float* buf = (float*)malloc((size_t)P * 14 * sizeof(float));
    float *m3 = buf, *sc = buf + 3 * (size_t)P, *rt = buf + 6 * (size_t)P,
          *sh = buf + 10 * (size_t)P, *op = buf + 13 * (size_t)P;
    gvf_oracle_activate(prm, P, xyz_raw, dc_raw, scaling_raw, rotation_raw, opacity_raw,
                        delta ? delta + (size_t)f * P * 14 : NULL, m3, sc, rt, sh, op);
    int64_t r = gvf_oracle_forward(prm, P, m3, sc, rt, sh, op, views + 16 * (size_t)f,
                                   projs + 16 * (size_t)f, NULL, out_rgba + (size_t)f * 4 * HW,
                                   radii ? radii + (size_t)f * P : NULL, NULL, NULL, NULL, NULL,
                                   NULL, NULL, 0, NULL, NULL);
    if (num_rendered) num_rendered[f] = r;
    if (r < 0) err = 1;
    free(buf);
  }
  return err;
}

/* exported so tests can check gvf_math.h against float64 libm */
float gvf_oracle_expf(float x) { return gvf_expf(x); }
float gvf_oracle_logf(float x) { return gvf_logf(x); }
float gvf_oracle_log1pf(float x) { return gvf_log1pf(x); }
float gvf_oracle_softplusf(float x) { return gvf_softplusf(x); }
float gvf_oracle_sigmoidf(float x) { return gvf_sigmoidf(x); }
