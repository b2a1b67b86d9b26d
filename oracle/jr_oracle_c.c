/*
 * CPU ORACLE (C) -- test infrastructure, NOT product code.
 *
 * Plain-C restatement of the reference's brute-force depth path
 * (renderer/pipeline.py:470-537 with shaders/depth.py:38-61): for every pixel
 * evaluate EVERY triangle (pipeline.py:332-335), first-index argmin over
 * `keep & inside & front` depths (shader.py:207-217), triangle-0 fallback,
 * no test against the incoming z (pipeline.py:401-440).  Arithmetic follows
 * the torch oracle (oracle/jr_oracle.py) operation for operation; compile with
 * -ffp-contract=off so no multiply-add is fused.  tests/test_oracle_c.py checks
 * it bit-for-bit against jr_oracle.py; bench.py times it as the CPU baseline
 * ("port": the reference's own algorithm on the host cores, pthreads over
 * image rows; the image has no OpenMP runtime for gcc).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

static void to_clip(const float* M, const float* p, float* out) {
  for (int r = 0; r < 4; ++r)
    out[r] = ((p[0] * M[4 * r + 0] + p[1] * M[4 * r + 1]) + p[2] * M[4 * r + 2]) + M[4 * r + 3];
}

static float det3(const float* a) {
  return a[0] * a[4] * a[8] + a[1] * a[5] * a[6] + a[2] * a[3] * a[7] - a[2] * a[4] * a[6] -
         a[0] * a[5] * a[7] - a[1] * a[3] * a[8];
}

static void swapf(float* a, float* b) { float t = *a; *a = *b; *b = t; }

/* jnp.linalg.inv 3x3: sgetrf2-order LU + strsm-order substitutions (pipeline.py:105). */
static void lu_inverse3(const float* A, float* inv) {
  float r0[6] = {A[0], A[1], A[2], 1.f, 0.f, 0.f};
  float r1[6] = {A[3], A[4], A[5], 0.f, 1.f, 0.f};
  float r2[6] = {A[6], A[7], A[8], 0.f, 0.f, 1.f};
  float a0 = fabsf(r0[0]), a1 = fabsf(r1[0]), a2 = fabsf(r2[0]);
  int p1 = a1 > a0;
  float best = p1 ? a1 : a0;
  int p2 = a2 > best;
  p1 = p1 && !p2;
  if (p1) for (int c = 0; c < 6; ++c) swapf(&r0[c], &r1[c]);
  if (p2) for (int c = 0; c < 6; ++c) swapf(&r0[c], &r2[c]);
  float r00 = 1.0f / r0[0];
  float l10 = r1[0] * r00, l20 = r2[0] * r00;
  float u00 = r0[0], u01 = r0[1], u02 = r0[2];
  float a11 = r1[1] - l10 * u01, a12 = r1[2] - l10 * u02;
  float a21 = r2[1] - l20 * u01, a22 = r2[2] - l20 * u02;
  if (fabsf(a21) > fabsf(a11)) {
    swapf(&l10, &l20); swapf(&a11, &a21); swapf(&a12, &a22);
    for (int c = 3; c < 6; ++c) swapf(&r1[c], &r2[c]);
  }
  float u11 = a11, u12 = a12;
  float l21 = a21 * (1.0f / u11);
  float u22 = a22 - l21 * u12;
  for (int j = 0; j < 3; ++j) {
    float y0 = r0[3 + j];
    float y1 = r1[3 + j] - y0 * l10;
    float y2 = (r2[3 + j] - y0 * l20) - y1 * l21;
    float x2 = y2 / u22;
    float t1 = y1 - x2 * u12;
    float t0 = y0 - x2 * u02;
    float x1 = t1 / u11;
    t0 = t0 - x1 * u01;
    float x0 = t0 / u00;
    inv[0 + j] = x0; inv[3 + j] = x1; inv[6 + j] = x2;
  }
}

typedef struct RowJob {
  int W, H, T, keep0, tid, nthreads;
  size_t Tp;
  const float* vp;
  float* const* inv;
  float* const* zc;
  const float* candf;
  float* zb;
  int32_t* tb;
  int rc;
  /* general visibility mode (jr_oracle_visibility): per-pixel argmin index, "some candidate", keep&inside of
   * the chosen triangle, second-best minus best depth -- what oracle/jr_oracle.py::visibility returns */
  int32_t* idx_out;
  uint8_t* has_out;
  uint8_t* kc_out;
  float* gap_out;
} RowJob;

static void* row_worker(void* arg) {
  RowJob* j = (RowJob*)arg;
  const int H = j->H, T = j->T;
  const size_t Tp = j->Tp;
  const float* vp = j->vp;
  float* const* inv = j->inv;
  float* const* zc = j->zc;
  const float* candf = j->candf;
  const float vp22 = vp[10], vp23 = vp[11];
  float* restrict depth = (float*)aligned_alloc(64, sizeof(float) * Tp + 64);
  if (!depth) { j->rc = -1; return NULL; }
  const float* restrict i0 = inv[0]; const float* restrict i1 = inv[1]; const float* restrict i2 = inv[2];
  const float* restrict i3 = inv[3]; const float* restrict i4 = inv[4]; const float* restrict i5 = inv[5];
  const float* restrict i6 = inv[6]; const float* restrict i7 = inv[7]; const float* restrict i8 = inv[8];
  const float* restrict cand = candf;
  const float inf = INFINITY;
  const float* restrict z0 = zc[0]; const float* restrict z1 = zc[1]; const float* restrict z2 = zc[2];
  for (int x = j->tid; x < j->W; x += j->nthreads) {
    const float xn = ((float)x - vp[3]) / vp[0];
    for (int y = 0; y < H; ++y) {
      const float yn = ((float)y - vp[7]) / vp[5];
      /* every triangle (pipeline.py:332-335): vectorisable */
      for (size_t t = 0; t < Tp; ++t) {
        const float c0 = (xn * i0[t] + yn * i3[t]) + i6[t];
        const float c1 = (xn * i1[t] + yn * i4[t]) + i7[t];
        const float c2 = (xn * i2[t] + yn * i5[t]) + i8[t];
        const float z = (c0 * z0[t] + c1 * z1[t]) + c2 * z2[t];
        const float zw = z * vp22 + vp23;
        float d = (c0 >= 0.f) ? zw : inf;   /* NaN-safe selects, if-converted by gcc */
        d = (c1 >= 0.f) ? d : inf;
        d = (c2 >= 0.f) ? d : inf;
        depth[t] = (cand[t] != 0.f) ? d : inf;
      }
      /* first-index argmin (shader.py:217) */
      int idx = 0;
      float bestd = depth[0];
      if (j->idx_out) {
        /* general mode: also the runner-up depth (ties count: equal depths give gap 0) */
        float second = inf;
        for (int t = 1; t < T; ++t) {
          const float d = depth[t];
          if (d < bestd) { second = bestd; bestd = d; idx = t; }
          else if (d < second) second = d;
        }
        const size_t p = (size_t)x * H + y;
        j->idx_out[p] = idx;
        j->has_out[p] = bestd < INFINITY;
        uint8_t kc = 1;
        if (!(bestd < INFINITY)) {
          const float c0 = (xn * inv[0][0] + yn * inv[3][0]) + inv[6][0];
          const float c1 = (xn * inv[1][0] + yn * inv[4][0]) + inv[7][0];
          const float c2 = (xn * inv[2][0] + yn * inv[5][0]) + inv[8][0];
          kc = j->keep0 && c0 >= 0.f && c1 >= 0.f && c2 >= 0.f;
        }
        j->kc_out[p] = kc;
        j->gap_out[p] = (second < INFINITY) ? second - bestd : inf;
        continue;
      }
      for (int t = 1; t < T; ++t)
        if (depth[t] < bestd) { bestd = depth[t]; idx = t; }
      int written = -1;
      if (bestd < INFINITY) {
        j->zb[(size_t)x * H + y] = bestd;
        written = idx;
      } else if (j->keep0) {
        /* chosen = triangle 0; DepthShader keeps iff keep0 & inside0 */
        const float c0 = (xn * inv[0][0] + yn * inv[3][0]) + inv[6][0];
        const float c1 = (xn * inv[1][0] + yn * inv[4][0]) + inv[7][0];
        const float c2 = (xn * inv[2][0] + yn * inv[5][0]) + inv[8][0];
        if (c0 >= 0.f && c1 >= 0.f && c2 >= 0.f) {
          const float z = (c0 * zc[0][0] + c1 * zc[1][0]) + c2 * zc[2][0];
          j->zb[(size_t)x * H + y] = z * vp22 + vp23;
          written = 0;
        }
      }
      if (j->tb) j->tb[(size_t)x * H + y] = written;
    }
  }
  free(depth);
  return NULL;
}

int jr_oracle_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

/*
 * Depth render of B images.  Strides are in elements; 0 broadcasts.
 * zbuffer (B,W,H) in/out, tri_id (B,W,H) out (-1 = pixel not written), may be NULL.
 * num_threads <= 0: all online cores.  Returns 0, or -1 on allocation failure.
 */
static int run_images(int B, int W, int H, int T, const float* w2c, long long w2c_bs,
                      const float* viewport, long long vp_bs, const float* position, long long pos_bs,
                      const int32_t* faces, long long faces_bs, float* zbuffer, int32_t* tri_id,
                      int32_t* idx_out, uint8_t* has_out, uint8_t* kc_out, float* gap_out, int num_threads) {
  if (num_threads <= 0) num_threads = jr_oracle_max_threads();
  if (num_threads > W) num_threads = W;
  if (num_threads > 1024) num_threads = 1024;
  /* SoA per-triangle tables (PerPrimitive, pipeline.py:49-113) */
  size_t Tp = (size_t)((T + 7) / 8) * 8;
  float* tab = (float*)aligned_alloc(64, sizeof(float) * Tp * 13 + 64);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)num_threads);
  RowJob* jobs = (RowJob*)malloc(sizeof(RowJob) * (size_t)num_threads);
  if (!tab || !th || !jobs) { free(tab); free(th); free(jobs); return -1; }
  float* inv[9];
  for (int k = 0; k < 9; ++k) inv[k] = tab + Tp * k;
  float* zc[3] = {tab + Tp * 9, tab + Tp * 10, tab + Tp * 11};
  float* candf = tab + Tp * 12; /* 1.0 = keep & front */
  int rc = 0;
  for (int b = 0; b < B; ++b) {
    const float* M = w2c + b * w2c_bs;
    const float* vp = viewport + b * vp_bs;
    const float* pos = position + b * pos_bs;
    const int32_t* f = faces + b * faces_bs;
    int keep0 = 0;
    for (int t = 0; t < T; ++t) {
      float c[3][4], A[9], iv[9];
      for (int k = 0; k < 3; ++k) to_clip(M, pos + 3 * (size_t)f[3 * t + k], c[k]);
      for (int k = 0; k < 3; ++k) { A[3 * k] = c[k][0]; A[3 * k + 1] = c[k][1]; A[3 * k + 2] = c[k][3]; }
      float det = det3(A);
      int keep = fabsf(det) > 1e-6f;
      lu_inverse3(A, iv);
      for (int k = 0; k < 9; ++k) inv[k][t] = iv[k];
      for (int k = 0; k < 3; ++k) zc[k][t] = c[k][2];
      candf[t] = (keep && det >= 0.f) ? 1.f : 0.f;
      if (t == 0) keep0 = keep;
    }
    for (size_t t = T; t < Tp; ++t) {
      for (int k = 0; k < 9; ++k) inv[k][t] = 0.f;
      for (int k = 0; k < 3; ++k) zc[k][t] = 0.f;
      candf[t] = 0.f;
    }
    for (int i = 0; i < num_threads; ++i) {
      const size_t o = (size_t)b * W * H;
      RowJob j = {W, H, T, keep0, i, num_threads, Tp, vp, inv, zc, candf,
                  zbuffer ? zbuffer + o : NULL, tri_id ? tri_id + o : NULL, 0,
                  idx_out ? idx_out + o : NULL, has_out ? has_out + o : NULL, kc_out ? kc_out + o : NULL,
                  gap_out ? gap_out + o : NULL};
      jobs[i] = j;
      if (i > 0 && pthread_create(&th[i], NULL, row_worker, &jobs[i]) != 0) {
        row_worker(&jobs[i]);       /* could not spawn: do the rows here */
        th[i] = 0;
        jobs[i].tid = -1;
      }
    }
    row_worker(&jobs[0]);
    for (int i = 1; i < num_threads; ++i)
      if (jobs[i].tid >= 0) pthread_join(th[i], NULL);
    for (int i = 0; i < num_threads; ++i)
      if (jobs[i].rc) rc = -1;
  }
  free(tab); free(th); free(jobs);
  return rc;
}

int jr_oracle_depth(int B, int W, int H, int T, const float* w2c, long long w2c_bs,
                    const float* viewport, long long vp_bs, const float* position, long long pos_bs,
                    const int32_t* faces, long long faces_bs, float* zbuffer, int32_t* tri_id,
                    int num_threads) {
  return run_images(B, W, H, T, w2c, w2c_bs, viewport, vp_bs, position, pos_bs, faces, faces_bs, zbuffer, tri_id,
                    NULL, NULL, NULL, NULL, num_threads);
}

/*
 * The visibility stage alone, for ANY built-in shader (pipeline.py:332-336 + shader.py:159-251): per pixel the
 * first-index argmin over `keep & inside & front` depths (`idx`, 0 when there is no candidate), whether a candidate
 * exists (`has`), `(keep & inside)[idx]` (`kc`) and the runner-up depth minus the best (`gap`, +inf when there is
 * no runner-up) -- the tuple oracle/jr_oracle.py::visibility computes with torch, bit for bit
 * (tests/test_oracle_c.py).  Lets the full-size configurations (960x540 x 19 980 triangles) be shaded by the torch
 * oracle on the chosen fragments.  Outputs (B,W,H).
 */
int jr_oracle_visibility(int B, int W, int H, int T, const float* w2c, long long w2c_bs,
                         const float* viewport, long long vp_bs, const float* position, long long pos_bs,
                         const int32_t* faces, long long faces_bs, int32_t* idx, uint8_t* has, uint8_t* kc,
                         float* gap, int num_threads) {
  if (!idx || !has || !kc || !gap) return -1;
  return run_images(B, W, H, T, w2c, w2c_bs, viewport, vp_bs, position, pos_bs, faces, faces_bs, NULL, NULL,
                    idx, has, kc, gap, num_threads);
}
