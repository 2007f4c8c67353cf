/* C fp32 restatement of the engine semantics.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * PARITY UNPINNED: the reference arithmetic is in the absent PyPI dependency fdtdz>=1.1.3
 * (/root/reference/setup.py:26; call site /root/reference/src/pjz/_field.py:254-269).  This file
 * restates oracle/fdtd_numpy.py (the spec) in single precision with the
 * per-cell operation order written out with explicit fmaf(), which is the order the CUDA
 * kernels in pjz_b200/csrc use -- so the GPU result can be compared BIT-FOR-BIT.
 *
 * Build: see oracle/Makefile (gcc -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp).
 * Also serves as the timed CPU baseline (bench.py cpu_baseline / --impl reference, kind "port").
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int32_t X, Y, Z;          /* full domain                                   */
  int32_t xx, yy, zz;       /* epsilon sub-volume                            */
  int32_t ox, oy, oz;       /* its offset in the domain                      */
  int32_t tt;               /* number of steps = rows of source_waveform     */
  int32_t src_axis;         /* 0,1,2                                         */
  int32_t src_pos;
  int32_t pml_lo, pml_hi;
  int32_t out_start, out_stop, out_step;
  int32_t reduced;          /* 1: E,H and dt/eps coefficient stored as fp16  */
  float dt;
} oracle_desc;

static inline float rstore(float v, int reduced) {
  return reduced ? (float)(_Float16)v : v;
}

static inline int wrap(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

/* Per-z CPML tables: b=exp(-(s/k+al)dt), a=s(b-1)/(k(s+k al)) (0 where s==0 or outside the
 * pml cells), ik=1/k (0 for k=inf).  Double math, rounded to float. */
static void cpml_tables(const oracle_desc* d, const float* kappa, const float* sigma,
                        const float* alpha, int col, float* a, float* b, float* ik) {
  for (int z = 0; z < d->Z; ++z) {
    double k = kappa[2 * z + col], s = sigma[2 * z + col], al = alpha[2 * z + col];
    double dt = (double)d->dt;
    double inv = isinf(k) ? 0.0 : 1.0 / k;
    double bb = exp(-(s * inv + al) * dt);
    int in_pml = (z < d->pml_lo) || (z >= d->Z - d->pml_hi);
    double den = k * (s + k * al);
    double aa = 0.0;
    if (s != 0.0 && in_pml && isfinite(den) && den != 0.0) aa = s * (bb - 1.0) / den;
    if (!isfinite(aa)) aa = 0.0;
    a[z] = (float)aa; b[z] = (float)bb; ik[z] = (float)inv;
  }
}

int oracle_fdtd_num_outputs(const oracle_desc* d) {
  if (d->out_step <= 0 || d->out_stop <= d->out_start) return 0;
  return (d->out_stop - d->out_start + d->out_step - 1) / d->out_step;
}

/* All pointers are host memory, C-order, z fastest:
 * eps (3,xx,yy,zz); source_field (2,1,Y,Z)|(2,X,1,Z)|(2,2,X,Y,1); waveform (tt,2);
 * mask (3,X,Y); kappa/sigma/alpha (Z,2); out (n_out,3,xx,yy,zz).
 * If steps_override >= 0 only that many steps are run (0 = set-up only; CPU-baseline timing).
 * Returns 0 on success. */
int oracle_fdtd_run(const oracle_desc* d, const float* eps, const float* source_field,
                    const float* waveform, const float* mask, const float* kappa,
                    const float* sigma, const float* alpha, float* out, int nthreads,
                    int steps_override) {
  const int X = d->X, Y = d->Y, Z = d->Z;
  const size_t P = (size_t)Y * Z, N = (size_t)X * P;
  const int rp = d->reduced;
  const float dt = d->dt;
  if (X <= 0 || Y <= 0 || Z <= 0) return 1;
  if (d->ox < 0 || d->oy < 0 || d->oz < 0 || d->ox + d->xx > X || d->oy + d->yy > Y ||
      d->oz + d->zz > Z) return 2;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  float* E = (float*)calloc(3 * N, sizeof(float));
  float* H = (float*)calloc(3 * N, sizeof(float));
  float* psiH = (float*)calloc(2 * N, sizeof(float));
  float* psiE = (float*)calloc(2 * N, sizeof(float));
  float* B = (float*)malloc(3 * N * sizeof(float));
  float* A = (float*)malloc(3 * (size_t)X * Y * sizeof(float));
  float* S = (float*)malloc(3 * (size_t)X * Y * sizeof(float));
  float* tab = (float*)malloc(6 * (size_t)Z * sizeof(float));
  if (!E || !H || !psiH || !psiE || !B || !A || !S || !tab) return 3;
  float *ae = tab, *be = tab + Z, *ike = tab + 2 * Z, *ah = tab + 3 * Z, *bh = tab + 4 * Z,
        *ikh = tab + 5 * Z;
  cpml_tables(d, kappa, sigma, alpha, 0, ae, be, ike);
  cpml_tables(d, kappa, sigma, alpha, 1, ah, bh, ikh);
  for (size_t i = 0; i < 3 * (size_t)X * Y; ++i) {
    double s = mask[i], h = (double)dt / 2;
    A[i] = (float)((1 - s * h) / (1 + s * h));
    S[i] = (float)(1 / (1 + s * h));
  }
  /* B = (dt/eps_ext) * S, eps edge-replicated outside the sub-volume. */
#pragma omp parallel for collapse(2) schedule(static)
  for (int c = 0; c < 3; ++c)
    for (int x = 0; x < X; ++x) {
      int ex = x - d->ox; ex = ex < 0 ? 0 : (ex >= d->xx ? d->xx - 1 : ex);
      for (int y = 0; y < Y; ++y) {
        int ey = y - d->oy; ey = ey < 0 ? 0 : (ey >= d->yy ? d->yy - 1 : ey);
        float s = S[((size_t)c * X + x) * Y + y];
        const float* er = eps + (((size_t)c * d->xx + ex) * d->yy + ey) * d->zz;
        float* br = B + (size_t)c * N + (size_t)x * P + (size_t)y * Z;
        for (int z = 0; z < Z; ++z) {
          int ez = z - d->oz; ez = ez < 0 ? 0 : (ez >= d->zz ? d->zz - 1 : ez);
          br[z] = rstore((dt / er[ez]) * s, rp);
        }
      }
    }
  float *Ex = E, *Ey = E + N, *Ez = E + 2 * N, *Hx = H, *Hy = H + N, *Hz = H + 2 * N;
  float *pHx = psiH, *pHy = psiH + N, *pEx = psiE, *pEy = psiE + N;
  const int tt = steps_override >= 0 && steps_override < d->tt ? steps_override : d->tt;
  const int nout = oracle_fdtd_num_outputs(d);
  int oi = 0;
  for (int n = 0; n < tt; ++n) {
    /* (1) H update: forward differences, x-y periodic, E[z=Z] := 0. */
#pragma omp parallel for collapse(2) schedule(static)
    for (int x = 0; x < X; ++x)
      for (int y = 0; y < Y; ++y) {
        const size_t o = (size_t)x * P + (size_t)y * Z;
        const size_t ox1 = (size_t)wrap(x + 1, X) * P + (size_t)y * Z;
        const size_t oy1 = (size_t)x * P + (size_t)wrap(y + 1, Y) * Z;
        for (int z = 0; z < Z; ++z) {
          float ey = Ey[o + z], ex = Ex[o + z], ez = Ez[o + z];
          float dzEy = (z + 1 < Z ? Ey[o + z + 1] : 0.0f) - ey;
          float dzEx = (z + 1 < Z ? Ex[o + z + 1] : 0.0f) - ex;
          float px = fmaf(bh[z], pHx[o + z], ah[z] * dzEy);
          float py = fmaf(bh[z], pHy[o + z], ah[z] * dzEx);
          pHx[o + z] = px; pHy[o + z] = py;
          float cx = (Ez[oy1 + z] - ez) - fmaf(dzEy, ikh[z], px);
          float cy = fmaf(dzEx, ikh[z], py) - (Ez[ox1 + z] - ez);
          float cz = (Ey[ox1 + z] - ey) - (Ex[oy1 + z] - ex);
          Hx[o + z] = rstore(fmaf(-dt, cx, Hx[o + z]), rp);
          Hy[o + z] = rstore(fmaf(-dt, cy, Hy[o + z]), rp);
          Hz[o + z] = rstore(fmaf(-dt, cz, Hz[o + z]), rp);
        }
      }
    /* (2) E update: backward differences, H[z=-1] := 0; (3) source; fp16 rounding last. */
    const float w0 = waveform[2 * n], w1 = waveform[2 * n + 1];
#pragma omp parallel for collapse(2) schedule(static)
    for (int x = 0; x < X; ++x)
      for (int y = 0; y < Y; ++y) {
        const size_t o = (size_t)x * P + (size_t)y * Z;
        const size_t ox1 = (size_t)wrap(x - 1, X) * P + (size_t)y * Z;
        const size_t oy1 = (size_t)x * P + (size_t)wrap(y - 1, Y) * Z;
        const float a0 = A[((size_t)0 * X + x) * Y + y], a1 = A[((size_t)1 * X + x) * Y + y],
                    a2 = A[((size_t)2 * X + x) * Y + y];
        for (int z = 0; z < Z; ++z) {
          float hx = Hx[o + z], hy = Hy[o + z], hz = Hz[o + z];
          float dzHy = hy - (z > 0 ? Hy[o + z - 1] : 0.0f);
          float dzHx = hx - (z > 0 ? Hx[o + z - 1] : 0.0f);
          float px = fmaf(be[z], pEx[o + z], ae[z] * dzHy);
          float py = fmaf(be[z], pEy[o + z], ae[z] * dzHx);
          pEx[o + z] = px; pEy[o + z] = py;
          float cx = (hz - Hz[oy1 + z]) - fmaf(dzHy, ike[z], px);
          float cy = fmaf(dzHx, ike[z], py) - (hz - Hz[ox1 + z]);
          float cz = (hy - Hy[ox1 + z]) - (hx - Hx[oy1 + z]);
          float e0 = fmaf(B[o + z], cx, a0 * Ex[o + z]);
          float e1 = fmaf(B[N + o + z], cy, a1 * Ey[o + z]);
          float e2 = fmaf(B[2 * N + o + z], cz, a2 * Ez[o + z]);
          Ex[o + z] = e0; Ey[o + z] = e1; Ez[o + z] = e2;
        }
        /* (3) source: channel 0 then channel 1, each a single fmaf */
        if (d->src_axis == 0) {
          for (int ch = 0; ch < 2; ++ch)
            if (x == wrap(d->src_pos - ch, X)) {
              const float w = ch ? w1 : w0;
              for (int z = 0; z < Z; ++z) {
                Ey[o + z] = fmaf(w, source_field[(size_t)0 * P + (size_t)y * Z + z], Ey[o + z]);
                Ez[o + z] = fmaf(w, source_field[(size_t)1 * P + (size_t)y * Z + z], Ez[o + z]);
              }
            }
        } else if (d->src_axis == 1) {
          for (int ch = 0; ch < 2; ++ch)
            if (y == wrap(d->src_pos - ch, Y)) {
              const float w = ch ? w1 : w0;
              for (int z = 0; z < Z; ++z) {
                Ex[o + z] = fmaf(w, source_field[((size_t)0 * X + x) * Z + z], Ex[o + z]);
                Ez[o + z] = fmaf(w, source_field[((size_t)1 * X + x) * Z + z], Ez[o + z]);
              }
            }
        } else {
          const size_t XY = (size_t)X * Y, xy = (size_t)x * Y + y;
          const int z = d->src_pos;
          float e0 = Ex[o + z], e1 = Ey[o + z];
          e0 = fmaf(w0, source_field[0 * XY + xy], e0);
          e1 = fmaf(w0, source_field[1 * XY + xy], e1);
          e0 = fmaf(w1, source_field[2 * XY + xy], e0);
          e1 = fmaf(w1, source_field[3 * XY + xy], e1);
          Ex[o + z] = e0; Ey[o + z] = e1;
        }
        if (rp)
          for (int z = 0; z < Z; ++z) {
            Ex[o + z] = rstore(Ex[o + z], 1); Ey[o + z] = rstore(Ey[o + z], 1);
            Ez[o + z] = rstore(Ez[o + z], 1);
          }
      }
    /* (4) snapshot */
    if (out && oi < nout && n == d->out_start + oi * d->out_step) {
      for (int c = 0; c < 3; ++c)
        for (int x = 0; x < d->xx; ++x)
          for (int y = 0; y < d->yy; ++y)
            memcpy(out + ((((size_t)oi * 3 + c) * d->xx + x) * d->yy + y) * d->zz,
                   E + (size_t)c * N + (size_t)(x + d->ox) * P + (size_t)(y + d->oy) * Z + d->oz,
                   (size_t)d->zz * sizeof(float));
      ++oi;
    }
  }
  free(E); free(H); free(psiH); free(psiE); free(B); free(A); free(S); free(tab);
  return 0;
}

int oracle_fdtd_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}


/* Fused frequency projection (engine extension, include/b200fdtd.h `proj_rows`): out[r] =
 * sum_s w[r][s] * snaps[s], accumulated from zero in snapshot order with one fmaf per term --
 * the order the CUDA kernels use when they accumulate in place at each output step.
 * (Restates the pinv einsum of /root/reference/src/pjz/_field.py:276-277 as a running sum.) */
void oracle_fdtd_project(const float* snaps, const float* w, int rows, int nout, size_t n,
                         float* out) {
  for (int r = 0; r < rows; ++r) {
#pragma omp parallel for
    for (long long i = 0; i < (long long)n; ++i) {
      float acc = 0.f;
      for (int s = 0; s < nout; ++s) acc = fmaf(w[(size_t)r * nout + s], snaps[(size_t)s * n + i], acc);
      out[(size_t)r * n + i] = acc;
    }
  }
}
