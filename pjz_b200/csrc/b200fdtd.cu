// libb200fdtd.so -- C ABI (include/b200fdtd.h) and host-side driver of the B200 FDTD engine.
//
// Replaces the engine behind fdtdz_jax.fdtdz (/root/reference/src/pjz/_field.py:254-269; the
// reference implementation is the absent PyPI package fdtdz>=1.1.3, /root/reference/setup.py:26).
// Everything here is stream-ordered: coefficient preparation runs as small kernels on the
// caller's stream, so b200fdtd_run never synchronises with the host.
#include "../../include/b200fdtd.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <limits.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <new>
#include <string>

#include "fdtd_common.cuh"
#include "kernels_systolic.cuh"
#include "kernels_systolic2.cuh"
#include "kernels_lean.cuh"
#include "kernels_lean16.cuh"
#include "kernels_twopass.cuh"
#include "postproc.cuh"
#include "render.cuh"

namespace b200 {

static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(B200FDTD_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                   \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int num_outputs(const b200fdtd_desc* d) {
  if (d->out_stop <= d->out_start) return 0;
  return (d->out_stop - d->out_start + d->out_step - 1) / d->out_step;
}

static int validate(const b200fdtd_desc* d) {
  if (!d) return fail(B200FDTD_EINVAL, "desc is NULL");
  if (d->struct_bytes != sizeof(b200fdtd_desc) || d->abi_version != B200FDTD_ABI_VERSION)
    return fail(B200FDTD_EINVAL, "descriptor ABI mismatch (struct_bytes=%u, abi_version=%u)",
                d->struct_bytes, d->abi_version);
  if (d->X < 1 || d->Y < 1 || d->Z < 1)
    return fail(B200FDTD_EINVAL, "domain must be positive, got (%d,%d,%d)", d->X, d->Y, d->Z);
  if (d->xx < 1 || d->yy < 1 || d->zz < 1 || d->off_x < 0 || d->off_y < 0 || d->off_z < 0 ||
      d->off_x + d->xx > d->X || d->off_y + d->yy > d->Y || d->off_z + d->zz > d->Z)
    return fail(B200FDTD_EINVAL,
                "epsilon sub-volume (%d,%d,%d) at offset (%d,%d,%d) does not fit domain (%d,%d,%d)",
                d->xx, d->yy, d->zz, d->off_x, d->off_y, d->off_z, d->X, d->Y, d->Z);
  if (d->tt < 0) return fail(B200FDTD_EINVAL, "tt must be >= 0");
  if (d->source_axis < 0 || d->source_axis > 2)
    return fail(B200FDTD_EINVAL, "source_axis must be 0, 1 or 2");
  const int ext = d->source_axis == 0 ? d->X : (d->source_axis == 1 ? d->Y : d->Z);
  if (d->source_position < 0 || d->source_position >= ext)
    return fail(B200FDTD_EINVAL, "source_position %d outside [0,%d)", d->source_position, ext);
  if (d->pml_lo < 0 || d->pml_hi < 0 || d->pml_lo + d->pml_hi > d->Z)
    return fail(B200FDTD_EINVAL, "pml_widths (%d,%d) do not fit Z=%d", d->pml_lo, d->pml_hi, d->Z);
  if (d->out_step < 1) return fail(B200FDTD_EINVAL, "output_steps step must be >= 1");
  // every snapshot step must be a step the run performs: the last one is
  // out_start + (n_out - 1) * out_step (oracle/fdtd_numpy.py rejects outs[-1] >= tt the same way)
  if (d->out_stop > d->out_start &&
      (d->out_start < 0 ||
       (long long)d->out_start + (long long)(num_outputs(d) - 1) * d->out_step >= d->tt))
    return fail(B200FDTD_EINVAL, "output_steps (%d,%d,%d) outside [0,tt=%d)", d->out_start,
                d->out_stop, d->out_step, d->tt);
  if (!(d->dt > 0.f) || !isfinite(d->dt)) return fail(B200FDTD_EINVAL, "dt must be > 0");
  if (d->kernel < 0 || d->kernel > 5 || d->kernel == B200FDTD_KERNEL_RESERVED4)
    return fail(B200FDTD_EINVAL, "unknown kernel %d", d->kernel);
  if (d->cols < 0 || d->cols > 2) return fail(B200FDTD_EINVAL, "cols must be 0, 1 or 2");
  if (d->proj_rows < 0 || d->proj_rows > 64)
    return fail(B200FDTD_EINVAL, "proj_rows must be in [0, 64], got %d", d->proj_rows);
  if (d->proj_rows > 0 && d->out_stop <= d->out_start)
    return fail(B200FDTD_EINVAL, "fused projection needs at least one output step");
  return B200FDTD_OK;
}

static Geom make_geom(const b200fdtd_desc* d) {
  Geom g;
  const int VW = d->use_reduced_precision ? 8 : 4;
  g.X = d->X; g.Y = d->Y; g.Z = d->Z;
  g.Zp = (d->Z + VW - 1) / VW * VW;
  g.Zq = g.Zp / VW;
  g.nlo = (d->pml_lo + VW - 1) / VW;
  g.hi0 = d->pml_hi > 0 ? (d->Z - d->pml_hi) / VW : g.Zq;
  if (g.hi0 < g.nlo) g.hi0 = g.nlo;
  g.npg = g.nlo + (g.Zq - g.hi0);
  g.xx = d->xx; g.yy = d->yy; g.zz = d->zz;
  g.ox = d->off_x; g.oy = d->off_y; g.oz = d->off_z;
  g.src_axis = d->source_axis; g.src_pos = d->source_position;
  g.out_start = d->out_start; g.out_stop = d->out_stop; g.out_step = d->out_step;
  g.proj_rows = d->proj_rows;
  g.n_out = num_outputs(d);
  g.tt = d->tt;
  g.n0 = 0;
  g.ylo = 0; g.yhi = d->Y;
  g.dt = d->dt;
  g.P = (long long)g.Y * g.Zp;
  g.N = (long long)g.X * g.P;
  return g;
}

// ---- launch plan ---------------------------------------------------------------------------

struct Plan {
  int kernel;            // resolved B200FDTD_KERNEL_*
  int depth;             // prefetch distance (SYSTOLIC_ASYNC)
  SystolicCfg sys;       // valid for the systolic kernels
};

static bool is_systolic(int k) {
  return k == B200FDTD_KERNEL_SYSTOLIC || k == B200FDTD_KERNEL_SYSTOLIC_ASYNC ||
         k == B200FDTD_KERNEL_SYSTOLIC_LEAN;
}

static int device_props(int* sms, int* l2_bytes) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  CUDA_TRY(cudaDeviceGetAttribute(l2_bytes, cudaDevAttrL2CacheSize, dev));
  return B200FDTD_OK;
}

template <typename T>
static bool configure_async(const Geom& g, const b200fdtd_desc* d, int depth, int sms, int l2,
                            SystolicCfg* cfg, std::string* why) {
  const int K = d->cols == 1 ? 1 : 2;            // columns per compute thread (0 = default 2)
  switch (depth) {
    case 1: return systolic2_configure_d<T, 1>(g, d->tile_y, d->stages, d->threads, sms, l2, K, cfg, why);
    case 2: return systolic2_configure_d<T, 2>(g, d->tile_y, d->stages, d->threads, sms, l2, K, cfg, why);
    case 3: return systolic2_configure_d<T, 3>(g, d->tile_y, d->stages, d->threads, sms, l2, K, cfg, why);
  }
  *why = "prefetch must be 1, 2 or 3";
  return false;
}

// AUTO: the warp-per-column-pair kernel when the geometry allows it (fp32, 32 z-vectors), else the
// cp.async-staged systolic kernel (prefetch distance 1 measured fastest: a deeper ring only
// shrinks the tile), else the register-staged one, else the per-step kernels.
template <typename T>
static int make_plan_t(const b200fdtd_desc* d, const Geom& g, Plan* plan) {
  plan->kernel = d->kernel;
  plan->depth = 0;
  if (d->kernel == B200FDTD_KERNEL_TWOPASS) return B200FDTD_OK;
  int sms = 0, l2 = 0;
  int rc = device_props(&sms, &l2);
  if (rc) return rc;
  std::string why;
  if (d->kernel == B200FDTD_KERNEL_SYSTOLIC_LEAN) {
    bool ok = lean_configure(g, sizeof(T) == 2, d->tile_y, d->stages, sms, l2, &plan->sys, &why);
    if (!ok && g.Zq <= kL16ZR) {                   // short columns / fp16 storage: half-warp variant
      std::string why16;
      ok = lean16_configure<T>(g, d->tile_y, d->stages, sms, l2, &plan->sys, &why16);
      if (!ok) why += "; half-warp variant: " + why16;
    }
    if (!ok)
      return fail(B200FDTD_EUNSUPPORTED, "systolic_lean kernel unavailable: %s", why.c_str());
    plan->depth = 1;
    return B200FDTD_OK;
  }
  if (d->kernel == B200FDTD_KERNEL_AUTO &&
      lean_configure(g, sizeof(T) == 2, d->tile_y, d->stages, sms, l2, &plan->sys, &why)) {
    plan->kernel = B200FDTD_KERNEL_SYSTOLIC_LEAN;
    plan->depth = 1;
    return B200FDTD_OK;
  }
  // Columns of at most 16 vectors (everything fdtd-z itself accepts): the sub-warp lean kernel
  // where most of its lanes carry a vector.  Measured on 256x256 in x-y (Gcell/s, lean16 vs the
  // cp.async kernel, profiles/r01_lean16_pjz_geometries_final.txt): fp16 Z=128 119 vs 86, fp16
  // Z=96 (pjz's default; 12 of 16 lanes busy) 85 vs 82, fp16 Z=64 (8 lanes per column) 109 vs 77,
  // fp32 Z=64 83 vs 73; fp32 columns of <= 12 vectors are slower (Z=48 62 vs 67, Z=32 67 vs 70)
  // and stay on the cp.async kernel.
  if (d->kernel == B200FDTD_KERNEL_AUTO && g.Zq <= kL16ZR) {
    const bool wide = sizeof(T) == 2 ? ((g.Zq >= 12) || (g.Zq >= 7 && g.Zq <= 8)) : g.Zq >= 15;
    if (wide && lean16_configure<T>(g, d->tile_y, d->stages, sms, l2, &plan->sys, &why)) {
      plan->kernel = B200FDTD_KERNEL_SYSTOLIC_LEAN;
      plan->depth = 1;
      return B200FDTD_OK;
    }
  }
  if (d->kernel == B200FDTD_KERNEL_AUTO || d->kernel == B200FDTD_KERNEL_SYSTOLIC_ASYNC) {
    const int depth = d->prefetch > 0 ? d->prefetch : 1;
    if (configure_async<T>(g, d, depth, sms, l2, &plan->sys, &why)) {
      // Short x extent: the stages of the pipeline trail each other by ~depth+6 planes around a
      // ring of X planes, so only X/(depth+6) of them can run at once.  Rather than leaving SMs
      // to stages that only wait, cap the stage count and cut the domain into narrower y-tiles.
      const int cap = g.X / (depth + 6) > 1 ? g.X / (depth + 6) : 1;
      if (d->tile_y == 0 && d->stages == 0 && plan->sys.stages > cap && plan->sys.tile_y > 2) {
        b200fdtd_desc d2 = *d;
        const int want_tiles = (sms + cap - 1) / cap;
        int tile = (g.Y + want_tiles - 1) / want_tiles;
        if (tile < 2) tile = 2;
        if (tile < plan->sys.tile_y) {
          d2.tile_y = tile;
          d2.stages = cap;
          SystolicCfg alt;
          std::string why2;
          if (configure_async<T>(g, &d2, depth, sms, l2, &alt, &why2) && alt.stages <= cap &&
              (long long)alt.stages * alt.ntiles > (long long)cap * plan->sys.ntiles)
            plan->sys = alt;
        }
      }
      plan->kernel = B200FDTD_KERNEL_SYSTOLIC_ASYNC;
      plan->depth = depth;
      return B200FDTD_OK;
    }
    if (d->kernel == B200FDTD_KERNEL_SYSTOLIC_ASYNC)
      return fail(B200FDTD_EUNSUPPORTED, "systolic_async kernel unavailable: %s", why.c_str());
  }
  if (systolic_configure<T>(g, d->tile_y, d->stages, d->threads, sms, l2, &plan->sys, &why)) {
    plan->kernel = B200FDTD_KERNEL_SYSTOLIC;
    return B200FDTD_OK;
  }
  if (d->kernel == B200FDTD_KERNEL_SYSTOLIC)
    return fail(B200FDTD_EUNSUPPORTED, "systolic kernel unavailable: %s", why.c_str());
  plan->kernel = B200FDTD_KERNEL_TWOPASS;
  return B200FDTD_OK;
}

static int make_plan(const b200fdtd_desc* d, const Geom& g, Plan* plan) {
  if (g.yhi - g.ylo != g.Y && (d->use_reduced_precision || g.Zq != 32))
    return fail(B200FDTD_EUNSUPPORTED,
                "y-slab sessions need the warp-per-column-pair kernel (fp32, 125 <= Z <= 128)");
  return d->use_reduced_precision ? make_plan_t<__half>(d, g, plan) : make_plan_t<float>(d, g, plan);
}

// ---- workspace -----------------------------------------------------------------------------

struct Workspace {
  size_t fields;   // E[3], H[3]        (zeroed)
  size_t fields2;  // E2[3], H2[3]      (zeroed; systolic only)
  size_t psi;      // psiH[2], psiE[2]  (zeroed)
  size_t sync;     // progress counters + status (zeroed)
  size_t zero_end; // end of the zero-initialised prefix
  size_t B, A, A4, S4, S, tab;
  size_t total;
};

static Workspace carve(const Geom& g, bool reduced, bool systolic, const SystolicCfg* sys) {
  Workspace w;
  const size_t el = reduced ? 2 : 4;
  const size_t VW = reduced ? 8 : 4;
  size_t o = 0;
  w.fields = o;  o = align_up(o + 6 * (size_t)g.N * el, 256);
  w.fields2 = o; if (systolic) o = align_up(o + 6 * (size_t)g.N * el, 256);
  w.psi = o;     o = align_up(o + (systolic ? 6 : 4) * (size_t)g.X * g.Y * g.npg * VW * sizeof(float), 256);
  w.sync = o;    o = align_up(o + (systolic ? systolic_sync_bytes(*sys) : 0) + 256, 256);
  w.zero_end = o;
  w.B = o;       o = align_up(o + 3 * (size_t)g.N * el, 256);
  w.A = o;       o = align_up(o + 3 * (size_t)g.X * g.Y * sizeof(float), 256);
  w.A4 = o;      o = align_up(o + 4 * (size_t)g.X * g.Y * sizeof(float), 256);
  w.S4 = o;      o = align_up(o + 4 * (size_t)g.X * g.Y * sizeof(float), 256);
  w.S = o;       o = align_up(o + 3 * (size_t)g.X * g.Y * sizeof(float), 256);
  w.tab = o;     o = align_up(o + 6 * (size_t)g.Zp * sizeof(float), 256);
  w.total = o;
  return w;
}

// ---- coefficient preparation kernels ------------------------------------------------------------

// CPML tables, same double-precision formulas as oracle/fdtd_c.c:cpml_tables.
__global__ void prep_tables_kernel(Geom g, const float* __restrict__ kappa,
                                   const float* __restrict__ sigma,
                                   const float* __restrict__ alpha, int pml_lo, int pml_hi,
                                   float* __restrict__ tab) {
  for (int z = blockIdx.x * blockDim.x + threadIdx.x; z < g.Zp; z += gridDim.x * blockDim.x) {
    for (int col = 0; col < 2; ++col) {
      float fa = 0.f, fb = 0.f, fik = 0.f;
      if (z < g.Z) {
        const double k = kappa[2 * z + col], s = sigma[2 * z + col], al = alpha[2 * z + col];
        const double dt = (double)g.dt;
        const double inv = isinf(k) ? 0.0 : 1.0 / k;
        const double bb = exp(-(s * inv + al) * dt);
        const bool in_pml = (z < pml_lo) || (z >= g.Z - pml_hi);
        const double den = k * (s + k * al);
        double aa = 0.0;
        if (s != 0.0 && in_pml && isfinite(den) && den != 0.0) aa = s * (bb - 1.0) / den;
        if (!isfinite(aa)) aa = 0.0;
        fa = (float)aa; fb = (float)bb; fik = (float)inv;
      }
      tab[(3 * col + 0) * g.Zp + z] = fa;
      tab[(3 * col + 1) * g.Zp + z] = fb;
      tab[(3 * col + 2) * g.Zp + z] = fik;
    }
  }
}

__global__ void prep_absorber_kernel(Geom g, const float* __restrict__ mask,
                                     float* __restrict__ A, float* __restrict__ A4,
                                     float* __restrict__ S) {
  const size_t XY = (size_t)g.X * g.Y, n = 3 * XY;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const double s = mask[i], h = (double)g.dt / 2;
    const float a = (float)((1 - s * h) / (1 + s * h));
    A[i] = a;
    A4[4 * (i % XY) + i / XY] = a;
    if (i < XY) A4[4 * i + 3] = 0.f;
    S[i] = (float)(1 / (1 + s * h));
  }
}

// z-plane source (2,2,X,Y,1) repacked per column so that a kernel fetches it with one 16-byte
// load: S4[x][y] = (ch0 Ex, ch0 Ey, ch1 Ex, ch1 Ey).
__global__ void prep_zsource_kernel(Geom g, const float* __restrict__ src, float* __restrict__ S4) {
  const size_t XY = (size_t)g.X * g.Y;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < XY;
       i += (size_t)gridDim.x * blockDim.x) {
    S4[4 * i + 0] = src[i]; S4[4 * i + 1] = src[XY + i];
    S4[4 * i + 2] = src[2 * XY + i]; S4[4 * i + 3] = src[3 * XY + i];
  }
}

// B_c = (dt / eps_c) * S_c with eps edge-replicated outside the sub-volume; 0 in the z padding.
template <typename T>
__global__ void prep_b_kernel(Geom g, const float* __restrict__ eps, const float* __restrict__ S,
                              T* __restrict__ B) {
  const size_t n = 3 * (size_t)g.N;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const int z = (int)(i % g.Zp);
    size_t r = i / g.Zp;
    const int y = (int)(r % g.Y); r /= g.Y;
    const int x = (int)(r % g.X);
    const int c = (int)(r / g.X);
    float v = 0.f;
    if (z < g.Z) {
      int ex = x - g.ox; ex = ex < 0 ? 0 : (ex >= g.xx ? g.xx - 1 : ex);
      int ey = y - g.oy; ey = ey < 0 ? 0 : (ey >= g.yy ? g.yy - 1 : ey);
      int ez = z - g.oz; ez = ez < 0 ? 0 : (ez >= g.zz ? g.zz - 1 : ez);
      const float e = eps[(((size_t)c * g.xx + ex) * g.yy + ey) * g.zz + ez];
      const float s = S[((size_t)c * g.X + x) * g.Y + y];
      v = __fmul_rn(__fdiv_rn(g.dt, e), s);
    }
    if constexpr (sizeof(T) == 4) B[i] = v;
    else B[i] = __float2half_rn(v);
  }
}

// ---- the run ----------------------------------------------------------------------------------

// Binds the workspace regions and the caller's arrays into `p`, zeroes the state and runs the
// coefficient-preparation kernels (all stream-ordered).
template <typename T>
static int prepare_typed(const b200fdtd_desc* d, const Geom& g, const Workspace& w,
                         const void* const* in, void* const* out, char* ws, cudaStream_t st,
                         Ptrs<T>* pp) {
  constexpr int VW = VecTraits<T>::VW;
  Ptrs<T>& p = *pp;
  T* fields = reinterpret_cast<T*>(ws + w.fields);
  T* fields2 = reinterpret_cast<T*>(ws + w.fields2);
  T* B = reinterpret_cast<T*>(ws + w.B);
  float* psi = reinterpret_cast<float*>(ws + w.psi);
  const size_t psi_n = (size_t)g.X * g.Y * g.npg * VW;
  for (int c = 0; c < 3; ++c) {
    p.E[c] = fields + (size_t)c * g.N;
    p.H[c] = fields + (size_t)(3 + c) * g.N;
    p.E2[c] = fields2 + (size_t)c * g.N;
    p.H2[c] = fields2 + (size_t)(3 + c) * g.N;
    p.B[c] = B + (size_t)c * g.N;
  }
  p.A = reinterpret_cast<float*>(ws + w.A);
  p.A4 = reinterpret_cast<float*>(ws + w.A4);
  p.S4 = reinterpret_cast<float*>(ws + w.S4);
  p.tab = reinterpret_cast<float*>(ws + w.tab);
  p.psiH[0] = psi; p.psiH[1] = psi + psi_n; p.psiE[0] = psi + 2 * psi_n; p.psiE[1] = psi + 3 * psi_n;
  p.psiH2[0] = psi + 4 * psi_n; p.psiH2[1] = psi + 5 * psi_n;   // carved only for the systolic kernel
  for (int c = 0; c < 3; ++c) {
    p.Es[0][c] = p.E[c]; p.Es[1][c] = p.E2[c];
    p.Hs[0][c] = p.H[c]; p.Hs[1][c] = p.H2[c];
  }
  for (int c = 0; c < 2; ++c) { p.psiHs[0][c] = p.psiH[c]; p.psiHs[1][c] = p.psiH2[c]; }
  p.src = static_cast<const float*>(in[B200FDTD_IN_SOURCE_FIELD]);
  p.wave = static_cast<const float*>(in[B200FDTD_IN_SOURCE_WAVEFORM]);
  p.out = static_cast<float*>(out[0]);
  p.proj = d->proj_rows > 0 ? static_cast<const float*>(in[B200FDTD_IN_PROJECTION]) : nullptr;
  if (d->proj_rows > 0)   // the accumulators start from zero
    CUDA_TRY(cudaMemsetAsync(p.out, 0, (size_t)d->proj_rows * 3 * d->xx * d->yy * d->zz * sizeof(float), st));

  CUDA_TRY(cudaMemsetAsync(ws, 0, w.zero_end, st));
  float* S = reinterpret_cast<float*>(ws + w.S);
  prep_tables_kernel<<<1, 256, 0, st>>>(
      g, static_cast<const float*>(in[B200FDTD_IN_PML_KAPPA]),
      static_cast<const float*>(in[B200FDTD_IN_PML_SIGMA]),
      static_cast<const float*>(in[B200FDTD_IN_PML_ALPHA]), d->pml_lo, d->pml_hi,
      const_cast<float*>(p.tab));
  prep_absorber_kernel<<<256, 256, 0, st>>>(
      g, static_cast<const float*>(in[B200FDTD_IN_ABSORPTION_MASK]), const_cast<float*>(p.A),
      const_cast<float*>(p.A4), S);
  if (g.src_axis == 2)
    prep_zsource_kernel<<<256, 256, 0, st>>>(g, p.src, const_cast<float*>(p.S4));
  prep_b_kernel<T><<<148 * 8, 256, 0, st>>>(
      g, static_cast<const float*>(in[B200FDTD_IN_EPSILON]), S, B);
  CUDA_TRY(cudaGetLastError());
  return B200FDTD_OK;
}

template <typename T>
static int run_typed(const b200fdtd_desc* d, const Geom& g, const Plan& plan, const Workspace& w,
                     const void* const* in, void* const* out, char* ws, cudaStream_t st) {
  Ptrs<T> p;
  int prc = prepare_typed<T>(d, g, w, in, out, ws, st, &p);
  if (prc) return prc;
  if (g.tt == 0) return B200FDTD_OK;
  if (is_systolic(plan.kernel)) {
    unsigned* sync = reinterpret_cast<unsigned*>(ws + w.sync);
    int rc;
    if (plan.kernel == B200FDTD_KERNEL_SYSTOLIC) rc = systolic_launch<T>(g, p, plan.sys, sync, st);
    else if (plan.kernel == B200FDTD_KERNEL_SYSTOLIC_LEAN) {
      if (plan.sys.cols == 16) rc = lean16_launch<T>(g, p, plan.sys, sync, st);
      else if constexpr (sizeof(T) == 4) rc = lean_launch(g, p, plan.sys, sync, st);
      else rc = (int)cudaErrorInvalidValue;
    }
    else if (plan.depth == 1) rc = systolic2_launch_d<T, 1>(g, p, plan.sys, sync, st);
    else if (plan.depth == 2) rc = systolic2_launch_d<T, 2>(g, p, plan.sys, sync, st);
    else rc = systolic2_launch_d<T, 3>(g, p, plan.sys, sync, st);
    if (rc != 0)
      return fail(B200FDTD_ECUDA, "systolic launch failed: %s",
                  cudaGetErrorString((cudaError_t)rc));
    return B200FDTD_OK;
  }
  const int per_plane = g.Y * g.Zq;
  dim3 grid((per_plane + kTwoPassThreads - 1) / kTwoPassThreads, g.X);
  for (int n = 0; n < g.tt; ++n) {
    twopass_h_kernel<T><<<grid, kTwoPassThreads, 0, st>>>(g, p);
    twopass_e_kernel<T><<<grid, kTwoPassThreads, 0, st>>>(g, p, n);
  }
  CUDA_TRY(cudaGetLastError());
  return B200FDTD_OK;
}

static int run_impl(const b200fdtd_desc* d, const void* const* in, void* const* out, void* ws_in,
                    size_t ws_bytes, cudaStream_t st) {
  int rc = validate(d);
  if (rc) return rc;
  if (!in || !out) return fail(B200FDTD_EINVAL, "inputs/outputs is NULL");
  for (int i = 0; i < B200FDTD_NUM_INPUTS; ++i)
    if (!in[i]) return fail(B200FDTD_EINVAL, "inputs[%d] is NULL", i);
  if (num_outputs(d) > 0 && !out[0]) return fail(B200FDTD_EINVAL, "outputs[0] is NULL");
  if (d->proj_rows > 0 && !in[B200FDTD_IN_PROJECTION])
    return fail(B200FDTD_EINVAL, "inputs[%d] (projection weights) is NULL", B200FDTD_IN_PROJECTION);
  const Geom g = make_geom(d);
  if ((long long)g.Y * g.Zq > INT_MAX / 8) return fail(B200FDTD_EUNSUPPORTED, "plane too large");
  if (g.X > 65535 && d->kernel == B200FDTD_KERNEL_TWOPASS)
    return fail(B200FDTD_EUNSUPPORTED, "two-pass kernel supports X <= 65535");
  Plan plan;
  rc = make_plan(d, g, &plan);
  if (rc) return rc;
  if (plan.kernel == B200FDTD_KERNEL_TWOPASS && g.X > 65535)
    return fail(B200FDTD_EUNSUPPORTED, "two-pass kernel supports X <= 65535");
  const Workspace w = carve(g, d->use_reduced_precision != 0,
                            is_systolic(plan.kernel), &plan.sys);
  char* ws = static_cast<char*>(ws_in);
  bool own = false;
  if (!ws) {
    CUDA_TRY(cudaMallocAsync((void**)&ws, w.total, st));
    own = true;
  } else {
    if (ws_bytes < w.total)
      return fail(B200FDTD_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total,
                  ws_bytes);
    if ((uintptr_t)ws % 256 != 0) return fail(B200FDTD_EINVAL, "workspace must be 256-byte aligned");
  }
  rc = d->use_reduced_precision ? run_typed<__half>(d, g, plan, w, in, out, ws, st)
                                : run_typed<float>(d, g, plan, w, in, out, ws, st);
  if (own) cudaFreeAsync(ws, st);
  return rc;
}

// Private stream-ordered memory pool of b200fdtd_run_host, one per device, created on first use.
// Its release threshold is unlimited so that freed blocks stay cached between calls; the
// device's default pool (and therefore every other user of cudaMallocAsync in the process) is
// left alone.  b200fdtd_host_pool_trim() hands the cached blocks back to the driver.
static std::mutex g_pool_mu;
static cudaMemPool_t g_pools[64] = {nullptr};

static int host_pool(int device, cudaMemPool_t* out) {
  if (device < 0 || device >= 64) return fail(B200FDTD_EINVAL, "device %d out of range", device);
  std::lock_guard<std::mutex> lock(g_pool_mu);
  if (!g_pools[device]) {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof props);
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    cudaMemPool_t pool = nullptr;
    CUDA_TRY(cudaMemPoolCreate(&pool, &props));
    unsigned long long keep = ~0ull;
    cudaError_t e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    if (e != cudaSuccess) {
      cudaMemPoolDestroy(pool);
      return fail(B200FDTD_ECUDA, "cudaMemPoolSetAttribute failed: %s", cudaGetErrorString(e));
    }
    g_pools[device] = pool;
  }
  *out = g_pools[device];
  return B200FDTD_OK;
}

}  // namespace b200

// ---- C ABI --------------------------------------------------------------------------------------

using namespace b200;

extern "C" {

int b200fdtd_abi_version(void) { return B200FDTD_ABI_VERSION; }

const char* b200fdtd_last_error(void) { return g_last_error.c_str(); }

int b200fdtd_validate(const b200fdtd_desc* desc) { return validate(desc); }

int b200fdtd_num_outputs(const b200fdtd_desc* desc) {
  if (validate(desc)) return -1;
  return num_outputs(desc);
}

size_t b200fdtd_output_bytes(const b200fdtd_desc* desc) {
  if (validate(desc)) return 0;
  const size_t n = desc->proj_rows > 0 ? (size_t)desc->proj_rows : (size_t)num_outputs(desc);
  return n * 3 * desc->xx * desc->yy * desc->zz * sizeof(float);
}

size_t b200fdtd_workspace_bytes(const b200fdtd_desc* desc) {
  if (validate(desc)) return 0;
  const Geom g = make_geom(desc);
  Plan plan;
  if (make_plan(desc, g, &plan)) return 0;
  return carve(g, desc->use_reduced_precision != 0, is_systolic(plan.kernel), &plan.sys).total;
}

int b200fdtd_run(const b200fdtd_desc* desc, const void* const* inputs, void* const* outputs,
                 void* workspace, size_t workspace_bytes, void* stream) {
  return run_impl(desc, inputs, outputs, workspace, workspace_bytes,
                  static_cast<cudaStream_t>(stream));
}

int b200fdtd_run_host(const b200fdtd_desc* d, const void* const* hin, void* const* hout,
                      int device) {
  int rc = validate(d);
  if (rc) return rc;
  if (!hin || !hout) return fail(B200FDTD_EINVAL, "inputs/outputs is NULL");
  CUDA_TRY(cudaSetDevice(device));
  const size_t XY = (size_t)d->X * d->Y;
  size_t bytes[B200FDTD_MAX_INPUTS];
  const int nin = d->proj_rows > 0 ? B200FDTD_MAX_INPUTS : B200FDTD_NUM_INPUTS;
  bytes[B200FDTD_IN_PROJECTION] = (size_t)d->proj_rows * (num_outputs(d) > 0 ? num_outputs(d) : 0) * 4;
  bytes[B200FDTD_IN_EPSILON] = 3 * (size_t)d->xx * d->yy * d->zz * 4;
  bytes[B200FDTD_IN_SOURCE_FIELD] =
      (d->source_axis == 0 ? 2 * (size_t)d->Y * d->Z
                           : d->source_axis == 1 ? 2 * (size_t)d->X * d->Z : 4 * XY) * 4;
  bytes[B200FDTD_IN_SOURCE_WAVEFORM] = 2 * (size_t)(d->tt > 0 ? d->tt : 1) * 4;
  bytes[B200FDTD_IN_ABSORPTION_MASK] = 3 * XY * 4;
  bytes[B200FDTD_IN_PML_KAPPA] = bytes[B200FDTD_IN_PML_SIGMA] = bytes[B200FDTD_IN_PML_ALPHA] =
      2 * (size_t)d->Z * 4;
  const size_t out_bytes = b200fdtd_output_bytes(d);
  const size_t ws_bytes = b200fdtd_workspace_bytes(d);
  if (ws_bytes == 0) return B200FDTD_EINVAL;
  cudaStream_t st = nullptr;
  void* din[B200FDTD_MAX_INPUTS] = {nullptr};
  void* dout[1] = {nullptr};
  void* ws = nullptr;
  rc = B200FDTD_OK;
  auto cleanup = [&]() {
    // stream-ordered frees: the blocks go back to the library's private pool, which keeps them
    // cached so that the next call does not pay for fresh allocations
    for (auto p : din) if (p) cudaFreeAsync(p, st);
    if (dout[0]) cudaFreeAsync(dout[0], st);
    if (ws) cudaFreeAsync(ws, st);
    if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  };
#define TRY_CLEAN(expr)                                                                     \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      cleanup();                                                                            \
      return fail(B200FDTD_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));          \
    }                                                                                       \
  } while (0)
  TRY_CLEAN(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaMemPool_t pool = nullptr;
  rc = host_pool(device, &pool);
  if (rc) { cleanup(); return rc; }
  for (int i = 0; i < nin; ++i) {
    if (!hin[i]) { cleanup(); return fail(B200FDTD_EINVAL, "inputs[%d] is NULL", i); }
    TRY_CLEAN(cudaMallocFromPoolAsync(&din[i], bytes[i] ? bytes[i] : 4, pool, st));
    TRY_CLEAN(cudaMemcpyAsync(din[i], hin[i], d->tt > 0 || i != B200FDTD_IN_SOURCE_WAVEFORM
                                                  ? bytes[i] : 0,
                              cudaMemcpyHostToDevice, st));
  }
  TRY_CLEAN(cudaMallocFromPoolAsync(&dout[0], out_bytes ? out_bytes : 4, pool, st));
  TRY_CLEAN(cudaMallocFromPoolAsync(&ws, ws_bytes, pool, st));
  rc = run_impl(d, din, dout, ws, ws_bytes, st);
  if (rc == B200FDTD_OK && out_bytes) {
    if (!hout[0]) { cleanup(); return fail(B200FDTD_EINVAL, "outputs[0] is NULL"); }
    TRY_CLEAN(cudaMemcpyAsync(hout[0], dout[0], out_bytes, cudaMemcpyDeviceToHost, st));
  }
  TRY_CLEAN(cudaStreamSynchronize(st));
#undef TRY_CLEAN
  cleanup();
  return rc;
}

int b200fdtd_host_pool_trim(int device) {
  if (device < 0 || device >= 64) return fail(B200FDTD_EINVAL, "device %d out of range", device);
  std::lock_guard<std::mutex> lock(g_pool_mu);
  if (g_pools[device]) CUDA_TRY(cudaMemPoolTrimTo(g_pools[device], 0));
  return B200FDTD_OK;
}

// Body of both XLA custom-call entry points.  `opaque` is a b200fdtd_desc, optionally followed
// by a uint64 = the scratch bytes the wrapper declared at trace time (b200fdtd_workspace_bytes
// on the tracing device): the run is refused if this device's plan needs more.
static int custom_call_impl(void* stream, void** buffers, const char* opaque, size_t opaque_len) {
  const size_t with_ws = sizeof(b200fdtd_desc) + sizeof(uint64_t);
  if (!buffers || !opaque || (opaque_len != sizeof(b200fdtd_desc) && opaque_len != with_ws))
    return fail(B200FDTD_EINVAL,
                "custom call: opaque must be a b200fdtd_desc (%zu bytes) [+ uint64 scratch bytes], got %zu",
                sizeof(b200fdtd_desc), opaque_len);
  b200fdtd_desc d;
  memcpy(&d, opaque, sizeof d);
  int rc = validate(&d);
  if (rc) return rc;
  size_t ws_bytes = b200fdtd_workspace_bytes(&d);
  if (ws_bytes == 0) return B200FDTD_EUNSUPPORTED;       // last_error already says why
  if (opaque_len == with_ws) {
    uint64_t declared;
    memcpy(&declared, opaque + sizeof d, sizeof declared);
    if (declared < ws_bytes)
      return fail(B200FDTD_EWORKSPACE,
                  "custom call: scratch declared at trace time (%llu bytes) is smaller than this "
                  "device's plan needs (%zu bytes)", (unsigned long long)declared, ws_bytes);
    ws_bytes = (size_t)declared;
  }
  // operands: the 7 arrays (+ the projection matrix when proj_rows > 0), then result, scratch
  const int nin = d.proj_rows > 0 ? B200FDTD_MAX_INPUTS : B200FDTD_NUM_INPUTS;
  const void* ins[B200FDTD_MAX_INPUTS] = {nullptr};
  for (int i = 0; i < nin; ++i) ins[i] = buffers[i];
  void* outs[1] = {buffers[nin]};
  return run_impl(&d, ins, outs, buffers[nin + 1], ws_bytes, static_cast<cudaStream_t>(stream));
}

void b200fdtd_xla_custom_call(void* stream, void** buffers, const char* opaque,
                              size_t opaque_len) {
  custom_call_impl(stream, buffers, opaque, opaque_len);
}

void b200fdtd_xla_custom_call_status(void* stream, void** buffers, const char* opaque,
                                     size_t opaque_len, b200fdtd_xla_status* status) {
  const int rc = custom_call_impl(stream, buffers, opaque, opaque_len);
  if (rc != B200FDTD_OK && status) {
    // XlaCustomCallStatus is `struct { std::optional<std::string> message; }` on XLA's side and
    // must be set through XlaCustomCallStatusSetFailure, which lives in the host process
    // (jaxlib).  Resolve it at run time; without it the failure stays in last_error().
    typedef void (*set_failure_fn)(void*, const char*, size_t);
    static set_failure_fn set_failure =
        reinterpret_cast<set_failure_fn>(dlsym(RTLD_DEFAULT, "XlaCustomCallStatusSetFailure"));
    if (set_failure) set_failure(status, g_last_error.c_str(), g_last_error.size());
  }
}

int b200fdtd_plan_info(const b200fdtd_desc* desc, int64_t* info) {
  int rc = validate(desc);
  if (rc) return rc;
  if (!info) return fail(B200FDTD_EINVAL, "info is NULL");
  const Geom g = make_geom(desc);
  Plan plan;
  rc = make_plan(desc, g, &plan);
  if (rc) return rc;
  memset(info, 0, 8 * sizeof(int64_t));
  info[0] = plan.kernel;
  if (is_systolic(plan.kernel)) {
    info[1] = plan.sys.tile_y; info[2] = plan.sys.stages; info[3] = plan.sys.threads;
    info[4] = (int64_t)plan.sys.stages * plan.sys.ntiles; info[5] = plan.sys.smem_bytes;
    info[6] = plan.depth; info[7] = plan.sys.l2_window_bytes >> 20;
  } else {
    info[3] = kTwoPassThreads;
    info[4] = (int64_t)((g.Y * g.Zq + kTwoPassThreads - 1) / kTwoPassThreads) * g.X;
  }
  return B200FDTD_OK;
}


// ---- adjoint product-reduce -----------------------------------------------------

int b200fdtd_adjoint_reduce(int nports, int ww, size_t nvox, const void* const* fields,
                            const void* coef, void* out, void* stream) {
  if (nports < 1 || nports > kAdjMaxPorts)
    return fail(B200FDTD_EINVAL, "nports must be in [1, %d], got %d", kAdjMaxPorts, nports);
  if (ww < 1 || ww > 64) return fail(B200FDTD_EINVAL, "ww must be in [1, 64], got %d", ww);
  if (!fields || !coef || !out) return fail(B200FDTD_EINVAL, "fields/coef/out is NULL");
  AdjFields f;
  for (int i = 0; i < kAdjMaxPorts; ++i) f.f[i] = nullptr;
  for (int i = 0; i < nports; ++i) {
    if (!fields[i]) return fail(B200FDTD_EINVAL, "fields[%d] is NULL", i);
    f.f[i] = static_cast<const float2*>(fields[i]);
  }
  if (nvox == 0) return B200FDTD_OK;
  int sms = 0, l2 = 0;
  int rc = device_props(&sms, &l2);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float2* c = static_cast<const float2*>(coef);
  float* o = static_cast<float*>(out);
  cudaError_t e = cudaSuccess;
  switch (nports) {
#define ADJ_CASE(N) case N: e = adjoint_reduce_launch<N>(f, c, ww, nvox, o, sms, st); break;
    ADJ_CASE(1) ADJ_CASE(2) ADJ_CASE(3) ADJ_CASE(4) ADJ_CASE(5) ADJ_CASE(6) ADJ_CASE(7) ADJ_CASE(8)
    ADJ_CASE(9) ADJ_CASE(10) ADJ_CASE(11) ADJ_CASE(12) ADJ_CASE(13) ADJ_CASE(14) ADJ_CASE(15)
    ADJ_CASE(16)
#undef ADJ_CASE
  }
  if (e != cudaSuccess)
    return fail(B200FDTD_ECUDA, "adjoint_reduce launch failed: %s", cudaGetErrorString(e));
  return B200FDTD_OK;
}


// ---- snapshot projection + port overlaps -------------------------------------

int b200fdtd_project(int ww, int n_out, size_t nvox, const void* snapshots, const void* weights,
                     void* out, void* stream) {
  if (ww < 1 || n_out < 1) return fail(B200FDTD_EINVAL, "project: ww and n_out must be positive");
  if (!snapshots || !weights || !out) return fail(B200FDTD_EINVAL, "project: NULL argument");
  if (nvox == 0) return B200FDTD_OK;
  int sms = 0, l2 = 0;
  int rc = device_props(&sms, &l2);
  if (rc) return rc;
  const cudaError_t e = project_launch(static_cast<const float*>(snapshots),
                                       static_cast<const float*>(weights), ww, n_out, nvox,
                                       static_cast<float2*>(out), sms,
                                       static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess)
    return fail(B200FDTD_ECUDA, "project launch failed: %s", cudaGetErrorString(e));
  return B200FDTD_OK;
}

int b200fdtd_overlaps(int nfields, int nports, int ww, int xx, int yy, int zz,
                      const void* const* fields, const void* const* modes, const int* axes,
                      const int* planes, void* vals, void* stream) {
  if (nfields < 1 || nfields > kOvlMaxPorts || nports < 1 || nports > kOvlMaxPorts)
    return fail(B200FDTD_EINVAL, "overlaps: nfields and nports must be in [1, %d]", kOvlMaxPorts);
  if (ww < 1 || xx < 1 || yy < 1 || zz < 1) return fail(B200FDTD_EINVAL, "overlaps: bad extents");
  if (!fields || !modes || !axes || !planes || !vals)
    return fail(B200FDTD_EINVAL, "overlaps: NULL argument");
  OverlapFields f;
  OverlapJobs j;
  const int dims[3] = {xx, yy, zz};
  for (int i = 0; i < kOvlMaxPorts; ++i) { f.f[i] = nullptr; j.j[i].mode = nullptr; }
  for (int i = 0; i < nfields; ++i) {
    if (!fields[i]) return fail(B200FDTD_EINVAL, "overlaps: fields[%d] is NULL", i);
    f.f[i] = static_cast<const float2*>(fields[i]);
  }
  for (int m = 0; m < nports; ++m) {
    if (!modes[m]) return fail(B200FDTD_EINVAL, "overlaps: modes[%d] is NULL", m);
    if (axes[m] < 0 || axes[m] > 2) return fail(B200FDTD_EINVAL, "overlaps: axes[%d] must be 0..2", m);
    for (int k = 0; k < 2; ++k)
      if (planes[2 * m + k] < 0 || planes[2 * m + k] >= dims[axes[m]])
        return fail(B200FDTD_EINVAL, "overlaps: sample plane %d of port %d outside [0,%d)",
                    planes[2 * m + k], m, dims[axes[m]]);
    j.j[m].mode = static_cast<const float2*>(modes[m]);
    j.j[m].axis = axes[m];
    j.j[m].plane[0] = planes[2 * m]; j.j[m].plane[1] = planes[2 * m + 1];
  }
  overlap_kernel<<<(unsigned)(nfields * nports * 2 * ww), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      f, j, nfields, nports, ww, xx, yy, zz, static_cast<float2*>(vals));
  CUDA_TRY(cudaGetLastError());
  return B200FDTD_OK;
}


// ---- waveguide-mode operator ------------------------------------------------------

int b200fdtd_mode_operator(int ww, int uu, int vv, int mm, const void* eps, const void* omega,
                           const void* shift, const void* x, void* y, void* stream) {
  if (ww < 1 || uu < 1 || vv < 1 || mm < 1)
    return fail(B200FDTD_EINVAL, "mode_operator: extents must be positive");
  if (!eps || !omega || !shift || !x || !y) return fail(B200FDTD_EINVAL, "mode_operator: NULL argument");
  const size_t n = (size_t)ww * uu * vv * mm;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  mode_operator_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ww, uu, vv, mm, static_cast<const float*>(eps), static_cast<const float*>(omega),
      static_cast<const float*>(shift), static_cast<const float*>(x), static_cast<float*>(y));
  CUDA_TRY(cudaGetLastError());
  return B200FDTD_OK;
}


// ---- permittivity renderer --------------------------------------------------------

static size_t render_stats_bytes(int ll, int xx, int yy) {
  return align_up((size_t)3 * ll * xx * yy * sizeof(float4), 256);
}

size_t b200fdtd_render_workspace_bytes(int ll, int xx, int yy, int zz) {
  if (ll < 1 || xx < 1 || yy < 1 || zz < 1) return 0;
  return render_stats_bytes(ll, xx, yy) + ((size_t)4 * ll * zz + 2 * zz) * sizeof(double);
}

int b200fdtd_render(int ll, int xx, int yy, int zz, int m, const void* layers,
                    const void* layer_pos, const void* grid_start, const void* grid_end,
                    int use_simple_averaging, void* workspace, void* out, void* stream) {
  if (ll < 1 || xx < 1 || yy < 1 || zz < 1 || m < 1)
    return fail(B200FDTD_EINVAL, "render: ll, xx, yy, zz, m must be positive");
  if (!layers || !grid_start || !grid_end || !workspace || !out || (ll > 1 && !layer_pos))
    return fail(B200FDTD_EINVAL, "render: NULL argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto blocks = [](size_t n) { size_t b = (n + 255) / 256; return (unsigned)(b > 148 * 32 ? 148 * 32 : b); };
  double* tab = reinterpret_cast<double*>(static_cast<char*>(workspace) + render_stats_bytes(ll, xx, yy));
  tile_stats_kernel<<<blocks((size_t)3 * ll * xx * yy), 256, 0, st>>>(
      ll, xx, yy, m, static_cast<const float*>(layers), static_cast<float4*>(workspace));
  layer_overlap_kernel<<<blocks((size_t)2 * ll * zz), 256, 0, st>>>(
      ll, zz, static_cast<const float*>(layer_pos), static_cast<const float*>(grid_start),
      static_cast<const float*>(grid_end), tab);
  render_combine_kernel<<<blocks((size_t)3 * xx * yy * zz), 256, 0, st>>>(
      ll, xx, yy, zz, static_cast<const float4*>(workspace), tab, use_simple_averaging,
      static_cast<float*>(out));
  CUDA_TRY(cudaGetLastError());
  return B200FDTD_OK;
}

size_t b200fdtd_render_backward_workspace_bytes(int ll, int xx, int yy, int zz) {
  if (ll < 1 || xx < 1 || yy < 1 || zz < 1) return 0;
  return align_up(b200fdtd_render_workspace_bytes(ll, xx, yy, zz), 256) +
         (size_t)3 * ll * xx * yy * 4 * sizeof(double);
}

int b200fdtd_render_backward(int ll, int xx, int yy, int zz, int m, const void* layers,
                             const void* layer_pos, const void* grid_start, const void* grid_end,
                             int use_simple_averaging, const void* grad_out, void* workspace,
                             void* grad_layers, void* grad_overlap, void* stream) {
  if (ll < 1 || xx < 1 || yy < 1 || zz < 1 || m < 1)
    return fail(B200FDTD_EINVAL, "render_backward: ll, xx, yy, zz, m must be positive");
  if (!layers || !grid_start || !grid_end || !grad_out || !workspace || !grad_layers ||
      !grad_overlap || (ll > 1 && !layer_pos))
    return fail(B200FDTD_EINVAL, "render_backward: NULL argument");
  if ((size_t)2 * zz * sizeof(double) > 48 * 1024)
    return fail(B200FDTD_EUNSUPPORTED, "render_backward: zz too large for the block reduction");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto blocks = [](size_t n) { size_t b = (n + 255) / 256; return (unsigned)(b > 148 * 32 ? 148 * 32 : b); };
  char* ws = static_cast<char*>(workspace);
  float4* stats = reinterpret_cast<float4*>(ws);
  double* tab = reinterpret_cast<double*>(ws + render_stats_bytes(ll, xx, yy));
  double* dstats = reinterpret_cast<double*>(
      ws + align_up(b200fdtd_render_workspace_bytes(ll, xx, yy, zz), 256));
  // forward quantities again (cheap next to keeping them alive across the engine runs)
  tile_stats_kernel<<<blocks((size_t)3 * ll * xx * yy), 256, 0, st>>>(
      ll, xx, yy, m, static_cast<const float*>(layers), stats);
  layer_overlap_kernel<<<blocks((size_t)2 * ll * zz), 256, 0, st>>>(
      ll, zz, static_cast<const float*>(layer_pos), static_cast<const float*>(grid_start),
      static_cast<const float*>(grid_end), tab);
  CUDA_TRY(cudaMemsetAsync(grad_overlap, 0, (size_t)4 * ll * zz * sizeof(double), st));
  unsigned bx = (unsigned)(((size_t)xx * yy + 255) / 256);
  if (bx > 148 * 4) bx = 148 * 4;
  render_combine_bwd_kernel<<<dim3(bx, 3 * ll), 256, 2 * zz * sizeof(double), st>>>(
      ll, xx, yy, zz, stats, tab, use_simple_averaging, static_cast<const float*>(grad_out), dstats,
      static_cast<double*>(grad_overlap));
  tile_stats_bwd_kernel<<<blocks((size_t)ll * 4 * m * m * xx * yy), 256, 0, st>>>(
      ll, xx, yy, m, static_cast<const float*>(layers), dstats, static_cast<float*>(grad_layers));
  CUDA_TRY(cudaGetLastError());
  return B200FDTD_OK;
}


// ---- stepping sessions (host-driven time loops: domain decomposition with halo exchange) ------

struct b200fdtd_session {
  b200fdtd_desc d;
  Geom g;
  Workspace w;
  Plan plan;
  bool systolic;        // persistent-launch kernel (lean): ping-pong state, advance() only
  char* ws;
  bool reduced;
  Ptrs<float> pf;
  Ptrs<__half> ph;
  dim3 grid;
  bool slab;            // y-slab session: tiles cover [g.ylo, g.yhi), in-kernel halo exchange
  bool peers_set;
  bool counters_clear;  // slab: b200fdtd_session_slab_reset ran since the last launch
  SlabPeers peers;
};

// Sessions use the lean systolic kernels when the plan selects them (fp32 with 32 z-vectors, or
// the sub-warp variant for short columns in either storage type) and the per-step kernels
// otherwise.
static int session_plan(const b200fdtd_desc* desc, const Geom& g, Plan* plan, bool* systolic) {
  *systolic = false;
  plan->kernel = B200FDTD_KERNEL_TWOPASS;
  plan->depth = 0;
  if (desc->kernel == B200FDTD_KERNEL_TWOPASS) return B200FDTD_OK;
  Plan p;
  int rc = make_plan(desc, g, &p);
  if (rc) {
    if (desc->kernel == B200FDTD_KERNEL_AUTO) return B200FDTD_OK;
    return rc;
  }
  if (p.kernel == B200FDTD_KERNEL_SYSTOLIC_LEAN) {
    *plan = p;
    *systolic = true;
  } else if (desc->kernel != B200FDTD_KERNEL_AUTO) {
    return fail(B200FDTD_EUNSUPPORTED, "sessions support the twopass and systolic_lean kernels");
  }
  return B200FDTD_OK;
}

static int slab_geom(const b200fdtd_desc* desc, int ylo, int yhi, Geom* g) {
  *g = make_geom(desc);
  if (ylo == 0 && yhi == desc->Y) return B200FDTD_OK;     // the whole domain: an ordinary session
  if (ylo < 1 || yhi <= ylo || yhi > desc->Y - 1)
    return fail(B200FDTD_EINVAL, "slab columns [%d,%d) need one ghost column on each side of Y=%d",
                ylo, yhi, desc->Y);
  g->ylo = ylo; g->yhi = yhi;
  return B200FDTD_OK;
}

size_t b200fdtd_session_workspace_bytes_slab(const b200fdtd_desc* desc, int ylo, int yhi) {
  if (validate(desc)) return 0;
  Geom g;
  if (slab_geom(desc, ylo, yhi, &g)) return 0;
  Plan plan;
  bool systolic;
  if (session_plan(desc, g, &plan, &systolic)) return 0;
  if (g.yhi - g.ylo != g.Y && !systolic) return 0;
  return carve(g, desc->use_reduced_precision != 0, systolic, &plan.sys).total;
}

size_t b200fdtd_session_workspace_bytes(const b200fdtd_desc* desc) {
  if (validate(desc)) return 0;
  const Geom g = make_geom(desc);
  Plan plan;
  bool systolic;
  if (session_plan(desc, g, &plan, &systolic)) return 0;
  return carve(g, desc->use_reduced_precision != 0, systolic, &plan.sys).total;
}

int b200fdtd_session_create(const b200fdtd_desc* desc, const void* const* inputs,
                            void* const* outputs, void* workspace, size_t workspace_bytes,
                            void* stream, b200fdtd_session** session) {
  return b200fdtd_session_create_slab(desc, inputs, outputs, workspace, workspace_bytes, stream,
                                      0, desc ? desc->Y : 0, session);
}

int b200fdtd_session_create_slab(const b200fdtd_desc* desc, const void* const* inputs,
                                 void* const* outputs, void* workspace, size_t workspace_bytes,
                                 void* stream, int ylo, int yhi, b200fdtd_session** session) {
  int rc = validate(desc);
  if (rc) return rc;
  if (!session) return fail(B200FDTD_EINVAL, "session is NULL");
  if (!inputs || !outputs || !workspace) return fail(B200FDTD_EINVAL, "inputs/outputs/workspace is NULL");
  for (int i = 0; i < B200FDTD_NUM_INPUTS; ++i)
    if (!inputs[i]) return fail(B200FDTD_EINVAL, "inputs[%d] is NULL", i);
  if (num_outputs(desc) > 0 && !outputs[0]) return fail(B200FDTD_EINVAL, "outputs[0] is NULL");
  Geom g;
  rc = slab_geom(desc, ylo, yhi, &g);
  if (rc) return rc;
  const bool slab = g.yhi - g.ylo != g.Y;
  Plan plan;
  bool systolic;
  rc = session_plan(desc, g, &plan, &systolic);
  if (rc) return rc;
  if (slab && !(systolic && plan.sys.cols == 2))
    return fail(B200FDTD_EUNSUPPORTED,
                "y-slab sessions need the warp-per-column-pair kernel (fp32, 125 <= Z <= 128)");
  if (!systolic && g.X > 65535) return fail(B200FDTD_EUNSUPPORTED, "per-step sessions support X <= 65535");
  const Workspace w = carve(g, desc->use_reduced_precision != 0, systolic, &plan.sys);
  if (workspace_bytes < w.total)
    return fail(B200FDTD_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total,
                workspace_bytes);
  if ((uintptr_t)workspace % 256 != 0) return fail(B200FDTD_EINVAL, "workspace must be 256-byte aligned");
  b200fdtd_session* s = new (std::nothrow) b200fdtd_session();
  if (!s) return fail(B200FDTD_EINVAL, "out of host memory");
  s->d = *desc; s->g = g; s->w = w; s->ws = static_cast<char*>(workspace);
  s->plan = plan; s->systolic = systolic;
  s->slab = slab; s->peers_set = false; s->counters_clear = false;
  s->peers.delta_lo = 0; s->peers.delta_hi = 0; s->peers.enabled = slab ? 1 : 0;
  s->reduced = desc->use_reduced_precision != 0;
  s->grid = dim3((g.Y * g.Zq + kTwoPassThreads - 1) / kTwoPassThreads, g.X);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rc = s->reduced ? prepare_typed<__half>(desc, g, w, inputs, outputs, s->ws, st, &s->ph)
                  : prepare_typed<float>(desc, g, w, inputs, outputs, s->ws, st, &s->pf);
  if (rc) { delete s; return rc; }
  *session = s;
  return B200FDTD_OK;
}

int b200fdtd_session_step_h(b200fdtd_session* s, void* stream) {
  if (!s) return fail(B200FDTD_EINVAL, "session is NULL");
  if (s->systolic) return fail(B200FDTD_EUNSUPPORTED, "systolic sessions advance whole steps");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (s->reduced) twopass_h_kernel<__half><<<s->grid, kTwoPassThreads, 0, st>>>(s->g, s->ph);
  else twopass_h_kernel<float><<<s->grid, kTwoPassThreads, 0, st>>>(s->g, s->pf);
  CUDA_TRY(cudaGetLastError());
  return B200FDTD_OK;
}

int b200fdtd_session_step_e(b200fdtd_session* s, int n, void* stream) {
  if (!s) return fail(B200FDTD_EINVAL, "session is NULL");
  if (s->systolic) return fail(B200FDTD_EUNSUPPORTED, "systolic sessions advance whole steps");
  if (n < 0 || n >= s->g.tt) return fail(B200FDTD_EINVAL, "step %d outside [0,%d)", n, s->g.tt);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (s->reduced) twopass_e_kernel<__half><<<s->grid, kTwoPassThreads, 0, st>>>(s->g, s->ph, n);
  else twopass_e_kernel<float><<<s->grid, kTwoPassThreads, 0, st>>>(s->g, s->pf, n);
  CUDA_TRY(cudaGetLastError());
  return B200FDTD_OK;
}

int b200fdtd_session_advance(b200fdtd_session* s, int n0, int nsteps, void* stream) {
  if (!s) return fail(B200FDTD_EINVAL, "session is NULL");
  if (n0 < 0 || nsteps < 0 || n0 + nsteps > s->d.tt)
    return fail(B200FDTD_EINVAL, "steps [%d,%d) outside [0,%d)", n0, n0 + nsteps, s->d.tt);
  if (nsteps == 0) return B200FDTD_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!s->systolic) {
    for (int n = n0; n < n0 + nsteps; ++n) {
      int rc = b200fdtd_session_step_h(s, stream);
      if (rc) return rc;
      rc = b200fdtd_session_step_e(s, n, stream);
      if (rc) return rc;
    }
    return B200FDTD_OK;
  }
  // One persistent launch over steps [n0, n0+nsteps): the state of step n lives in buffer set
  // n & 1, so consecutive calls chain without copies.  Progress counters restart at zero.
  Geom g = s->g;
  g.n0 = n0;
  g.tt = n0 + nsteps;
  if (s->slab && !s->peers_set)
    return fail(B200FDTD_EINVAL, "slab session: call b200fdtd_session_set_peers first");
  // A slab's counters (and the mirror slots its neighbours write) are cleared by
  // b200fdtd_session_slab_reset, between two barriers of the caller -- not here.  A launch over
  // counters the previous launch left behind would see every dependency as already met.
  if (s->slab && !s->counters_clear)
    return fail(B200FDTD_EINVAL, "slab session: call b200fdtd_session_slab_reset (between two "
                                 "barriers of all ranks) before every advance");
  if (!s->slab)
    CUDA_TRY(cudaMemsetAsync(s->ws + s->w.sync, 0, systolic_sync_bytes(s->plan.sys), st));
  s->counters_clear = false;
  unsigned* const sync = reinterpret_cast<unsigned*>(s->ws + s->w.sync);
  int rc;
  if (s->plan.sys.cols == 16)                      // sub-warp variant: fp16 or fp32 storage
    rc = s->reduced ? lean16_launch<__half>(g, s->ph, s->plan.sys, sync, st)
                    : lean16_launch<float>(g, s->pf, s->plan.sys, sync, st);
  else
    rc = lean_launch(g, s->pf, s->plan.sys, sync, st, s->slab ? &s->peers : nullptr);
  if (rc != 0)
    return fail(B200FDTD_ECUDA, "systolic launch failed: %s", cudaGetErrorString((cudaError_t)rc));
  return B200FDTD_OK;
}

int b200fdtd_session_set_peers(b200fdtd_session* s, void* workspace_lo, void* workspace_hi) {
  if (!s) return fail(B200FDTD_EINVAL, "session is NULL");
  if (!s->slab) return fail(B200FDTD_EINVAL, "not a y-slab session");
  if (!workspace_lo || !workspace_hi) return fail(B200FDTD_EINVAL, "peer workspace is NULL");
  s->peers.delta_lo = static_cast<char*>(workspace_lo) - s->ws;
  s->peers.delta_hi = static_cast<char*>(workspace_hi) - s->ws;
  s->peers.enabled = 1;
  s->peers_set = true;
  return B200FDTD_OK;
}

int b200fdtd_session_slab_reset(b200fdtd_session* s, void* stream) {
  if (!s) return fail(B200FDTD_EINVAL, "session is NULL");
  if (!s->slab) return fail(B200FDTD_EINVAL, "not a y-slab session");
  CUDA_TRY(cudaMemsetAsync(s->ws + s->w.sync, 0, systolic_sync_bytes(s->plan.sys),
                           static_cast<cudaStream_t>(stream)));
  s->counters_clear = true;
  return B200FDTD_OK;
}

int b200fdtd_session_layout(const b200fdtd_session* s, int64_t* info) {
  if (!s || !info) return fail(B200FDTD_EINVAL, "session/info is NULL");
  const int64_t el = s->reduced ? 2 : 4;
  info[0] = (int64_t)s->w.fields;                 // byte offset of Ex in the workspace
  info[1] = (int64_t)s->w.fields + 3 * s->g.N * el;   // byte offset of Hx
  info[2] = s->g.N * el;                          // bytes between components
  info[3] = s->g.P * el;                          // bytes per x-plane
  info[4] = (int64_t)s->g.Zp;                     // padded z extent (elements per column)
  info[5] = el;                                   // bytes per element (4 fp32, 2 fp16 storage)
  info[6] = s->g.X;
  info[7] = s->g.Y;
  return B200FDTD_OK;
}

int b200fdtd_session_layout2(const b200fdtd_session* s, int64_t* info) {
  int rc = b200fdtd_session_layout(s, info);
  if (rc) return rc;
  const int64_t VW = s->reduced ? 8 : 4;
  const int64_t psi_bytes = (int64_t)s->g.X * s->g.Y * s->g.npg * VW * 4;
  info[8] = s->systolic ? (int64_t)s->w.fields2 : -1;   // Ex of buffer set 1 (Hx follows as in set 0)
  info[9] = (int64_t)s->w.psi;                    // psiH[0] of set 0; then psiH[1], psiE[0], psiE[1]
  info[10] = psi_bytes;                           // bytes per psi array
  info[11] = s->systolic ? (int64_t)s->w.psi + 4 * psi_bytes : -1;   // psiH[0] of set 1, then psiH[1]
  info[12] = (int64_t)s->g.npg * VW;              // psi floats per (x, y) column
  info[13] = s->systolic ? 1 : 0;                 // 1: state of step n lives in set n & 1
  info[14] = s->plan.kernel;
  info[15] = s->systolic ? s->plan.sys.stages : 0;
  return B200FDTD_OK;
}

int b200fdtd_peer_alloc(size_t bytes, void** ptr) {
  if (!ptr || bytes == 0) return fail(B200FDTD_EINVAL, "peer_alloc: bad argument");
  CUDA_TRY(cudaMalloc(ptr, bytes));
  return B200FDTD_OK;
}

int b200fdtd_peer_free(void* ptr) {
  CUDA_TRY(cudaFree(ptr));
  return B200FDTD_OK;
}

int b200fdtd_peer_export(void* ptr, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!ptr || !handle) return fail(B200FDTD_EINVAL, "peer_export: NULL argument");
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle, &h, sizeof h);
  return B200FDTD_OK;
}

int b200fdtd_peer_open(const unsigned char handle[64], void** ptr) {
  if (!ptr || !handle) return fail(B200FDTD_EINVAL, "peer_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return B200FDTD_OK;
}

int b200fdtd_peer_close(void* ptr) {
  CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return B200FDTD_OK;
}

void b200fdtd_session_destroy(b200fdtd_session* s) { delete s; }

}  // extern "C"
