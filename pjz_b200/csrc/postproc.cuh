// Adjoint product-reduce: the gradient of pjz.scatter's custom_vjp,
//     dL/d eps[v] = sum_{i,j} sum_w Re( c_ij[w] * F_i[w][v] * F_j[w][v] ),   c_ij = conj(g_ij) / a_i
// (/root/reference/src/pjz/_field.py:380-382 forms grads[i][j] = F_i F_j / a_i as N^2 full-volume
// complex temporaries and :393-398 reduces them against the cotangents), as ONE pass over the N
// phasor fields: every voxel-component is read once per port and frequency and written once.
// HBM-bound: N*ww*8 bytes in, 4 bytes out per voxel-component.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr int kAdjMaxPorts = 16;

struct AdjFields {
  const float2* f[kAdjMaxPorts];   // port i: (ww, nvox) complex64
};

// coef: (nports, nports, ww) complex64, staged in shared memory; symmetrised on the fly
// (F_i F_j = F_j F_i), so only pairs i <= j are multiplied.
template <int NP>
__global__ void __launch_bounds__(256)
adjoint_reduce_kernel(const AdjFields fields, const float2* __restrict__ coef, int ww,
                      size_t nvox, float* __restrict__ out) {
  extern __shared__ float2 sc[];     // [ww][NP][NP], upper triangle holds c_ij + c_ji (i < j)
  for (int k = threadIdx.x; k < ww * NP * NP; k += blockDim.x) {
    const int w = k / (NP * NP), i = (k / NP) % NP, j = k % NP;
    float2 c = make_float2(0.f, 0.f);
    if (i == j) c = coef[((size_t)i * NP + j) * ww + w];
    else if (i < j) {
      const float2 a = coef[((size_t)i * NP + j) * ww + w], b = coef[((size_t)j * NP + i) * ww + w];
      c = make_float2(a.x + b.x, a.y + b.y);
    }
    sc[k] = c;
  }
  __syncthreads();
  for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvox;
       v += (size_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int w = 0; w < ww; ++w) {
      float2 F[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) F[i] = __ldcs(fields.f[i] + (size_t)w * nvox + v);
      const float2* c = sc + (size_t)w * NP * NP;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        // t = sum_{j >= i} c_ij F_j ; acc += Re(F_i * t)
        float tx = 0.f, ty = 0.f;
#pragma unroll
        for (int j = i; j < NP; ++j) {
          const float2 cc = c[i * NP + j];
          tx += cc.x * F[j].x - cc.y * F[j].y;
          ty += cc.x * F[j].y + cc.y * F[j].x;
        }
        acc += F[i].x * tx - F[i].y * ty;
      }
    }
    out[v] = acc;
  }
}

template <int NP>
inline cudaError_t adjoint_reduce_launch(const AdjFields& f, const float2* coef, int ww,
                                         size_t nvox, float* out, int sms, cudaStream_t st) {
  const size_t smem = sizeof(float2) * (size_t)ww * NP * NP;
  size_t blocks = (nvox + 255) / 256;
  const size_t cap = (size_t)sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (smem > 48 * 1024) {                          // up to 16 ports x 64 frequencies = 128 KB
    const cudaError_t e = cudaFuncSetAttribute(adjoint_reduce_kernel<NP>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  adjoint_reduce_kernel<NP><<<(unsigned)blocks, 256, smem, st>>>(f, coef, ww, nvox, out);
  return cudaGetLastError();
}

}  // namespace b200

// ---- waveguide-mode operator ---------------------------------------------------
// y = op(x) of the shifted subspace iteration of /root/reference/src/pjz/_mode.py:22-51, batched
// over the ww frequencies and the mm trial vectors in ONE launch:
//   a = (omega^2 eps_yx - shift) x
//   b = (-d-_v x0 + d-_u x1) / eps_z ;  b = (-d+_v b, d+_u b) * eps_yx
//   c = d+_u x0 + d+_v x1            ;  c = (d-_u c, d-_v c)
//   y = a + b + c                      (d+/d- = periodic forward/backward differences)
// x, y: (ww, 2, uu, vv, mm) float32; eps: (3, uu, vv) in "propagate-along-z" form.
// One thread per (w, u, v, m); the 13-point neighbourhood is re-read through L1/L2 (a port
// cross-section is a few hundred KB at most).
namespace b200 {

__global__ void __launch_bounds__(256)
mode_operator_kernel(int ww, int uu, int vv, int mm, const float* __restrict__ eps,
                     const float* __restrict__ omega, const float* __restrict__ shift,
                     const float* __restrict__ x, float* __restrict__ y) {
  const size_t n = (size_t)ww * uu * vv * mm;
  const size_t plane = (size_t)uu * vv;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i % mm);
    size_t r = i / mm;
    const int v = (int)(r % vv); r /= vv;
    const int u = (int)(r % uu);
    const int w = (int)(r / uu);
    const float* x0 = x + ((size_t)w * 2 + 0) * plane * mm + m;
    const float* x1 = x + ((size_t)w * 2 + 1) * plane * mm + m;
    auto X0 = [&](int a, int b) { return x0[((size_t)a * vv + b) * mm]; };
    auto X1 = [&](int a, int b) { return x1[((size_t)a * vv + b) * mm]; };
    auto E = [&](int c, int a, int b) { return eps[(size_t)c * plane + (size_t)a * vv + b]; };
    const int up = u + 1 == uu ? 0 : u + 1, um = u == 0 ? uu - 1 : u - 1;
    const int vp = v + 1 == vv ? 0 : v + 1, vm = v == 0 ? vv - 1 : v - 1;
    // b(a, b) = (-(x0[a][b] - x0[a][b-1]) + (x1[a][b] - x1[a-1][b])) / eps_z[a][b]
    auto Bq = [&](int a, int b) {
      const int am = a == 0 ? uu - 1 : a - 1, bm = b == 0 ? vv - 1 : b - 1;
      return (-(X0(a, b) - X0(a, bm)) + (X1(a, b) - X1(am, b))) / E(2, a, b);
    };
    // c(a, b) = (x0[a+1][b] - x0[a][b]) + (x1[a][b+1] - x1[a][b])
    auto Cq = [&](int a, int b) {
      const int ap = a + 1 == uu ? 0 : a + 1, bp = b + 1 == vv ? 0 : b + 1;
      return (X0(ap, b) - X0(a, b)) + (X1(a, bp) - X1(a, b));
    };
    const float om2 = omega[w] * omega[w], sh = shift[w];
    const float e_y = E(1, u, v), e_x = E(0, u, v);
    const float b00 = Bq(u, v), c00 = Cq(u, v);
    const float y0 = (om2 * e_y - sh) * X0(u, v) + (-(Bq(u, vp) - b00)) * e_y + (c00 - Cq(um, v));
    const float y1 = (om2 * e_x - sh) * X1(u, v) + (Bq(up, v) - b00) * e_x + (c00 - Cq(u, vm));
    const size_t o = ((size_t)u * vv + v) * mm + m;
    y[((size_t)w * 2 + 0) * plane * mm + o] = y0;
    y[((size_t)w * 2 + 1) * plane * mm + o] = y1;
  }
}


// ---- snapshot -> phasor projection -------------------------------------------
// field() dumps n_out = 2ww+1 real snapshots and projects them onto the ww complex phasors with
// the pseudo-inverse of the sampled phases (/root/reference/src/pjz/_field.py:272-279:
// `jnp.einsum("ij,j...->i...", pinv, fields)` followed by `outputs[:ww] + 1j * outputs[ww:]`).  One
// pass: every snapshot value is read once (n_out * 4 bytes per voxel-component) and the ww complex
// phasors are written once, interleaved (ww * 8 bytes) -- the GEMM + complex() pair of the eager
// path moves 2 * ww * 8 bytes more.  HBM-bound streaming kernel, 16-byte accesses.
//   snaps (n_out, nvox) float32; W (2ww, n_out) float32 row-major; out (ww, nvox) complex64.
constexpr int kProjMaxWW = 8;        // register-resident accumulators: 2 * WW float4 per thread
constexpr int kProjMaxOut = 64;

template <int WW>
__global__ void __launch_bounds__(256)
project_kernel(const float* __restrict__ snaps, const float* __restrict__ W, int n_out, size_t nvox,
               float2* __restrict__ out) {
  __shared__ float sw[2 * WW * kProjMaxOut];
  for (int k = threadIdx.x; k < 2 * WW * n_out; k += blockDim.x) sw[k] = W[k];
  __syncthreads();
  const size_t nvec = nvox / 4;      // whole float4 groups (nvox % 4 == 0 on this path)
  for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvec;
       v += (size_t)gridDim.x * blockDim.x) {
    float4 acc[2 * WW];
#pragma unroll
    for (int r = 0; r < 2 * WW; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* src = reinterpret_cast<const float4*>(snaps) + v;
#pragma unroll 3
    for (int s = 0; s < n_out; ++s) {
      const float4 x = __ldcs(src + (size_t)s * nvec);
#pragma unroll
      for (int r = 0; r < 2 * WW; ++r) {
        const float w = sw[r * n_out + s];
        acc[r].x = fmaf(w, x.x, acc[r].x); acc[r].y = fmaf(w, x.y, acc[r].y);
        acc[r].z = fmaf(w, x.z, acc[r].z); acc[r].w = fmaf(w, x.w, acc[r].w);
      }
    }
#pragma unroll
    for (int w = 0; w < WW; ++w) {
      float4* o = reinterpret_cast<float4*>(out + (size_t)w * nvox) + 2 * v;
      __stcs(o, make_float4(acc[w].x, acc[WW + w].x, acc[w].y, acc[WW + w].y));
      __stcs(o + 1, make_float4(acc[w].z, acc[WW + w].z, acc[w].w, acc[WW + w].w));
    }
  }
}

// Any ww / unaligned volumes: one voxel-component per thread, frequencies looped (the snapshots
// of a voxel are re-read from L1/L2 for every frequency).
__global__ void __launch_bounds__(256)
project_generic_kernel(const float* __restrict__ snaps, const float* __restrict__ W, int ww, int n_out,
                       size_t nvox, float2* __restrict__ out) {
  for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvox;
       v += (size_t)gridDim.x * blockDim.x) {
    for (int w = 0; w < ww; ++w) {
      float re = 0.f, im = 0.f;
      for (int s = 0; s < n_out; ++s) {
        const float x = snaps[(size_t)s * nvox + v];
        re = fmaf(__ldg(W + (size_t)w * n_out + s), x, re);
        im = fmaf(__ldg(W + (size_t)(ww + w) * n_out + s), x, im);
      }
      out[(size_t)w * nvox + v] = make_float2(re, im);
    }
  }
}

inline cudaError_t project_launch(const float* snaps, const float* W, int ww, int n_out, size_t nvox,
                                  float2* out, int sms, cudaStream_t st) {
  const bool vec = nvox % 4 == 0 && ww <= kProjMaxWW && n_out <= kProjMaxOut &&
                   (reinterpret_cast<uintptr_t>(snaps) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  const size_t work = vec ? nvox / 4 : nvox;
  size_t blocks = (work + 255) / 256;
  const size_t cap = (size_t)sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (!vec) {
    project_generic_kernel<<<(unsigned)blocks, 256, 0, st>>>(snaps, W, ww, n_out, nvox, out);
    return cudaGetLastError();
  }
  switch (ww) {
#define PROJ_CASE(N) case N: project_kernel<N><<<(unsigned)blocks, 256, 0, st>>>(snaps, W, n_out, nvox, out); break;
    PROJ_CASE(1) PROJ_CASE(2) PROJ_CASE(3) PROJ_CASE(4) PROJ_CASE(5) PROJ_CASE(6) PROJ_CASE(7) PROJ_CASE(8)
#undef PROJ_CASE
  }
  return cudaGetLastError();
}

// ---- port overlaps (/root/reference/src/pjz/_field.py:305-338) -------------------------------------
// vals[f][m][k][w] = sum over the two transverse components c and the port plane of
//     mode_m[w][c][u][v] * field_f[w][comp_c][plane k of port m]
// for every (phasor field f, port m, sample plane k in {0,1}, frequency w): one block per value,
// fixed-order tree reduction (deterministic).  Replaces nports^2 x 2 eager slice-multiply-sum chains.
struct OverlapJob {
  const float2* mode;     // (ww, 2, U, V) complex64: transverse mode profile of this port
  int axis;               // propagation axis 0 | 1 | 2
  int plane[2];           // the two sample planes along `axis`
};
constexpr int kOvlMaxPorts = 16;
struct OverlapJobs { OverlapJob j[kOvlMaxPorts]; };
struct OverlapFields { const float2* f[kOvlMaxPorts]; };   // (ww, 3, xx, yy, zz) complex64 each

__global__ void __launch_bounds__(256)
overlap_kernel(const OverlapFields fields, const OverlapJobs jobs, int nfields, int nports, int ww,
               int xx, int yy, int zz, float2* __restrict__ vals) {
  int b = blockIdx.x;
  const int w = b % ww; b /= ww;
  const int k = b % 2; b /= 2;
  const int m = b % nports; const int f = b / nports;
  const OverlapJob job = jobs.j[m];
  const int dims[3] = {xx, yy, zz};
  const int a = job.axis, ua = a == 0 ? 1 : 0, va = a == 2 ? 1 : 2;   // transverse axes (ascending)
  const int U = dims[ua], V = dims[va];
  const size_t comp = (size_t)xx * yy * zz;
  const float2* F = fields.f[f] + (size_t)w * 3 * comp;
  const float2* M = job.mode + (size_t)w * 2 * U * V;
  const int p = job.plane[k];
  double re = 0.0, im = 0.0;
  for (int idx = threadIdx.x; idx < 2 * U * V; idx += blockDim.x) {
    const int c = idx / (U * V), r = idx % (U * V), u = r / V, v = r % V;
    const int fc = c == 0 ? ua : va;                 // transverse components in ascending order
    int pos[3];
    pos[a] = p; pos[ua] = u; pos[va] = v;
    const float2 x = F[(size_t)fc * comp + ((size_t)pos[0] * yy + pos[1]) * zz + pos[2]];
    const float2 mm = M[idx];
    re += (double)mm.x * x.x - (double)mm.y * x.y;
    im += (double)mm.x * x.y + (double)mm.y * x.x;
  }
  __shared__ double sre[256], sim[256];
  sre[threadIdx.x] = re; sim[threadIdx.x] = im;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { sre[threadIdx.x] += sre[threadIdx.x + o]; sim[threadIdx.x] += sim[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) vals[blockIdx.x] = make_float2((float)sre[0], (float)sim[0]);
}

}  // namespace b200
