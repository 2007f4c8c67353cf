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
  adjoint_reduce_kernel<NP><<<(unsigned)blocks, 256, smem, st>>>(f, coef, ww, nvox, out);
  return cudaGetLastError();
}

}  // namespace b200
