// Permittivity renderer: restates /root/reference/src/pjz/_epsilon.py:10-101
// (`_render_single` for the three Yee components) as two CUDA kernels that write epsilon
// directly in the engine's input layout (3, xx, yy, zz) float32, z fastest.
//
//   tile_stats_kernel : per (component, layer, X, Y) reduce the 2m x 2m tile of the (half-cell
//                       shifted, edge-replicated) layer image to {avg, avg of inverse, d/dx, d/dy}
//                       (:13-45: offsets, "layer-chunked" form, gradient weights);
//   combine_kernel    : per (component, X, Y, z) blend the layers by their overlap with the cell
//                       (:47-66), form the z gradient (:80-85) and the anisotropic average
//                       1 / (p aoi + (1 - p) ioa) with p the on-axis projection weight (:87-92).
// Accumulation is in double so the reference's golden values (tests/test_layers.py) are met to
// float32 rounding.  HBM traffic: the layer image once per component plus the output volume.
// The backward pass (below) gives d/d(layers) and the gradient with respect to the layer-overlap
// table, from which the python wrapper obtains d/d(layer_pos).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace b200 {

__global__ void __launch_bounds__(256)
tile_stats_kernel(int ll, int xx, int yy, int m, const float* __restrict__ layers,
                  float4* __restrict__ stats) {
  const int W = 2 * m, LX = W * xx, LY = W * yy;
  const size_t n = (size_t)3 * ll * xx * yy;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const int Y = (int)(i % yy);
    size_t r = i / yy;
    const int X = (int)(r % xx); r /= xx;
    const int l = (int)(r % ll);
    const int a = (int)(r / ll);                   // 0: Ex, 1: Ey, 2: Ez
    const float* img = layers + (size_t)l * LX * LY;
    double s = 0, si = 0, gx = 0, gy = 0;
    for (int di = 0; di < W; ++di) {
      int pi = X * W + di;
      if (a != 0) pi = pi - m < 0 ? 0 : pi - m;    // in-plane half-cell offset, edge-replicated
      const double wi = (di - (m - 0.5)) / ((double)W * W);
      for (int dj = 0; dj < W; ++dj) {
        int pj = Y * W + dj;
        if (a != 1) pj = pj - m < 0 ? 0 : pj - m;
        const double wj = (dj - (m - 0.5)) / ((double)W * W);
        const double v = img[(size_t)pi * LY + pj];
        s += v; si += 1.0 / v; gx += v * wi; gy += v * wj;
      }
    }
    const double inv = 1.0 / ((double)W * W);
    stats[i] = make_float4((float)(s * inv), (float)(si * inv), (float)(12.0 * W * gx * inv),
                           (float)(12.0 * W * gy * inv));
  }
}

// Overlap of every layer with every cell along z: u = covered fraction, uz = u * (centroid offset
// from the cell centre), and the cell's 12 / dz^2 -- a (2, ll, zz) table for the two z staggerings
// (column 0: Ex/Ey cells, column 1: Ez cells; _epsilon.py:47-66, 80-85).
__global__ void layer_overlap_kernel(int ll, int zz, const float* __restrict__ layer_pos,
                                     const float* __restrict__ grid_start,
                                     const float* __restrict__ grid_end, double* __restrict__ tab) {
  const int n = 2 * ll * zz;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int z = i % zz, l = (i / zz) % ll, col = i / (zz * ll);
    const double gs = grid_start[2 * z + col], ge = grid_end[2 * z + col];
    const double lo = l == 0 ? -INFINITY : (double)layer_pos[l - 1];
    const double hi = l == ll - 1 ? INFINITY : (double)layer_pos[l];
    const double p0 = fmin(fmax(lo, gs), ge), p1 = fmin(fmax(hi, gs), ge);
    const double u = (p1 - p0) / (ge - gs);
    tab[i] = u;
    tab[n + i] = u * (0.5 * (p0 + p1) - 0.5 * (gs + ge));
    if (l == 0) tab[2 * n + col * zz + z] = 12.0 / ((ge - gs) * (ge - gs));
  }
}

__global__ void __launch_bounds__(256)
render_combine_kernel(int ll, int xx, int yy, int zz, const float4* __restrict__ stats,
                      const double* __restrict__ tab, int simple, float* __restrict__ out) {
  const size_t n = (size_t)3 * xx * yy * zz;
  const int nt = 2 * ll * zz;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const int z = (int)(i % zz);
    size_t r = i / zz;
    const int Y = (int)(r % yy); r /= yy;
    const int X = (int)(r % xx);
    const int a = (int)(r / xx);
    const int col = a == 2 ? 1 : 0;                // Ez sits half a cell up in z (:22-27)
    double avg = 0, aoi = 0, gx = 0, gy = 0, gz = 0;
    for (int l = 0; l < ll; ++l) {
      const double u = tab[(col * ll + l) * zz + z], uz = tab[nt + (col * ll + l) * zz + z];
      const float4 t = stats[(((size_t)a * ll + l) * xx + X) * yy + Y];
      avg += t.x * u; aoi += t.y * u; gx += t.z * u; gy += t.w * u; gz += t.x * uz;
    }
    double v;
    if (simple) {
      v = avg;
    } else {
      gz *= tab[2 * nt + col * zz + z];
      const double g[3] = {gx, gy, gz};
      const double ss = gx * gx + gy * gy + gz * gz;
      const double pii = g[a] * g[a] / (ss == 0 ? 1.0 : ss);
      v = 1.0 / (pii * aoi + (1.0 - pii) * (1.0 / avg));
    }
    out[i] = (float)v;
  }
}

// ---- backward pass (the reference differentiates pjz.render with jax.grad,
// /root/reference/tests/test_layers.py:179-188) ----------------------------------------------------
//
//   render_combine_bwd_kernel : per (component, layer, X, Y): sweeps z, re-forms the forward sums of
//                       render_combine_kernel, and accumulates d(loss)/d(tile statistics) of its
//                       own layer; the gradient with respect to the overlap table (u, u*z) -- the
//                       path to layer_pos -- is reduced per block in shared memory and added to
//                       the (2, 2, ll, zz) result with one double atomic per entry and block.
//   tile_stats_bwd_kernel : per layer pixel, gathers the contributions of every tile sample that
//                       read it (the half-cell shifts are edge-replicated, so pixel 0 is read
//                       m + 1 times and the last m pixels never) -- no atomics.

// Partial derivatives of out = v(avg, aoi, gx, gy, gz) for component a, times the upstream gradient.
struct RenderPartials { double d_avg, d_aoi, d_gx, d_gy, d_gz; };

__device__ __forceinline__ RenderPartials render_partials(int a, int simple, double G, double avg,
                                                          double aoi, double gx, double gy, double gz) {
  RenderPartials r = {0, 0, 0, 0, 0};
  if (simple) { r.d_avg = G; return r; }
  const double g[3] = {gx, gy, gz};
  const double ss = gx * gx + gy * gy + gz * gz;
  const double ssd = ss == 0 ? 1.0 : ss;
  const double pii = g[a] * g[a] / ssd;
  const double ioa = 1.0 / avg;
  const double v = 1.0 / (pii * aoi + (1.0 - pii) * ioa);
  const double Gv = -G * v * v;                    // d loss / d D,  D = pii aoi + (1 - pii) / avg
  r.d_aoi = Gv * pii;
  r.d_avg = -Gv * (1.0 - pii) * ioa * ioa;
  const double dP = Gv * (aoi - ioa);
  if (ss != 0) {
    double dg[3];
#pragma unroll
    for (int b = 0; b < 3; ++b) dg[b] = dP * (-g[a] * g[a] * 2.0 * g[b] / (ssd * ssd));
    dg[a] += dP * 2.0 * g[a] / ssd;
    r.d_gx = dg[0]; r.d_gy = dg[1]; r.d_gz = dg[2];
  }
  return r;
}

// grid: (chunks of X*Y, 3 * ll); dynamic shared memory: 2 * zz doubles.
__global__ void __launch_bounds__(256)
render_combine_bwd_kernel(int ll, int xx, int yy, int zz, const float4* __restrict__ stats,
                          const double* __restrict__ tab, int simple,
                          const float* __restrict__ gout, double* __restrict__ dstats,
                          double* __restrict__ dtab) {
  extern __shared__ double sred[];                 // [zz] d/du, [zz] d/d(u z) of this block's (a, l)
  const int a = blockIdx.y / ll, l = blockIdx.y % ll;
  const int col = a == 2 ? 1 : 0;
  const int nt = 2 * ll * zz;
  for (int z = threadIdx.x; z < 2 * zz; z += blockDim.x) sred[z] = 0.0;
  __syncthreads();
  const int nxy = xx * yy;
  for (int base = blockIdx.x * blockDim.x; base < nxy; base += gridDim.x * blockDim.x) {
    const int xy = base + threadIdx.x;
    const bool live = xy < nxy;
    const size_t so = ((size_t)a * ll) * nxy + (live ? xy : 0);       // + l' * nxy
    const float4 mine = stats[so + (size_t)l * nxy];
    double dsx = 0, dsy = 0, dsz = 0, dsw = 0;
    for (int z = 0; z < zz; ++z) {
      double avg = 0, aoi = 0, gx = 0, gy = 0, gz = 0;
      for (int k = 0; k < ll; ++k) {
        const double u = tab[(col * ll + k) * zz + z], uz = tab[nt + (col * ll + k) * zz + z];
        const float4 t = stats[so + (size_t)k * nxy];
        avg += t.x * u; aoi += t.y * u; gx += t.z * u; gy += t.w * u; gz += t.x * uz;
      }
      const double c = tab[2 * nt + col * zz + z];
      gz *= c;
      const double G = live ? (double)gout[((size_t)a * nxy + (live ? xy : 0)) * zz + z] : 0.0;
      const RenderPartials r = render_partials(a, simple, G, avg, aoi, gx, gy, gz);
      const double u = tab[(col * ll + l) * zz + z], uz = tab[nt + (col * ll + l) * zz + z];
      dsx += r.d_avg * u + r.d_gz * c * uz;
      dsy += r.d_aoi * u;
      dsz += r.d_gx * u;
      dsw += r.d_gy * u;
      double du = r.d_avg * mine.x + r.d_aoi * mine.y + r.d_gx * mine.z + r.d_gy * mine.w;
      double duz = r.d_gz * c * mine.x;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        du += __shfl_xor_sync(0xffffffffu, du, o);
        duz += __shfl_xor_sync(0xffffffffu, duz, o);
      }
      if ((threadIdx.x & 31) == 0) { atomicAdd(&sred[z], du); atomicAdd(&sred[zz + z], duz); }
    }
    if (live) {
      double* d = dstats + 4 * (so + (size_t)l * nxy);
      d[0] = dsx; d[1] = dsy; d[2] = dsz; d[3] = dsw;
    }
  }
  __syncthreads();
  for (int z = threadIdx.x; z < zz; z += blockDim.x) {
    atomicAdd(&dtab[(col * ll + l) * zz + z], sred[z]);
    atomicAdd(&dtab[nt + (col * ll + l) * zz + z], sred[zz + z]);
  }
}

__global__ void __launch_bounds__(256)
tile_stats_bwd_kernel(int ll, int xx, int yy, int m, const float* __restrict__ layers,
                      const double* __restrict__ dstats, float* __restrict__ dlayers) {
  const int W = 2 * m, LX = W * xx, LY = W * yy;
  const size_t n = (size_t)ll * LX * LY;
  const double inv = 1.0 / ((double)W * W);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const int pj = (int)(i % LY);
    size_t r = i / LY;
    const int pi = (int)(r % LX);
    const int l = (int)(r / LX);
    const double v = layers[i];
    double acc = 0;
    for (int a = 0; a < 3; ++a) {
      const int sx = a != 0 ? m : 0, sy = a != 1 ? m : 0;
      // tile samples t = X*W + d that read this pixel: max(t - shift, 0) == p
      const int tx0 = pi == 0 ? 0 : pi + sx, tx1 = pi + sx;
      const int ty0 = pj == 0 ? 0 : pj + sy, ty1 = pj + sy;
      for (int tx = tx0; tx <= tx1 && tx < LX; ++tx) {
        const int X = tx / W, di = tx % W;
        const double wi = (di - (m - 0.5)) * inv;
        for (int ty = ty0; ty <= ty1 && ty < LY; ++ty) {
          const int Y = ty / W, dj = ty % W;
          const double wj = (dj - (m - 0.5)) * inv;
          const double* d = dstats + 4 * ((((size_t)a * ll + l) * xx + X) * yy + Y);
          acc += inv * (d[0] - d[1] / (v * v) + 12.0 * W * (wi * d[2] + wj * d[3]));
        }
      }
    }
    dlayers[i] = (float)acc;
  }
}

}  // namespace b200
