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

}  // namespace b200
