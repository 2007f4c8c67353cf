// Per-step kernels: one H launch + one E launch per time step, updating in place.
//
// One thread owns one 16-byte z-vector (VW cells) of one (x, y) column.  z is the fastest
// axis, so a warp reads 512 contiguous bytes per component; y+-1 and z+-1 neighbours come
// out of L1 (loaded by the neighbouring warps of the same CTA as their own cells), x+-1 out of
// L2 (the neighbouring CTA's own plane).  HBM traffic: H pass 9 words, E pass 12 words per
// cell = 84 B per cell-update in fp32 (+ psi in the PML groups).  This is the simple, always
// available path and the on-device cross-check for the systolic kernel.
#pragma once

#include "fdtd_common.cuh"

namespace b200 {

constexpr int kTwoPassThreads = 256;

// grid = (ceil(Y*Zq / blockDim), X)
template <typename T>
__global__ void __launch_bounds__(kTwoPassThreads)
twopass_h_kernel(Geom g, Ptrs<T> p) {
  constexpr int VW = VecTraits<T>::VW;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;   // vector index inside the plane
  if (f >= g.Y * g.Zq) return;
  const int x = blockIdx.y;
  const int y = f / g.Zq, q = f - y * g.Zq;
  const size_t base = (size_t)x * g.P + (size_t)f * VW;
  const size_t base_xp = (size_t)wrapi(x + 1, g.X) * g.P + (size_t)f * VW;
  const size_t base_yp = (size_t)x * g.P + ((size_t)wrapi(y + 1, g.Y) * g.Zq + q) * VW;

  float ex[VW], ey[VW], ez[VW], ez_yp[VW], ex_yp[VW], ey_xp[VW], ez_xp[VW];
  float hx[VW], hy[VW], hz[VW];
  load_vec<T, LD_DEFAULT>(p.E[0] + base, ex);
  load_vec<T, LD_DEFAULT>(p.E[1] + base, ey);
  load_vec<T, LD_DEFAULT>(p.E[2] + base, ez);
  load_vec<T, LD_DEFAULT>(p.E[2] + base_yp, ez_yp);
  load_vec<T, LD_DEFAULT>(p.E[0] + base_yp, ex_yp);
  load_vec<T, LD_DEFAULT>(p.E[1] + base_xp, ey_xp);
  load_vec<T, LD_DEFAULT>(p.E[2] + base_xp, ez_xp);
  load_vec<T, LD_DEFAULT>(p.H[0] + base, hx);
  load_vec<T, LD_DEFAULT>(p.H[1] + base, hy);
  load_vec<T, LD_DEFAULT>(p.H[2] + base, hz);
  float ex_top = 0.f, ey_top = 0.f;                      // value at z+1 of the last lane
  if (q + 1 < g.Zq) {
    ex_top = load_one<T, LD_DEFAULT>(p.E[0] + base + VW);
    ey_top = load_one<T, LD_DEFAULT>(p.E[1] + base + VW);
  }
  float ah[VW], bh[VW], ikh[VW];
  {
    float4 t;
#pragma unroll
    for (int v = 0; v < VW; v += 4) {
      t = __ldg(reinterpret_cast<const float4*>(p.tab + 3 * g.Zp + q * VW + v));
      ah[v] = t.x; ah[v + 1] = t.y; ah[v + 2] = t.z; ah[v + 3] = t.w;
      t = __ldg(reinterpret_cast<const float4*>(p.tab + 4 * g.Zp + q * VW + v));
      bh[v] = t.x; bh[v + 1] = t.y; bh[v + 2] = t.z; bh[v + 3] = t.w;
      t = __ldg(reinterpret_cast<const float4*>(p.tab + 5 * g.Zp + q * VW + v));
      ikh[v] = t.x; ikh[v + 1] = t.y; ikh[v + 2] = t.z; ikh[v + 3] = t.w;
    }
  }
  const int slot = psi_slot(g, q);
  float psx[VW], psy[VW];
#pragma unroll
  for (int i = 0; i < VW; ++i) { psx[i] = 0.f; psy[i] = 0.f; }
  size_t pbase = 0;
  if (slot >= 0) {
    pbase = (((size_t)x * g.Y + y) * g.npg + slot) * VW;
#pragma unroll
    for (int v = 0; v < VW; v += 4) {
      float4 t = *reinterpret_cast<const float4*>(p.psiH[0] + pbase + v);
      psx[v] = t.x; psx[v + 1] = t.y; psx[v + 2] = t.z; psx[v + 3] = t.w;
      t = *reinterpret_cast<const float4*>(p.psiH[1] + pbase + v);
      psy[v] = t.x; psy[v + 1] = t.y; psy[v + 2] = t.z; psy[v + 3] = t.w;
    }
  }
#pragma unroll
  for (int i = 0; i < VW; ++i) {
    const float exz = (i + 1 < VW) ? ex[(i + 1) % VW] : ex_top;
    const float eyz = (i + 1 < VW) ? ey[(i + 1) % VW] : ey_top;
    h_cell(ex[i], ey[i], ez[i], exz, eyz, ez_yp[i], ex_yp[i], ey_xp[i], ez_xp[i], ah[i], bh[i],
           ikh[i], g.dt, psx[i], psy[i], hx[i], hy[i], hz[i]);
  }
  store_vec<T, LD_DEFAULT>(p.H[0] + base, hx);
  store_vec<T, LD_DEFAULT>(p.H[1] + base, hy);
  store_vec<T, LD_DEFAULT>(p.H[2] + base, hz);
  if (slot >= 0) {
#pragma unroll
    for (int v = 0; v < VW; v += 4) {
      *reinterpret_cast<float4*>(p.psiH[0] + pbase + v) =
          make_float4(psx[v], psx[v + 1], psx[v + 2], psx[v + 3]);
      *reinterpret_cast<float4*>(p.psiH[1] + pbase + v) =
          make_float4(psy[v], psy[v + 1], psy[v + 2], psy[v + 3]);
    }
  }
}

// E update of step n (+ source injection, + snapshot when n is an output step).
template <typename T>
__global__ void __launch_bounds__(kTwoPassThreads)
twopass_e_kernel(Geom g, Ptrs<T> p, int n) {
  constexpr int VW = VecTraits<T>::VW;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= g.Y * g.Zq) return;
  const int x = blockIdx.y;
  const int y = f / g.Zq, q = f - y * g.Zq;
  const size_t base = (size_t)x * g.P + (size_t)f * VW;
  const size_t base_xm = (size_t)wrapi(x - 1, g.X) * g.P + (size_t)f * VW;
  const size_t base_ym = (size_t)x * g.P + ((size_t)wrapi(y - 1, g.Y) * g.Zq + q) * VW;

  float hx[VW], hy[VW], hz[VW], hz_ym[VW], hx_ym[VW], hy_xm[VW], hz_xm[VW];
  float ex[VW], ey[VW], ez[VW], b0[VW], b1[VW], b2[VW];
  load_vec<T, LD_DEFAULT>(p.H[0] + base, hx);
  load_vec<T, LD_DEFAULT>(p.H[1] + base, hy);
  load_vec<T, LD_DEFAULT>(p.H[2] + base, hz);
  load_vec<T, LD_DEFAULT>(p.H[2] + base_ym, hz_ym);
  load_vec<T, LD_DEFAULT>(p.H[0] + base_ym, hx_ym);
  load_vec<T, LD_DEFAULT>(p.H[1] + base_xm, hy_xm);
  load_vec<T, LD_DEFAULT>(p.H[2] + base_xm, hz_xm);
  load_vec<T, LD_DEFAULT>(p.E[0] + base, ex);
  load_vec<T, LD_DEFAULT>(p.E[1] + base, ey);
  load_vec<T, LD_DEFAULT>(p.E[2] + base, ez);
  load_vec<T, LD_NC>(p.B[0] + base, b0);
  load_vec<T, LD_NC>(p.B[1] + base, b1);
  load_vec<T, LD_NC>(p.B[2] + base, b2);
  float hx_bot = 0.f, hy_bot = 0.f;                      // value at z-1 of the first lane
  if (q > 0) {
    hx_bot = load_one<T, LD_DEFAULT>(p.H[0] + base - 1);
    hy_bot = load_one<T, LD_DEFAULT>(p.H[1] + base - 1);
  }
  const size_t xy = (size_t)x * g.Y + y, XY = (size_t)g.X * g.Y;
  const float a0 = __ldg(p.A + xy), a1 = __ldg(p.A + XY + xy), a2 = __ldg(p.A + 2 * XY + xy);
  float ae[VW], be[VW], ike[VW];
  {
    float4 t;
#pragma unroll
    for (int v = 0; v < VW; v += 4) {
      t = __ldg(reinterpret_cast<const float4*>(p.tab + 0 * g.Zp + q * VW + v));
      ae[v] = t.x; ae[v + 1] = t.y; ae[v + 2] = t.z; ae[v + 3] = t.w;
      t = __ldg(reinterpret_cast<const float4*>(p.tab + 1 * g.Zp + q * VW + v));
      be[v] = t.x; be[v + 1] = t.y; be[v + 2] = t.z; be[v + 3] = t.w;
      t = __ldg(reinterpret_cast<const float4*>(p.tab + 2 * g.Zp + q * VW + v));
      ike[v] = t.x; ike[v + 1] = t.y; ike[v + 2] = t.z; ike[v + 3] = t.w;
    }
  }
  const int slot = psi_slot(g, q);
  float psx[VW], psy[VW];
#pragma unroll
  for (int i = 0; i < VW; ++i) { psx[i] = 0.f; psy[i] = 0.f; }
  size_t pbase = 0;
  if (slot >= 0) {
    pbase = (xy * g.npg + slot) * VW;
#pragma unroll
    for (int v = 0; v < VW; v += 4) {
      float4 t = *reinterpret_cast<const float4*>(p.psiE[0] + pbase + v);
      psx[v] = t.x; psx[v + 1] = t.y; psx[v + 2] = t.z; psx[v + 3] = t.w;
      t = *reinterpret_cast<const float4*>(p.psiE[1] + pbase + v);
      psy[v] = t.x; psy[v + 1] = t.y; psy[v + 2] = t.z; psy[v + 3] = t.w;
    }
  }
#pragma unroll
  for (int i = 0; i < VW; ++i) {
    const float hxz = (i > 0) ? hx[(i + VW - 1) % VW] : hx_bot;
    const float hyz = (i > 0) ? hy[(i + VW - 1) % VW] : hy_bot;
    e_cell(hx[i], hy[i], hz[i], hxz, hyz, hz_ym[i], hx_ym[i], hy_xm[i], hz_xm[i], ae[i], be[i],
           ike[i], a0, a1, a2, b0[i], b1[i], b2[i], psx[i], psy[i], ex[i], ey[i], ez[i]);
  }
  const float w0 = __ldg(p.wave + 2 * (size_t)n), w1 = __ldg(p.wave + 2 * (size_t)n + 1);
  add_source<VW>(g, p.src, w0, w1, x, y, q, ex, ey, ez);
  store_vec<T, LD_DEFAULT>(p.E[0] + base, ex);
  store_vec<T, LD_DEFAULT>(p.E[1] + base, ey);
  store_vec<T, LD_DEFAULT>(p.E[2] + base, ez);
  if (slot >= 0) {
#pragma unroll
    for (int v = 0; v < VW; v += 4) {
      *reinterpret_cast<float4*>(p.psiE[0] + pbase + v) =
          make_float4(psx[v], psx[v + 1], psx[v + 2], psx[v + 3]);
      *reinterpret_cast<float4*>(p.psiE[1] + pbase + v) =
          make_float4(psy[v], psy[v + 1], psy[v + 2], psy[v + 3]);
    }
  }
  const int oi = snapshot_index(g, n);
  if (oi >= 0) {
#pragma unroll
    for (int i = 0; i < VW; ++i) {
      ex[i] = round_store<T>(ex[i]); ey[i] = round_store<T>(ey[i]); ez[i] = round_store<T>(ez[i]);
    }
    write_snapshot<VW>(g, p.out, oi, x, y, q, ex, ey, ez, p.proj);
  }
}

}  // namespace b200
