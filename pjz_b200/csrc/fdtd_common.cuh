// Shared device-side pieces of the B200 FDTD engine: geometry, 16-byte vector I/O, and the
// per-cell Yee/CPML arithmetic.
//
// The per-cell operation order below is THE engine arithmetic.  oracle/fdtd_c.c restates it
// with the same explicit fmaf() calls, and the library is compiled with --fmad=false, so the
// GPU result can be compared bit-for-bit with the CPU oracle.  Semantics: SURVEY.md 8(c) /
// oracle/fdtd_numpy.py (replaces the engine behind /root/reference/src/pjz/_field.py:254-269).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

// Everything a kernel needs to know about one run; passed by value.
struct Geom {
  int X, Y, Z;         // full domain
  int Zp, Zq;          // z padded to the vector width; vector groups per z-column
  int nlo, hi0, npg;   // PML groups: q < nlo (bottom), q >= hi0 (top); npg = psi groups/column
  int xx, yy, zz;      // epsilon / output sub-volume
  int ox, oy, oz;      // ... and its offset
  int src_axis, src_pos;
  int out_start, out_stop, out_step;
  int proj_rows, n_out; // fused projection: rows of W (0 = plain snapshots), snapshots per run
  int tt;              // one past the last step of this launch
  int n0;              // first step of this launch (stepping sessions; 0 for a whole run)
  int ylo, yhi;        // y-columns the tiles cover: [0, Y), or the owned range of a slab whose
                       // ghost columns ylo-1 and yhi are filled by the neighbouring GPUs (lean kernel)
  float dt;
  long long P;         // elements per x-plane   = Y * Zp
  long long N;         // elements per component = X * P
};

// Device pointers of one run (T = float or __half for the field/coefficient storage).
template <typename T>
struct Ptrs {
  T* E[3];             // Ex,Ey,Ez   [X][Y][Zp]
  T* H[3];             // Hx,Hy,Hz
  T* E2[3];            // ping-pong copies (systolic kernel only)
  T* H2[3];
  const T* B[3];       // (dt/eps)/(1+s dt/2), 0 in the z padding
  const float* A;      // [3][X][Y]  (1-s dt/2)/(1+s dt/2)
  const float* A4;     // [X][Y][4]  the same three values packed per column (+ one pad word)
  const float* S4;     // [X][Y][4]  z-plane source packed per column: (ch0 Ex, ch0 Ey, ch1 Ex, ch1 Ey)
  const float* tab;    // [6][Zp]    a_e,b_e,ik_e,a_h,b_h,ik_h
  float* psiH[2];      // [X][Y][npg][VW]  (psiHx, psiHy), fp32 always
  float* psiH2[2];     // ping-pong copy (systolic kernel only)
  T* Es[2][3];         // the same pointers indexed [buffer set][component] (constant-bank lookup)
  T* Hs[2][3];
  float* psiHs[2][2];
  float* psiE[2];
  const float* src;    // source_field, caller layout
  const float* wave;   // (tt,2)
  float* out;          // (n_out,3,xx,yy,zz), or (proj_rows,3,xx,yy,zz) with the fused projection
  const float* proj;   // (proj_rows, n_out) projection weights, or nullptr
};

template <typename T> struct VecTraits;
template <> struct VecTraits<float> { static constexpr int VW = 4; };
template <> struct VecTraits<__half> { static constexpr int VW = 8; };

// Cache policy of global accesses.
//  LD_DEFAULT: plain ld.global (L1 allowed) -- per-step kernels, coherent across launches.
//  LD_CG     : ld.global.cg (L2 only)       -- data other CTAs rewrite during the same launch.
//  LD_NC     : ld.global.nc                 -- immutable for the whole run (coefficients).
enum { LD_DEFAULT = 0, LD_CG = 1, LD_NC = 2 };

template <int MODE>
__device__ __forceinline__ float4 ld16(const void* p) {
  if constexpr (MODE == LD_CG) return __ldcg(reinterpret_cast<const float4*>(p));
  else if constexpr (MODE == LD_NC) return __ldg(reinterpret_cast<const float4*>(p));
  else return *reinterpret_cast<const float4*>(p);
}

template <int MODE>
__device__ __forceinline__ void st16(void* p, float4 v) {
  if constexpr (MODE == LD_CG) __stcg(reinterpret_cast<float4*>(p), v);
  else *reinterpret_cast<float4*>(p) = v;
}

__device__ __forceinline__ void unpack(float4 r, float (&v)[4], float /*tag*/) {
  v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}
__device__ __forceinline__ void unpack(float4 r, float (&v)[8], __half /*tag*/) {
  const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ float4 pack(const float (&v)[4], float) {
  return make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ float4 pack(const float (&v)[8], __half) {
  float4 r;
  __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  return r;
}

// 16-byte vector of T at p -> VW floats.
template <typename T, int MODE>
__device__ __forceinline__ void load_vec(const T* p, float (&v)[VecTraits<T>::VW]) {
  unpack(ld16<MODE>(p), v, T());
}
template <typename T, int MODE>
__device__ __forceinline__ void store_vec(T* p, const float (&v)[VecTraits<T>::VW]) {
  st16<MODE>(p, pack(v, T()));
}

template <typename T, int MODE>
__device__ __forceinline__ float load_one(const T* p) {
  if constexpr (sizeof(T) == 4) {
    if constexpr (MODE == LD_CG) return __ldcg(reinterpret_cast<const float*>(p));
    else if constexpr (MODE == LD_NC) return __ldg(reinterpret_cast<const float*>(p));
    else return *reinterpret_cast<const float*>(p);
  } else {
    unsigned short u;
    if constexpr (MODE == LD_CG) u = __ldcg(reinterpret_cast<const unsigned short*>(p));
    else if constexpr (MODE == LD_NC) u = __ldg(reinterpret_cast<const unsigned short*>(p));
    else u = *reinterpret_cast<const unsigned short*>(p);
    return __half2float(__ushort_as_half(u));
  }
}

// Value a float takes once it has been through the storage type.
template <typename T>
__device__ __forceinline__ float round_store(float v) {
  if constexpr (sizeof(T) == 4) return v;
  else return __half2float(__float2half_rn(v));
}

__host__ __device__ __forceinline__ int wrapi(int i, int n) {
  return i < 0 ? i + n : (i >= n ? i - n : i);
}

// psi group slot of z-group q, or -1 if the group holds no PML cell.
__device__ __forceinline__ int psi_slot(const Geom& g, int q) {
  if (q < g.nlo) return q;
  if (q >= g.hi0) return g.nlo + (q - g.hi0);
  return -1;
}

// ---- per-cell arithmetic (same order as oracle/fdtd_c.c) ---------------------------------

// H half-step for one cell.  *_zp: value at z+1 (0 above the top), *_yp at y+1, *_xp at x+1.
__device__ __forceinline__ void h_cell(float ex, float ey, float ez, float ex_zp, float ey_zp,
                                       float ez_yp, float ex_yp, float ey_xp, float ez_xp,
                                       float ah, float bh, float ikh, float dt, float& psx,
                                       float& psy, float& hx, float& hy, float& hz) {
  const float dzEy = ey_zp - ey;
  const float dzEx = ex_zp - ex;
  psx = fmaf(bh, psx, ah * dzEy);
  psy = fmaf(bh, psy, ah * dzEx);
  const float cx = (ez_yp - ez) - fmaf(dzEy, ikh, psx);
  const float cy = fmaf(dzEx, ikh, psy) - (ez_xp - ez);
  const float cz = (ey_xp - ey) - (ex_yp - ex);
  hx = fmaf(-dt, cx, hx);
  hy = fmaf(-dt, cy, hy);
  hz = fmaf(-dt, cz, hz);
}

// E half-step for one cell.  *_zm: value at z-1 (0 below the bottom), *_ym at y-1, *_xm at x-1.
__device__ __forceinline__ void e_cell(float hx, float hy, float hz, float hx_zm, float hy_zm,
                                       float hz_ym, float hx_ym, float hy_xm, float hz_xm,
                                       float ae, float be, float ike, float a0, float a1,
                                       float a2, float b0, float b1, float b2, float& psx,
                                       float& psy, float& ex, float& ey, float& ez) {
  const float dzHy = hy - hy_zm;
  const float dzHx = hx - hx_zm;
  psx = fmaf(be, psx, ae * dzHy);
  psy = fmaf(be, psy, ae * dzHx);
  const float cx = (hz - hz_ym) - fmaf(dzHy, ike, psx);
  const float cy = fmaf(dzHx, ike, psy) - (hz - hz_xm);
  const float cz = (hy - hy_xm) - (hx - hx_ym);
  ex = fmaf(b0, cx, a0 * ex);
  ey = fmaf(b1, cy, a1 * ey);
  ez = fmaf(b2, cz, a2 * ez);
}

// Plane-source injection for the VW cells (x, y, q*VW .. q*VW+VW-1), channel 0 then channel 1.
// Source layouts are the caller's: (2,1,Y,Z) | (2,X,1,Z) | (2,2,X,Y,1).
template <int VW>
__device__ __forceinline__ void add_source(const Geom& g, const float* __restrict__ src,
                                           float w0, float w1, int x, int y, int q,
                                           float (&ex)[VW], float (&ey)[VW], float (&ez)[VW]) {
  // Written with selects instead of conditional element updates: a conditionally updated
  // register array makes ptxas demote the whole array to local memory (measured: 64 B of
  // local traffic per thread per plane in the systolic kernel).
  if (g.src_axis == 0 || g.src_axis == 1) {
    const bool ax0 = g.src_axis == 0;
    const int n = ax0 ? g.X : g.Y, pos = ax0 ? x : y, line = ax0 ? y : x;
    const float* s0 = src + (size_t)line * g.Z;
    const float* s1 = s0 + (size_t)(ax0 ? g.Y : g.X) * g.Z;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const bool plane_hit = pos == wrapi(g.src_pos - ch, n);
      const float w = ch ? w1 : w0;
#pragma unroll
      for (int i = 0; i < VW; ++i) {
        const int z = q * VW + i;
        const bool hit = plane_hit && z < g.Z;
        const float a = hit ? __ldg(s0 + z) : 0.f, b = hit ? __ldg(s1 + z) : 0.f;
        const float t0 = fmaf(w, a, ax0 ? ey[i] : ex[i]);
        const float t2 = fmaf(w, b, ez[i]);
        ey[i] = (hit && ax0) ? t0 : ey[i];
        ex[i] = (hit && !ax0) ? t0 : ex[i];
        ez[i] = hit ? t2 : ez[i];
      }
    }
  } else {
    const bool qhit = q == g.src_pos / VW;
    const int idx = g.src_pos % VW;
    const size_t XY = (size_t)g.X * g.Y, xy = (size_t)x * g.Y + y;
    const float s00 = __ldg(src + xy), s01 = __ldg(src + XY + xy);
    const float s10 = __ldg(src + 2 * XY + xy), s11 = __ldg(src + 3 * XY + xy);
#pragma unroll
    for (int i = 0; i < VW; ++i) {
      const bool hit = qhit && i == idx;
      float e0 = ex[i], e1 = ey[i];
      e0 = fmaf(w0, s00, e0);
      e1 = fmaf(w0, s01, e1);
      e0 = fmaf(w1, s10, e0);
      e1 = fmaf(w1, s11, e1);
      ex[i] = hit ? e0 : ex[i];
      ey[i] = hit ? e1 : ey[i];
    }
  }
}

// Snapshot of the VW cells (x, y, q*VW..) into out[oi] (n_out,3,xx,yy,zz), cropped.  With the
// fused projection (g.proj_rows = R > 0) the snapshot is not stored: every row r of the output
// (R,3,xx,yy,zz) accumulates W[r][oi] * E instead (one fmaf per row, snapshot order -- the order
// oracle/fdtd_c.c:oracle_fdtd_project restates).  Successive snapshot steps of one cell are
// ordered by the kernels' own step-to-step dependencies; the accumulators bypass L1.
template <int VW>
__device__ __forceinline__ void write_snapshot(const Geom& g, float* __restrict__ out, int oi,
                                               int x, int y, int q, const float (&ex)[VW],
                                               const float (&ey)[VW], const float (&ez)[VW],
                                               const float* __restrict__ proj = nullptr) {
  const int sx = x - g.ox, sy = y - g.oy;
  if (sx < 0 || sx >= g.xx || sy < 0 || sy >= g.yy) return;
  const size_t comp = (size_t)g.xx * g.yy * g.zz;
  if (g.proj_rows > 0) {
    float* o = out + ((size_t)sx * g.yy + sy) * g.zz;
    const int sz0 = q * VW - g.oz;
    if (VW == 4 && (g.oz & 3) == 0 && (g.zz & 3) == 0 && sz0 >= 0 && sz0 + VW <= g.zz) {
      // whole 16-byte vectors (aligned crop): one read-modify-write per component and row
      for (int r = 0; r < g.proj_rows; ++r, o += 3 * comp) {
        const float w = __ldg(proj + (size_t)r * g.n_out + oi);
        float4* const ox = reinterpret_cast<float4*>(o + sz0);
        float4* const oy = reinterpret_cast<float4*>(o + comp + sz0);
        float4* const oz_ = reinterpret_cast<float4*>(o + 2 * comp + sz0);
        float4 a = __ldcg(ox), b = __ldcg(oy), c = __ldcg(oz_);
        a.x = fmaf(w, ex[0], a.x); a.y = fmaf(w, ex[1], a.y); a.z = fmaf(w, ex[2], a.z); a.w = fmaf(w, ex[3], a.w);
        b.x = fmaf(w, ey[0], b.x); b.y = fmaf(w, ey[1], b.y); b.z = fmaf(w, ey[2], b.z); b.w = fmaf(w, ey[3], b.w);
        c.x = fmaf(w, ez[0], c.x); c.y = fmaf(w, ez[1], c.y); c.z = fmaf(w, ez[2], c.z); c.w = fmaf(w, ez[3], c.w);
        __stcg(ox, a); __stcg(oy, b); __stcg(oz_, c);
      }
      return;
    }
    for (int r = 0; r < g.proj_rows; ++r, o += 3 * comp) {
      const float w = __ldg(proj + (size_t)r * g.n_out + oi);
#pragma unroll
      for (int i = 0; i < VW; ++i) {
        const int sz = q * VW + i - g.oz;
        if (sz >= 0 && sz < g.zz) {
          __stcg(o + sz, fmaf(w, ex[i], __ldcg(o + sz)));
          __stcg(o + comp + sz, fmaf(w, ey[i], __ldcg(o + comp + sz)));
          __stcg(o + 2 * comp + sz, fmaf(w, ez[i], __ldcg(o + 2 * comp + sz)));
        }
      }
    }
    return;
  }
  float* o = out + (size_t)oi * 3 * comp + ((size_t)sx * g.yy + sy) * g.zz;
#pragma unroll
  for (int i = 0; i < VW; ++i) {
    const int sz = q * VW + i - g.oz;
    if (sz >= 0 && sz < g.zz) {
      o[sz] = ex[i];
      o[comp + sz] = ey[i];
      o[2 * comp + sz] = ez[i];
    }
  }
}

// Output index of step n, or -1.
__host__ __device__ __forceinline__ int snapshot_index(const Geom& g, int n) {
  if (n < g.out_start || n >= g.out_stop) return -1;
  const int d = n - g.out_start;
  return (d % g.out_step == 0) ? d / g.out_step : -1;
}

}  // namespace b200
