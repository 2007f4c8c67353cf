// Systolic kernel, second generation: same decomposition and dependency protocol as
// kernels_systolic.cuh (read its header first), re-plumbed so that neither memory latency nor
// inter-CTA synchronisation sits on the compute warps' critical path.
//
//  * Operand staging.  Every compute thread copies the 16-byte vectors of ITS OWN cells
//    (E^n[P+1], H^{n-1/2}[P], psi[P]) from global/L2 into a shared-memory ring with
//    cp.async.cg (LDGSTS, L1-bypassing, no registers held) D iterations before they are used.
//    x+1, y+1 and cross-warp z+1 neighbours are then plain shared-memory reads of the
//    neighbouring threads' slots: there is no exchange copy for E at all.
//      ring depth:  E needs D+2 plane slots (P and P+1 are both live), H and psi D+1.
//  * A dedicated sync warp (the last warp of the CTA) polls the predecessor stage's progress
//    counters D iterations ahead and publishes this CTA's own progress with st.release.gpu
//    after the barrier that follows the stores, so compute warps never execute an acquire
//    load, a fence or a spin.
//  * Two CTA barriers per plane: (A) "ring slot landed + previous stores issued",
//    (B) compute-warps-only exchange of the freshly formed H for the y-1 / z-1 neighbours.
#pragma once

#include <stdio.h>

#include <string>

#include "fdtd_common.cuh"
#include "kernels_systolic.cuh"

namespace b200 {

constexpr int kSys2MaxCompute = 512;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bar_compute(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

template <typename T, int D>
// 17 warps x 120 registers = 65 280 <= 64 Ki: __launch_bounds__(544) would round the CTA up to 20
// warps and cap the kernel at 96 registers (spills).
__global__ void __maxnreg__(120)
systolic2_kernel(const Geom g, const Ptrs<T> p, const SystolicCfg cfg, unsigned* sync) {
  constexpr int VW = VecTraits<T>::VW;
  constexpr int NE = D + 2, NH = D + 1;
  constexpr int PV = VW / 4;                     // float4 per psi vector
  extern __shared__ float4 smem[];
  const int NTc = blockDim.x - 32;               // compute threads
  const int tid = threadIdx.x, lane = tid & 31;
  const bool is_sync_warp = tid >= NTc;
  const int S = cfg.stages, NT = cfg.ntiles;
  const int t = blockIdx.x % NT, j = blockIdx.x / NT;
  const int y0 = (int)((long long)t * g.Y / NT);
  const int Yt = (int)((long long)(t + 1) * g.Y / NT) - y0;
  const int npsi = (cfg.tile_y + 2) * g.npg * PV; // float4 per psi array per slot

  float4* const sE = smem;                                   // [NE][3][NTc]
  float4* const sH = sE + (size_t)NE * 3 * NTc;              // [NH][3][NTc]
  float4* const sX = sH + (size_t)NH * 3 * NTc;              // [3][NTc]  Hz, Hx, Hy (new)
  float4* const sP = sX + (size_t)3 * NTc;                   // [NH][4][npsi]
  __shared__ int s_ok;

  unsigned* const status = sync + (size_t)S * NT * kSysFlagStride;
  unsigned* const my_prog = sync + ((size_t)j * NT + t) * kSysFlagStride;

  // =================================== sync warp =================================================
  if (is_sync_warp) {
    const int jp = (j + S - 1) % S, jn = (j + 1) % S;
    const unsigned* watch = nullptr;
    if (lane < 3) watch = sync + ((size_t)jp * NT + wrapi(t - 1 + lane, NT)) * kSysFlagStride;
    else if (lane == 3) watch = sync + ((size_t)jn * NT + t) * kSysFlagStride;
    bool ok = true;
    for (int n = j; n < g.tt && ok; n += S) {
      const int m = n / S;
      const unsigned base_prev = (unsigned)((j > 0 ? m : m - 1)) * (unsigned)g.X;
      const unsigned base_mine = (unsigned)m * (unsigned)g.X;
      const bool has_prev = n > 0, has_next = n + 1 < g.tt;
      // Iteration i (0 = prologue) works on sweep index k = i-1.  Before barrier A_i the loads of
      // iteration i+D are about to be issued: they touch planes up to index (i+D-1)+2 of the
      // previous stage's sweep.  The pre-loop (i = -1) covers the groups of iterations 0..D-1.
      for (int i = -1; i <= g.X && ok; ++i) {
        const int ahead = i + D;                   // last iteration whose loads get issued
        if (lane < 3 && has_prev) {
          const int need = min(max(ahead, 0) + 2, g.X);
          ok = wait_ge(watch, base_prev + (unsigned)need, status);
        } else if (lane == 3 && has_next && j + 1 < S) {
          const int kk = min(ahead, g.X) - 1;      // index this CTA is about to prefetch
          if (kk > cfg.max_lead)
            ok = wait_ge(watch, base_mine + (unsigned)(kk - cfg.max_lead), status);
        }
        ok = __all_sync(0xffffffffu, ok);
        if (lane == 0) s_ok = ok;
        __syncthreads();                           // A_i  (i = -1: the pre-loop barrier)
        // all stores of iteration i-1 (index i-2) are issued: publish it.
        if (lane == 0 && i >= 2) st_release_u32(my_prog, base_mine + (unsigned)(i - 1));
      }
      __syncthreads();                             // end of sweep: last iteration's stores issued
      if (lane == 0 && ok) st_release_u32(my_prog, base_mine + (unsigned)g.X);
    }
    return;
  }

  // ================================= compute warps ===============================================
  const int c = tid / g.Zq, q = tid - c * g.Zq;
  const bool active = c < Yt + 2;
  const bool doH = c <= Yt;
  const bool own = c >= 1 && c <= Yt;
  const int y = wrapi(y0 - 1 + (active ? c : 0), g.Y);
  const size_t coff = ((size_t)y * g.Zq + q) * VW;
  const int slot = psi_slot(g, q);
  const bool has_psi = slot >= 0;
  const size_t poff = ((size_t)y * g.npg + (has_psi ? slot : 0)) * VW;
  const size_t pplane = (size_t)g.Y * g.npg * VW;
  const int pidx = (c * g.npg + (has_psi ? slot : 0)) * PV;   // float4 index inside a psi slot
  const bool fix_up = cfg.need_zfix && lane == 31 && q + 1 < g.Zq;
  const bool fix_dn = cfg.need_zfix && lane == 0 && q > 0;
  const bool top = q + 1 == g.Zq, bottom = q == 0;
  const size_t XY = (size_t)g.X * g.Y;

  // CPML tables of this z-group are re-read from L1 (ld.global.nc, 3 KB total) where used:
  // holding all six in registers costs 6*VW registers for the whole kernel.
  auto load_tab = [&](int which, float (&dst)[VW]) {
#pragma unroll
    for (int v = 0; v < VW; v += 4) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(p.tab + which * g.Zp + q * VW + v));
      dst[v] = r.x; dst[v + 1] = r.y; dst[v + 2] = r.z; dst[v + 3] = r.w;
    }
  };

  for (int n = j; n < g.tt; n += S) {
    const int rb = n & 1;
    const T* const Er0 = rb ? p.E2[0] : p.E[0];
    const T* const Er1 = rb ? p.E2[1] : p.E[1];
    const T* const Er2 = rb ? p.E2[2] : p.E[2];
    const T* const Hr0 = rb ? p.H2[0] : p.H[0];
    const T* const Hr1 = rb ? p.H2[1] : p.H[1];
    const T* const Hr2 = rb ? p.H2[2] : p.H[2];
    T* const Ew0 = rb ? p.E[0] : p.E2[0];
    T* const Ew1 = rb ? p.E[1] : p.E2[1];
    T* const Ew2 = rb ? p.E[2] : p.E2[2];
    T* const Hw0 = rb ? p.H[0] : p.H2[0];
    T* const Hw1 = rb ? p.H[1] : p.H2[1];
    T* const Hw2 = rb ? p.H[2] : p.H2[2];
    const float* const pHr0 = rb ? p.psiH2[0] : p.psiH[0];
    const float* const pHr1 = rb ? p.psiH2[1] : p.psiH[1];
    float* const pHw0 = rb ? p.psiH[0] : p.psiH2[0];
    float* const pHw1 = rb ? p.psiH[1] : p.psiH2[1];
    const int cstart = n % g.X;
    const int oi = snapshot_index(g, n);
    const float w0 = __ldg(p.wave + 2 * (size_t)n), w1 = __ldg(p.wave + 2 * (size_t)n + 1);

    // Issues the async copies that iteration `it` consumes: E[P_it + 1], H[P_it], psi[P_it]
    // (+ E[P_0] for the very first one).  P_it = cstart - 1 + it.
    auto issue = [&](int it) {
      if (it <= g.X && active) {
        const int P = wrapi(cstart - 1 + it - (it > g.X ? g.X : 0), g.X);
        const int Pn = wrapi(P + 1, g.X);
        const size_t offP = (size_t)P * g.P + coff, offN = (size_t)Pn * g.P + coff;
        float4* e = sE + (size_t)((it + 1) % NE) * 3 * NTc + tid;
        cp_async16(e, Er0 + offN);
        cp_async16(e + 2 * NTc, Er2 + offN);
        if (doH) cp_async16(e + NTc, Er1 + offN);
        if (it == 0) {
          float4* e0 = sE + tid;                   // slot 0
          cp_async16(e0, Er0 + offP);
          cp_async16(e0 + 2 * NTc, Er2 + offP);
          if (doH) cp_async16(e0 + NTc, Er1 + offP);
        }
        if (doH) {
          float4* h = sH + (size_t)(it % NH) * 3 * NTc + tid;
          cp_async16(h, Hr0 + offP);
          cp_async16(h + NTc, Hr1 + offP);
          cp_async16(h + 2 * NTc, Hr2 + offP);
          if (has_psi) {
            float4* ps = sP + (size_t)(it % NH) * 4 * npsi + pidx;
            const size_t po = (size_t)P * pplane + poff;
#pragma unroll
            for (int v = 0; v < PV; ++v) {
              cp_async16(ps + v, pHr0 + po + 4 * v);
              cp_async16(ps + npsi + v, pHr1 + po + 4 * v);
              if (own && it >= 1) {
                cp_async16(ps + 2 * npsi + v, p.psiE[0] + po + 4 * v);
                cp_async16(ps + 3 * npsi + v, p.psiE[1] + po + 4 * v);
              }
            }
          }
        }
      }
      cp_async_commit();
    };

    __syncthreads();                               // A_{-1}: dependencies of iterations 0..D-1
    if (!s_ok) break;
#pragma unroll
    for (int it = 0; it < D; ++it) issue(it);

    float hyp[VW], hzp[VW];
#pragma unroll
    for (int i = 0; i < VW; ++i) { hyp[i] = 0.f; hzp[i] = 0.f; }

    for (int i = 0; i <= g.X; ++i) {               // i = 0 is the prologue plane cstart-1
      const int P = wrapi(cstart - 1 + i - (i > g.X ? g.X : 0), g.X);
      const size_t offP = (size_t)P * g.P + coff;
      const bool real = i >= 1;
      cp_async_wait<D - 1>();
      __syncthreads();                             // A_i
      if (!s_ok) break;
      issue(i + D);

      float4 bb0, bb1, bb2;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      bb0 = bb1 = bb2 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (own && real) {
        bb0 = ld16<LD_NC>(p.B[0] + offP);
        bb1 = ld16<LD_NC>(p.B[1] + offP);
        bb2 = ld16<LD_NC>(p.B[2] + offP);
        const size_t xy = (size_t)P * g.Y + y;
        a0 = __ldg(p.A + xy); a1 = __ldg(p.A + XY + xy); a2 = __ldg(p.A + 2 * XY + xy);
      }
      const float4* eC = sE + (size_t)(i % NE) * 3 * NTc;         // E^n[P]
      const float4* eN = sE + (size_t)((i + 1) % NE) * 3 * NTc;   // E^n[P+1]
      const float4* hO = sH + (size_t)(i % NH) * 3 * NTc;         // H^{n-1/2}[P]
      const float4* pS = sP + (size_t)(i % NH) * 4 * npsi + pidx;

      float ex[VW], ey[VW], ez[VW], ez_yp[VW], ex_yp[VW], ey_xp[VW], ez_xp[VW];
      float hx[VW], hy[VW], hz[VW], psx[VW], psy[VW];
      const int nb = doH ? tid + g.Zq : tid;
      unpack(eC[tid], ex, T()); unpack(eC[NTc + tid], ey, T()); unpack(eC[2 * NTc + tid], ez, T());
      unpack(eC[2 * NTc + nb], ez_yp, T()); unpack(eC[nb], ex_yp, T());
      unpack(eN[NTc + tid], ey_xp, T()); unpack(eN[2 * NTc + tid], ez_xp, T());
      unpack(hO[tid], hx, T()); unpack(hO[NTc + tid], hy, T()); unpack(hO[2 * NTc + tid], hz, T());
#pragma unroll
      for (int v = 0; v < VW; ++v) { psx[v] = 0.f; psy[v] = 0.f; }
      if (has_psi && doH) {
#pragma unroll
        for (int v = 0; v < PV; ++v) {
          float4 r = pS[v];
          psx[4 * v] = r.x; psx[4 * v + 1] = r.y; psx[4 * v + 2] = r.z; psx[4 * v + 3] = r.w;
          r = pS[npsi + v];
          psy[4 * v] = r.x; psy[4 * v + 1] = r.y; psy[4 * v + 2] = r.z; psy[4 * v + 3] = r.w;
        }
      }
      float ah[VW], bh[VW], ikh[VW];
      load_tab(3, ah); load_tab(4, bh); load_tab(5, ikh);
      float ex_top = __shfl_down_sync(0xffffffffu, ex[0], 1);
      float ey_top = __shfl_down_sync(0xffffffffu, ey[0], 1);
      if (fix_up) {
        float tmp[VW];
        unpack(eC[tid + 1], tmp, T()); ex_top = tmp[0];
        unpack(eC[NTc + tid + 1], tmp, T()); ey_top = tmp[0];
      }
      if (top) { ex_top = 0.f; ey_top = 0.f; }
#pragma unroll
      for (int v = 0; v < VW; ++v) {
        const float exz = (v + 1 < VW) ? ex[(v + 1) % VW] : ex_top;
        const float eyz = (v + 1 < VW) ? ey[(v + 1) % VW] : ey_top;
        h_cell(ex[v], ey[v], ez[v], exz, eyz, ez_yp[v], ex_yp[v], ey_xp[v], ez_xp[v], ah[v], bh[v],
               ikh[v], g.dt, psx[v], psy[v], hx[v], hy[v], hz[v]);
        hx[v] = round_store<T>(hx[v]); hy[v] = round_store<T>(hy[v]); hz[v] = round_store<T>(hz[v]);
      }

      if (real) {
        const float4 hxv = pack(hx, T()), hyv = pack(hy, T()), hzv = pack(hz, T());
        sX[tid] = hzv;
        sX[NTc + tid] = hxv;
        if (cfg.need_zfix) sX[2 * NTc + tid] = hyv;
        bar_compute(NTc);                          // B_i
        float hz_ym[VW], hx_ym[VW];
        const int nm = own ? tid - g.Zq : tid;
        unpack(sX[nm], hz_ym, T());
        unpack(sX[NTc + nm], hx_ym, T());
        float hx_bot = __shfl_up_sync(0xffffffffu, hx[VW - 1], 1);
        float hy_bot = __shfl_up_sync(0xffffffffu, hy[VW - 1], 1);
        if (fix_dn) {
          float tmp[VW];
          unpack(sX[NTc + tid - 1], tmp, T()); hx_bot = tmp[VW - 1];
          unpack(sX[2 * NTc + tid - 1], tmp, T()); hy_bot = tmp[VW - 1];
        }
        if (bottom) { hx_bot = 0.f; hy_bot = 0.f; }
        if (own) {
          float qsx[VW], qsy[VW], b0[VW], b1[VW], b2[VW];
#pragma unroll
          for (int v = 0; v < VW; ++v) { qsx[v] = 0.f; qsy[v] = 0.f; }
          if (has_psi) {
#pragma unroll
            for (int v = 0; v < PV; ++v) {
              float4 r = pS[2 * npsi + v];
              qsx[4 * v] = r.x; qsx[4 * v + 1] = r.y; qsx[4 * v + 2] = r.z; qsx[4 * v + 3] = r.w;
              r = pS[3 * npsi + v];
              qsy[4 * v] = r.x; qsy[4 * v + 1] = r.y; qsy[4 * v + 2] = r.z; qsy[4 * v + 3] = r.w;
            }
          }
          unpack(bb0, b0, T()); unpack(bb1, b1, T()); unpack(bb2, b2, T());
          float ae[VW], be[VW], ike[VW];
          load_tab(0, ae); load_tab(1, be); load_tab(2, ike);
#pragma unroll
          for (int v = 0; v < VW; ++v) {
            const float hxz = (v > 0) ? hx[(v + VW - 1) % VW] : hx_bot;
            const float hyz = (v > 0) ? hy[(v + VW - 1) % VW] : hy_bot;
            e_cell(hx[v], hy[v], hz[v], hxz, hyz, hz_ym[v], hx_ym[v], hyp[v], hzp[v], ae[v], be[v],
                   ike[v], a0, a1, a2, b0[v], b1[v], b2[v], qsx[v], qsy[v], ex[v], ey[v], ez[v]);
          }
          add_source<VW>(g, p.src, w0, w1, P, y, q, ex, ey, ez);
          st16<LD_CG>(Hw0 + offP, hxv);
          st16<LD_CG>(Hw1 + offP, hyv);
          st16<LD_CG>(Hw2 + offP, hzv);
          store_vec<T, LD_CG>(Ew0 + offP, ex);
          store_vec<T, LD_CG>(Ew1 + offP, ey);
          store_vec<T, LD_CG>(Ew2 + offP, ez);
          if (has_psi) {
            const size_t po = (size_t)P * pplane + poff;
#pragma unroll
            for (int v = 0; v < VW; v += 4) {
              __stcg(reinterpret_cast<float4*>(pHw0 + po + v),
                     make_float4(psx[v], psx[v + 1], psx[v + 2], psx[v + 3]));
              __stcg(reinterpret_cast<float4*>(pHw1 + po + v),
                     make_float4(psy[v], psy[v + 1], psy[v + 2], psy[v + 3]));
              __stcg(reinterpret_cast<float4*>(p.psiE[0] + po + v),
                     make_float4(qsx[v], qsx[v + 1], qsx[v + 2], qsx[v + 3]));
              __stcg(reinterpret_cast<float4*>(p.psiE[1] + po + v),
                     make_float4(qsy[v], qsy[v + 1], qsy[v + 2], qsy[v + 3]));
            }
          }
          if (oi >= 0) {
#pragma unroll
            for (int v = 0; v < VW; ++v) {
              ex[v] = round_store<T>(ex[v]); ey[v] = round_store<T>(ey[v]);
              ez[v] = round_store<T>(ez[v]);
            }
            write_snapshot<VW>(g, p.out, oi, P, y, q, ex, ey, ez);
          }
        }
      }
#pragma unroll
      for (int v = 0; v < VW; ++v) { hyp[v] = hy[v]; hzp[v] = hz[v]; }
    }
    cp_async_wait<0>();
    __syncthreads();                               // end of sweep (pairs with the sync warp)
    if (!s_ok) break;
  }
}

template <typename T, int D>
size_t systolic2_smem_bytes(const Geom& g, int compute_threads, int tile_y) {
  constexpr int PV = VecTraits<T>::VW / 4;
  const size_t npsi = (size_t)(tile_y + 2) * g.npg * PV;
  return sizeof(float4) * ((size_t)(D + 2) * 3 * compute_threads + (size_t)(D + 1) * 3 * compute_threads +
                           3 * (size_t)compute_threads + (size_t)(D + 1) * 4 * npsi);
}

template <typename T, int D>
bool systolic2_configure_d(const Geom& g, int tile_y_req, int stages_req, int threads_req, int sms,
                           int l2_bytes, SystolicCfg* cfg, std::string* why) {
  const int max_threads = threads_req > 0 ? (threads_req < kSys2MaxCompute ? threads_req : kSys2MaxCompute)
                                          : kSys2MaxCompute;
  if (g.Zq * 3 > max_threads) { *why = "z extent too large for one CTA"; return false; }
  int max_tile = max_threads / g.Zq - 2;
  if (tile_y_req > 0 && tile_y_req < max_tile) max_tile = tile_y_req;
  if (max_tile > g.Y) max_tile = g.Y;
  const int ntiles = (g.Y + max_tile - 1) / max_tile;
  const int widest = (g.Y + ntiles - 1) / ntiles;
  const int compute = ((widest + 2) * g.Zq + 31) / 32 * 32;
  cfg->tile_y = widest;
  cfg->ntiles = ntiles;
  cfg->threads = compute + 32;
  cfg->need_zfix = (32 % g.Zq != 0);
  cfg->smem_bytes = (int)systolic2_smem_bytes<T, D>(g, compute, widest);
  // >= 2D+4 is needed for deadlock freedom (a throttled stage has published i-2 indices while
  // its successor needs i'+D+2 of them to advance).
  cfg->max_lead = 2 * D + 6;
  cfg->trap_on_timeout = 1;
  int occ = 0;
  if (cfg->smem_bytes > 227 * 1024 ||
      cudaFuncSetAttribute(systolic2_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           cfg->smem_bytes) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, systolic2_kernel<T, D>, cfg->threads,
                                                    cfg->smem_bytes) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    *why = "staging ring does not fit in shared memory";
    return false;
  }
  const long long capacity = (long long)occ * sms;
  if (ntiles > capacity) { *why = "more y-tiles than co-resident CTAs"; return false; }
  int stages = (int)(capacity / ntiles);
  const long long plane_bytes = g.P * (long long)sizeof(T) * 15;
  long long by_l2 = (long long)(l2_bytes * 0.5) / ((3 + D) * plane_bytes);
  if (by_l2 < 1) by_l2 = 1;
  if (stages > by_l2) stages = (int)by_l2;
  if (stages_req > 0 && stages_req <= capacity / ntiles) stages = stages_req;
  if (stages > g.tt) stages = g.tt > 0 ? g.tt : 1;
  if (stages > g.X) stages = g.X;
  cfg->stages = stages;
  cfg->l2_window_bytes = (long long)stages * (3 + D) * plane_bytes;
  return true;
}

template <typename T, int D>
int systolic2_launch_d(const Geom& g, const Ptrs<T>& p, const SystolicCfg& cfg, unsigned* sync,
                       cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(systolic2_kernel<T, D>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, cfg.smem_bytes);
  if (e != cudaSuccess) return (int)e;
  Geom gg = g;
  Ptrs<T> pp = p;
  SystolicCfg cc = cfg;
  void* args[] = {&gg, &pp, &cc, &sync};
  e = cudaLaunchCooperativeKernel((const void*)systolic2_kernel<T, D>,
                                  dim3(cfg.stages * cfg.ntiles), dim3(cfg.threads), args,
                                  cfg.smem_bytes, st);
  if (e != cudaSuccess) return (int)e;
  systolic_check_kernel<<<1, 1, 0, st>>>(sync + (size_t)cfg.stages * cfg.ntiles * kSysFlagStride);
  return (int)cudaGetLastError();
}

}  // namespace b200
