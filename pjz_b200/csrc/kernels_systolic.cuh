// Persistent "systolic" kernel: ALL tt time steps in ONE launch.
//
// Decomposition.  The grid is `stages x ntiles` CTAs, all co-resident (cooperative launch).
// CTA (j, t) owns y-tile t (tile_y columns, all of z) and performs the time steps
// n = j, j+stages, j+2*stages, ...  One time step ("stage") is a full sweep over the X planes
// of a FUSED H+E update: at plane P the CTA
//     loads   E^n[P+1], H^{n-1/2}[P], coefficients[P]              (each exactly once)
//     forms   H^{n+1/2}[P]  from E^n[P], E^n[P+1]                  (x+1 = the thread's own next load)
//     forms   E^{n+1}[P]    from H^{n+1/2}[P], H^{n+1/2}[P-1]      (x-1 = the thread's own registers)
//     stores  H^{n+1/2}[P], E^{n+1}[P]
// so a cell-update moves 15 words = 60 B (fp32) through the SM.  y+-1 neighbours are exchanged
// through shared memory, z+-1 through warp shuffles; the one halo column on each side of the
// tile is loaded (and its H recomputed) redundantly.  Fields are ping-ponged between two
// buffer sets by step parity, which is what makes the fused update and the halo recompute
// race-free.
//
// Temporal blocking through L2.  Stage n+1 starts one plane later than stage n
// (start plane = n mod X; this also resolves the periodic x wrap with no special case) and may
// process sweep index k as soon as stage n has finished index k+2 of its own sweep on the three
// y-tiles t-1, t, t+1.  The `stages` steps in flight therefore trail each other by ~3 planes and
// touch a window of only stages*3 planes, which stays resident in the 126 MB L2: each plane is
// fetched from HBM once per `stages` time steps instead of once per step.  Ordering is carried
// by one progress counter per CTA (st.release.gpu / ld.acquire.gpu); mutable field data is read
// with ld.global.cg because L1 is not coherent across SMs within a launch.
#pragma once

#include <stdio.h>

#include <string>

#include "fdtd_common.cuh"

namespace b200 {

constexpr int kSysMaxThreads = 512;
constexpr int kSysFlagStride = 8;   // unsigned words between progress counters (32 B sectors)

struct SystolicCfg {
  int tile_y;            // nominal columns per tile (tiles are balanced: floor/ceil of Y/ntiles)
  int ntiles;
  int stages;
  int threads;           // CTA size = roundup32((max tile columns + 2) * Zq)
  int smem_bytes;
  int max_lead;          // a stage may run at most this many planes ahead of the next stage
  union {
    int need_zfix;       // systolic / systolic_async: z-columns straddle warps (Zq does not divide 32)
    int spin_ns0;        // systolic_lean: first back-off of a waiting compute warp, ns
  };
  int trap_on_timeout;
  int pf_ahead;          // systolic_async: planes of L2 prefetch beyond the staging ring
  int svc_sleep_ns;      // systolic_async: back-off of the poller / publisher warps
  int cols;              // systolic_async: adjacent columns per compute thread (1 or 2)
  int spin_ns_max;       // systolic_lean: back-off ceiling of a waiting compute warp
  int discard;           // systolic_lean: drop consumed field lines from L2 (discard.global.L2)
  int block_order;       // systolic_lean: blockIdx -> (stage, tile): 2 = stage-major, 1 = tile-major
  long long l2_window_bytes;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Spin until *flag >= need.  Returns false on timeout (5 s) or when another CTA flagged an error.
__device__ __forceinline__ bool wait_ge(const unsigned* flag, unsigned need, unsigned* status) {
  unsigned v = ld_acquire_u32(flag);
  if (v >= need) return true;
  const unsigned long long t0 = globaltimer_ns();
  unsigned spins = 0;
  while (true) {
    v = ld_acquire_u32(flag);
    if (v >= need) return true;
    if ((++spins & 63u) == 0) {
      if (ld_acquire_u32(status) != 0) return false;
      if (globaltimer_ns() - t0 > 5000000000ull) {
        atomicCAS(status, 0u, 1u + blockIdx.x);
        return false;
      }
    }
    __nanosleep(40);
  }
}

// Progress counters [stages][ntiles], the status word, then 4 x stages slots of a y-slab session
// (kernels_lean.cuh, SlabPeers): MIRROR slots, where the neighbouring GPUs' couriers store the
// counters of their edge tiles (low neighbour's last tile, then high neighbour's first tile),
// and this GPU's COURIER progress (low side, then high side).
inline size_t systolic_sync_bytes(const SystolicCfg& c) {
  return ((size_t)c.stages * c.ntiles + 1 + 4 * (size_t)c.stages) * kSysFlagStride * sizeof(unsigned);
}

template <typename T>
__global__ void __launch_bounds__(kSysMaxThreads)
systolic_kernel(const Geom g, const Ptrs<T> p, const SystolicCfg cfg, unsigned* sync) {
  constexpr int VW = VecTraits<T>::VW;
  extern __shared__ float4 smem[];
  const int nth = blockDim.x;
  float4* const sEz = smem;            // E^n[P] halo exchange (y+1): raw storage vectors
  float4* const sEx = smem + nth;
  float4* const sHz = smem + 2 * nth;  // H^{n+1/2}[P] exchange (y-1, and z-1 across warps)
  float4* const sHx = smem + 3 * nth;
  float4* const sHy = smem + 4 * nth;
  __shared__ int s_ok;

  const int tid = threadIdx.x, lane = tid & 31;
  const int S = cfg.stages, NT = cfg.ntiles;
  const int t = blockIdx.x % NT, j = blockIdx.x / NT;
  const int y0 = (int)((long long)t * g.Y / NT);
  const int Yt = (int)((long long)(t + 1) * g.Y / NT) - y0;
  const int c = tid / g.Zq, q = tid - c * g.Zq;
  const bool active = c < Yt + 2;              // thread maps onto a loaded column
  const bool doH = c <= Yt;                    // forms H (columns y0-1 .. y0+Yt-1)
  const bool own = c >= 1 && c <= Yt;          // owns the cell: forms E, stores E and H
  const int y = wrapi(y0 - 1 + (active ? c : 0), g.Y);
  const size_t coff = ((size_t)y * g.Zq + q) * VW;
  const int slot = psi_slot(g, q);
  const bool has_psi = slot >= 0;
  const size_t poff = (((size_t)y) * g.npg + (has_psi ? slot : 0)) * VW;  // + x*Y*npg*VW
  const size_t pplane = (size_t)g.Y * g.npg * VW;
  const bool fix_up = cfg.need_zfix && lane == 31 && q + 1 < g.Zq;
  const bool fix_dn = cfg.need_zfix && lane == 0 && q > 0;
  const bool top = q + 1 == g.Zq, bottom = q == 0;

  unsigned* const status = sync + (size_t)S * NT * kSysFlagStride;
  unsigned* const my_prog = sync + ((size_t)j * NT + t) * kSysFlagStride;
  const int jp = (j + S - 1) % S, jn = (j + 1) % S;
  // warp 0, lanes 0..2 watch the previous stage's tiles t-1, t, t+1; lane 3 the next stage.
  const unsigned* watch = nullptr;
  if (tid < 3) watch = sync + ((size_t)jp * NT + wrapi(t - 1 + tid, NT)) * kSysFlagStride;
  else if (tid == 3) watch = sync + ((size_t)jn * NT + t) * kSysFlagStride;

  // per-thread constants: CPML tables of this z-group, absorber row pointers
  float ah[VW], bh[VW], ikh[VW], ae[VW], be[VW], ike[VW];
#pragma unroll
  for (int v = 0; v < VW; v += 4) {
    float4 r;
    r = __ldg(reinterpret_cast<const float4*>(p.tab + 0 * g.Zp + q * VW + v));
    ae[v] = r.x; ae[v + 1] = r.y; ae[v + 2] = r.z; ae[v + 3] = r.w;
    r = __ldg(reinterpret_cast<const float4*>(p.tab + 1 * g.Zp + q * VW + v));
    be[v] = r.x; be[v + 1] = r.y; be[v + 2] = r.z; be[v + 3] = r.w;
    r = __ldg(reinterpret_cast<const float4*>(p.tab + 2 * g.Zp + q * VW + v));
    ike[v] = r.x; ike[v + 1] = r.y; ike[v + 2] = r.z; ike[v + 3] = r.w;
    r = __ldg(reinterpret_cast<const float4*>(p.tab + 3 * g.Zp + q * VW + v));
    ah[v] = r.x; ah[v + 1] = r.y; ah[v + 2] = r.z; ah[v + 3] = r.w;
    r = __ldg(reinterpret_cast<const float4*>(p.tab + 4 * g.Zp + q * VW + v));
    bh[v] = r.x; bh[v + 1] = r.y; bh[v + 2] = r.z; bh[v + 3] = r.w;
    r = __ldg(reinterpret_cast<const float4*>(p.tab + 5 * g.Zp + q * VW + v));
    ikh[v] = r.x; ikh[v + 1] = r.y; ikh[v + 2] = r.z; ikh[v + 3] = r.w;
  }
  const size_t XY = (size_t)g.X * g.Y;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int n = j; n < g.tt; n += S) {
    const int m = n / S;                         // sweep number of this CTA row
    const unsigned base_prev = (unsigned)((j > 0 ? m : m - 1)) * (unsigned)g.X;
    const unsigned base_mine = (unsigned)m * (unsigned)g.X;
    const bool has_prev = n > 0, has_next = n + 1 < g.tt;
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
    // psiH is ping-ponged like the fields (the halo column re-reads the old value);
    // psiE is only touched by its owner and is updated in place.
    const float* const pHr0 = rb ? p.psiH2[0] : p.psiH[0];
    const float* const pHr1 = rb ? p.psiH2[1] : p.psiH[1];
    float* const pHw0 = rb ? p.psiH[0] : p.psiH2[0];
    float* const pHw1 = rb ? p.psiH[1] : p.psiH2[1];
    const int cstart = n % g.X;
    const int oi = snapshot_index(g, n);
    const float w0 = __ldg(p.wave + 2 * (size_t)n), w1 = __ldg(p.wave + 2 * (size_t)n + 1);

    // ---- wait for the first planes of the previous stage (sweep indices 0..2) ---------------
    if (tid < 32) {
      bool ok = true;
      if (tid < 3 && has_prev)
        ok = wait_ge(watch, base_prev + (unsigned)min(3, g.X), status);
      ok = __all_sync(0xffffffffu, ok);
      if (tid == 0) s_ok = ok;
    }
    __syncthreads();
    if (!s_ok) break;

    // E^n of the plane being updated (raw storage vectors), carried across iterations.
    float4 e0c = zero4, e1c = zero4, e2c = zero4;
    float exf_c = 0.f, eyf_c = 0.f;              // z+1 fix-ups (element 0 of the next z-group)
    float hyp[VW], hzp[VW];                      // H^{n+1/2}[P-1] of this thread's cells
#pragma unroll
    for (int i = 0; i < VW; ++i) { hyp[i] = 0.f; hzp[i] = 0.f; }

    // Sweep index k = -1 is the prologue: it forms H^{n+1/2}[cstart-1] (needed by the first E
    // update) without storing anything.  k = 0..X-1 are the real iterations.
    for (int k = -1; k < g.X; ++k) {
      const int P = wrapi(cstart + k, g.X), Pn = wrapi(P + 1, g.X);
      const size_t offP = (size_t)P * g.P + coff, offN = (size_t)Pn * g.P + coff;
      const bool real = k >= 0;

      if (k == -1) {                             // first plane: E^n[P] has not been loaded yet
        if (active) {
          e0c = ld16<LD_CG>(Er0 + offP);
          e2c = ld16<LD_CG>(Er2 + offP);
          if (doH) e1c = ld16<LD_CG>(Er1 + offP);
          if (fix_up) {
            exf_c = load_one<T, LD_CG>(Er0 + offP + VW);
            eyf_c = load_one<T, LD_CG>(Er1 + offP + VW);
          }
        }
      }
      // ---- loads of this iteration ------------------------------------------------------------
      float4 e0n = zero4, e1n = zero4, e2n = zero4, h0 = zero4, h1 = zero4, h2 = zero4;
      float4 bb0 = zero4, bb1 = zero4, bb2 = zero4;
      float exf_n = 0.f, eyf_n = 0.f;
      float psx[VW], psy[VW], qsx[VW], qsy[VW];
#pragma unroll
      for (int i = 0; i < VW; ++i) { psx[i] = 0.f; psy[i] = 0.f; qsx[i] = 0.f; qsy[i] = 0.f; }
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      if (active) {
        e0n = ld16<LD_CG>(Er0 + offN);
        e2n = ld16<LD_CG>(Er2 + offN);
        if (doH) {
          e1n = ld16<LD_CG>(Er1 + offN);
          h0 = ld16<LD_CG>(Hr0 + offP);
          h1 = ld16<LD_CG>(Hr1 + offP);
          h2 = ld16<LD_CG>(Hr2 + offP);
          if (has_psi) {
            const size_t po = (size_t)P * pplane + poff;
#pragma unroll
            for (int v = 0; v < VW; v += 4) {
              float4 r = __ldcg(reinterpret_cast<const float4*>(pHr0 + po + v));
              psx[v] = r.x; psx[v + 1] = r.y; psx[v + 2] = r.z; psx[v + 3] = r.w;
              r = __ldcg(reinterpret_cast<const float4*>(pHr1 + po + v));
              psy[v] = r.x; psy[v + 1] = r.y; psy[v + 2] = r.z; psy[v + 3] = r.w;
            }
          }
        }
        if (fix_up) {
          exf_n = load_one<T, LD_CG>(Er0 + offN + VW);
          eyf_n = load_one<T, LD_CG>(Er1 + offN + VW);
        }
        if (own && real) {
          bb0 = ld16<LD_NC>(p.B[0] + offP);
          bb1 = ld16<LD_NC>(p.B[1] + offP);
          bb2 = ld16<LD_NC>(p.B[2] + offP);
          const size_t xy = (size_t)P * g.Y + y;
          a0 = __ldg(p.A + xy); a1 = __ldg(p.A + XY + xy); a2 = __ldg(p.A + 2 * XY + xy);
          if (has_psi) {
            const size_t po = (size_t)P * pplane + poff;
#pragma unroll
            for (int v = 0; v < VW; v += 4) {
              float4 r = __ldcg(reinterpret_cast<const float4*>(p.psiE[0] + po + v));
              qsx[v] = r.x; qsx[v + 1] = r.y; qsx[v + 2] = r.z; qsx[v + 3] = r.w;
              r = __ldcg(reinterpret_cast<const float4*>(p.psiE[1] + po + v));
              qsy[v] = r.x; qsy[v + 1] = r.y; qsy[v + 2] = r.z; qsy[v + 3] = r.w;
            }
          }
        }
      }

      // ---- y+1 exchange of E^n[P] (Ez, Ex) -----------------------------------------------------
      sEz[tid] = e2c;
      sEx[tid] = e0c;
      __syncthreads();                                                        // (A)
      float ex[VW], ey[VW], ez[VW];
      unpack(e0c, ex, T()); unpack(e1c, ey, T()); unpack(e2c, ez, T());
      float ez_yp[VW], ex_yp[VW], ey_xp[VW], ez_xp[VW], hx[VW], hy[VW], hz[VW];
      {
        const int nb = doH ? tid + g.Zq : tid;
        unpack(sEz[nb], ez_yp, T());
        unpack(sEx[nb], ex_yp, T());
      }
      unpack(e1n, ey_xp, T()); unpack(e2n, ez_xp, T());
      unpack(h0, hx, T()); unpack(h1, hy, T()); unpack(h2, hz, T());
      // z+1 neighbours of the last lane element: element 0 of the next z-group.
      float ex_top = __shfl_down_sync(0xffffffffu, ex[0], 1);
      float ey_top = __shfl_down_sync(0xffffffffu, ey[0], 1);
      if (fix_up) { ex_top = exf_c; ey_top = eyf_c; }
      if (top) { ex_top = 0.f; ey_top = 0.f; }
#pragma unroll
      for (int i = 0; i < VW; ++i) {
        const float exz = (i + 1 < VW) ? ex[(i + 1) % VW] : ex_top;
        const float eyz = (i + 1 < VW) ? ey[(i + 1) % VW] : ey_top;
        h_cell(ex[i], ey[i], ez[i], exz, eyz, ez_yp[i], ex_yp[i], ey_xp[i], ez_xp[i], ah[i], bh[i],
               ikh[i], g.dt, psx[i], psy[i], hx[i], hy[i], hz[i]);
        hx[i] = round_store<T>(hx[i]); hy[i] = round_store<T>(hy[i]); hz[i] = round_store<T>(hz[i]);
      }

      if (real) {
        // ---- y-1 / z-1 exchange of H^{n+1/2}[P] ------------------------------------------------
        const float4 hxv = pack(hx, T()), hyv = pack(hy, T()), hzv = pack(hz, T());
        sHz[tid] = hzv;
        sHx[tid] = hxv;
        if (cfg.need_zfix) sHy[tid] = hyv;
        // warp 0 checks the dependencies of the NEXT iteration while the others compute.
        if (tid < 32) {
          bool ok = true;
          const unsigned kk = (unsigned)(k + 1);
          if (kk < (unsigned)g.X) {
            if (tid < 3 && has_prev)
              ok = wait_ge(watch, base_prev + (unsigned)min(k + 4, g.X), status);
            else if (tid == 3 && has_next && j + 1 < S && (int)kk > cfg.max_lead)
              // do not run more than max_lead planes ahead of the next stage (same sweep): keeps
              // the planes in flight inside L2.  max_lead >= 4 cannot deadlock.
              ok = wait_ge(watch, base_mine + kk - (unsigned)cfg.max_lead, status);
          }
          ok = __all_sync(0xffffffffu, ok);
          if (tid == 0) s_ok = ok;
        }
        __syncthreads();                                                      // (B)
        if (!s_ok) break;

        float hz_ym[VW], hx_ym[VW];
        {
          const int nb = own ? tid - g.Zq : tid;
          unpack(sHz[nb], hz_ym, T());
          unpack(sHx[nb], hx_ym, T());
        }
        float hx_bot = __shfl_up_sync(0xffffffffu, hx[VW - 1], 1);
        float hy_bot = __shfl_up_sync(0xffffffffu, hy[VW - 1], 1);
        if (fix_dn) {
          float tmp[VW];
          unpack(sHx[tid - 1], tmp, T()); hx_bot = tmp[VW - 1];
          unpack(sHy[tid - 1], tmp, T()); hy_bot = tmp[VW - 1];
        }
        if (bottom) { hx_bot = 0.f; hy_bot = 0.f; }
        float b0[VW], b1[VW], b2[VW];
        unpack(bb0, b0, T()); unpack(bb1, b1, T()); unpack(bb2, b2, T());
#pragma unroll
        for (int i = 0; i < VW; ++i) {
          const float hxz = (i > 0) ? hx[(i + VW - 1) % VW] : hx_bot;
          const float hyz = (i > 0) ? hy[(i + VW - 1) % VW] : hy_bot;
          e_cell(hx[i], hy[i], hz[i], hxz, hyz, hz_ym[i], hx_ym[i], hyp[i], hzp[i], ae[i], be[i],
                 ike[i], a0, a1, a2, b0[i], b1[i], b2[i], qsx[i], qsy[i], ex[i], ey[i], ez[i]);
        }
        if (own) {
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
            for (int i = 0; i < VW; ++i) {
              ex[i] = round_store<T>(ex[i]); ey[i] = round_store<T>(ey[i]);
              ez[i] = round_store<T>(ez[i]);
            }
            write_snapshot<VW>(g, p.out, oi, P, y, q, ex, ey, ez, p.proj);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < VW; ++i) { hyp[i] = hy[i]; hzp[i] = hz[i]; }
      e0c = e0n; e1c = e1n; e2c = e2n; exf_c = exf_n; eyf_c = eyf_n;

      __syncthreads();                                                        // (C)
      if (real && tid == 0) {
        __threadfence();
        st_release_u32(my_prog, base_mine + (unsigned)(k + 1));
      }
    }
    if (!s_ok) break;
  }
}

// Chooses tile/stage counts.  Returns false (with a reason) if the kernel cannot be used.
template <typename T>
bool systolic_configure(const Geom& g, int tile_y_req, int stages_req, int threads_req, int sms,
                        int l2_bytes, SystolicCfg* cfg, std::string* why) {
  const int max_threads = threads_req > 0 ? (threads_req < kSysMaxThreads ? threads_req : kSysMaxThreads)
                                          : kSysMaxThreads;
  if (g.Zq > max_threads / 3) {
    *why = "z extent too large for one CTA";
    return false;
  }
  int max_tile = max_threads / g.Zq - 2;        // columns owned, + 2 halo columns
  if (max_tile < 1) { *why = "z extent too large for one CTA"; return false; }
  if (tile_y_req > 0 && tile_y_req < max_tile) max_tile = tile_y_req;
  if (max_tile > g.Y) max_tile = g.Y;
  int ntiles = (g.Y + max_tile - 1) / max_tile;
  const int widest = (g.Y + ntiles - 1) / ntiles;
  int threads = ((widest + 2) * g.Zq + 31) / 32 * 32;
  cfg->tile_y = widest;
  cfg->ntiles = ntiles;
  cfg->threads = threads;
  cfg->need_zfix = (32 % g.Zq != 0);
  cfg->smem_bytes = 5 * threads * (int)sizeof(float4);
  cfg->max_lead = 6;
  cfg->trap_on_timeout = 1;
  int occ = 0;
  if (cudaFuncSetAttribute(systolic_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           cfg->smem_bytes) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, systolic_kernel<T>, threads,
                                                    cfg->smem_bytes) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    *why = "kernel does not fit on an SM";
    return false;
  }
  const long long capacity = (long long)occ * sms;
  if (ntiles > capacity) { *why = "more y-tiles than co-resident CTAs"; return false; }
  int stages = (int)(capacity / ntiles);
  // keep the window of planes in flight (~3 planes per stage, both buffer sets + coefficients)
  // within ~half of L2
  const long long plane_bytes = g.P * (long long)sizeof(T) * 15;
  long long by_l2 = (long long)(l2_bytes * 0.5) / (3 * plane_bytes);
  if (by_l2 < 1) by_l2 = 1;
  if (stages > by_l2) stages = (int)by_l2;
  if (stages_req > 0 && stages_req < stages) stages = stages_req;
  if (stages_req > 0 && stages_req > stages && stages_req <= capacity / ntiles) stages = stages_req;
  if (stages > g.tt) stages = g.tt > 0 ? g.tt : 1;
  if (stages > g.X && g.X >= 1) stages = g.X;   // start planes must be distinct within a sweep
  cfg->stages = stages;
  cfg->l2_window_bytes = (long long)stages * 3 * plane_bytes;
  return true;
}

// Turns a timed-out wait (a dependency that never arrived) into a loud, sticky CUDA error.
__global__ void systolic_check_kernel(const unsigned* status) {
  if (*status != 0) {
    printf("b200fdtd: systolic kernel gave up waiting on a dependency (first CTA %u)\n",
           *status - 1);
    __trap();
  }
}

template <typename T>
int systolic_launch(const Geom& g, const Ptrs<T>& p, const SystolicCfg& cfg, unsigned* sync,
                    cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(systolic_kernel<T>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, cfg.smem_bytes);
  if (e != cudaSuccess) return (int)e;
  Geom gg = g;
  Ptrs<T> pp = p;
  SystolicCfg cc = cfg;
  void* args[] = {&gg, &pp, &cc, &sync};
  e = cudaLaunchCooperativeKernel((const void*)systolic_kernel<T>, dim3(cfg.stages * cfg.ntiles),
                                  dim3(cfg.threads), args, cfg.smem_bytes, st);
  if (e != cudaSuccess) return (int)e;
  systolic_check_kernel<<<1, 1, 0, st>>>(
      sync + (size_t)cfg.stages * cfg.ntiles * kSysFlagStride);
  return (int)cudaGetLastError();
}

}  // namespace b200
