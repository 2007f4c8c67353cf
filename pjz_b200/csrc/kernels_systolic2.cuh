// Systolic kernel, asynchronous generation ("systolic_async"): same decomposition and
// dependency protocol as kernels_systolic.cuh (read its header first), re-plumbed so that
// neither memory latency nor inter-CTA synchronisation sits on the compute warps' critical path.
//
//  * Operand staging.  Every compute thread copies the 16-byte vectors of ITS OWN cells
//    (E^n[P+1], H^{n-1/2}[P], B[P], psi[P]) from L2 into a shared-memory ring with cp.async.cg
//    (LDGSTS, L1-bypassing, no registers held) D iterations before they are used.  x+1, y+1
//    and cross-warp z+1 neighbours are then plain shared-memory reads of the neighbouring
//    threads' slots: there is no exchange copy for E at all.
//      ring depth:  E needs D+2 plane slots (P and P+1 are both live); H, B and psi D+1.
//  * L2 prefetch.  The ring only hides L2 latency.  The stage that currently leads the window
//    reads planes nobody touched for a whole sweep, i.e. from HBM; the poller warp therefore
//    issues cp.async.bulk.prefetch.L2 for the tile's (contiguous) column range several planes
//    ahead of the ring, so every ring load is an L2 hit.
//  * Decoupled synchronisation.  Two service warps replace the spin/fence of the register
//    kernel:
//      - the POLLER continuously ld.acquire.gpu's the three predecessor counters (+ the
//        successor's, for the max_lead throttle) and mirrors them into shared memory;
//      - the PUBLISHER watches a shared-memory "iterations finished" word and, whenever it
//        advances, performs the st.release.gpu (fence included) of this CTA's counter.
//    Compute warps only read/write those shared-memory words (acquire/release at CTA scope),
//    so an L2 round trip or a fence never stalls them; if a service warp is slower than an
//    iteration it simply coalesces updates.
//  * Two compute-only CTA barriers per plane: (A) "ring slot landed, previous iteration's
//    shared-memory reads done", (B) exchange of the freshly formed H for the y-1/z-1 neighbours.
//
// Register budget: registers are partitioned per SM sub-partition, so at most 16 warps can hold
// more than 96 registers each; the CTA is therefore capped at 14 compute + 2 service warps.
#pragma once

#include <stdio.h>
#include <stdlib.h>

#include <string>

#include "fdtd_common.cuh"
#include "kernels_systolic.cuh"

namespace b200 {

constexpr int kSys2MaxCompute = 448;   // 14 warps
constexpr int kSys2Service = 64;       // poller warp + publisher warp

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
// Barrier over the compute warps that also AND-reduces a predicate (uniform failure decisions).
__device__ __forceinline__ bool bar_compute_and(int nthreads, bool pred) {
  int r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "bar.red.and.pred p, 1, %1, q;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(r) : "r"(nthreads), "r"((int)pred) : "memory");
  return r != 0;
}
// Shared-memory control words are accessed with volatile loads/stores: an acquire/release at
// CTA scope costs a MEMBAR.ALL.CTA per access, and none is needed -- each consumer branches on
// the loaded value before issuing its dependent memory operations, each producer's stored value
// is data-dependent on the loads it summarises.
__device__ __forceinline__ unsigned ld_vol_s(const unsigned* smem_word) {
  return *reinterpret_cast<const volatile unsigned*>(smem_word);
}
__device__ __forceinline__ void st_vol_s(unsigned* smem_word, unsigned v) {
  *reinterpret_cast<volatile unsigned*>(smem_word) = v;
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* gptr, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

// Shared control block of one CTA.
struct Sys2Ctl {
  unsigned avail;     // min over the three predecessor counters (raw, cumulative)
  unsigned next;      // successor stage's counter on this tile
  unsigned done;      // this CTA's cumulative count of finished sweep indices
  unsigned front;     // cumulative iteration index the compute warps have reached (for prefetch)
  unsigned ok;        // 0 once any CTA gave up
  unsigned exit_;     // compute warps are finished
};

template <typename T, int D>
__global__ void __launch_bounds__(kSys2MaxCompute + kSys2Service)
systolic2_kernel(const Geom g, const Ptrs<T> p, const SystolicCfg cfg, unsigned* sync) {
  constexpr int VW = VecTraits<T>::VW;
  constexpr int NE = D + 2, NH = D + 1;
  constexpr int PV = VW / 4;                     // float4 per psi vector
  extern __shared__ float4 smem[];
  __shared__ Sys2Ctl ctl;
  const int NTc = blockDim.x - kSys2Service;     // compute threads
  const int tid = threadIdx.x, lane = tid & 31;
  const int S = cfg.stages, NT = cfg.ntiles;
  const int t = blockIdx.x % NT, j = blockIdx.x / NT;
  const int y0 = (int)((long long)t * g.Y / NT);
  const int Yt = (int)((long long)(t + 1) * g.Y / NT) - y0;
  const int npsi = (cfg.tile_y + 2) * g.npg * PV; // float4 per psi array per slot

  float4* const sE = smem;                                   // [NE][3][NTc]
  float4* const sH = sE + (size_t)NE * 3 * NTc;              // [NH][3][NTc]
  float4* const sB = sH + (size_t)NH * 3 * NTc;              // [NH][3][NTc]
  float4* const sX = sB + (size_t)NH * 3 * NTc;              // [3][NTc]  Hz, Hx, Hy (new)
  float4* const sP = sX + (size_t)3 * NTc;                   // [NH][4][npsi]
  float4* const sT = sP + (size_t)NH * 4 * npsi;             // [6][Zp/4] CPML tables

  unsigned* const status = sync + (size_t)S * NT * kSysFlagStride;
  unsigned* const my_prog = sync + ((size_t)j * NT + t) * kSysFlagStride;

  if (tid == 0) {
    ctl.avail = 0; ctl.next = 0; ctl.done = 0; ctl.front = 0; ctl.ok = 1; ctl.exit_ = 0;
  }
  for (int i = tid; i < 6 * g.Zp / 4; i += blockDim.x)
    sT[i] = __ldg(reinterpret_cast<const float4*>(p.tab) + i);
  __syncthreads();

  // =================================== poller warp ===============================================
  if (tid >= NTc && tid < NTc + 32) {
    const int jp = (j + S - 1) % S, jn = (j + 1) % S;
    const unsigned* watch = sync + ((size_t)jp * NT + wrapi(t - 1 + (lane < 3 ? lane : 1), NT)) *
                                       kSysFlagStride;
    if (lane == 3) watch = sync + ((size_t)jn * NT + t) * kSysFlagStride;
    if (lane == 4) watch = status;
    // L2 prefetch duty: lanes 8..16 own one array each (E0,E1,E2,H0,H1,H2 of the read set, B0..2).
    const int ylo = max(y0 - 1, 0), yhi = min(y0 + Yt, g.Y - 1);
    const unsigned pf_bytes = (unsigned)((yhi - ylo + 1) * g.Zp * (int)sizeof(T));
    const size_t pf_off = (size_t)ylo * g.Zp;
    unsigned pf_done = 0;                          // cumulative iterations already prefetched
    const unsigned sweep_iters = (unsigned)g.X + 1u;
    while (ld_vol_s(&ctl.exit_) == 0) {
      unsigned v = 0xffffffffu;
      if (lane < 5) v = ld_acquire_u32(watch);
      const unsigned v0 = __shfl_sync(0xffffffffu, v, 0), v1 = __shfl_sync(0xffffffffu, v, 1),
                     v2 = __shfl_sync(0xffffffffu, v, 2), v3 = __shfl_sync(0xffffffffu, v, 3),
                     v4 = __shfl_sync(0xffffffffu, v, 4);
      if (lane == 0) {
        st_vol_s(&ctl.avail, min(v0, min(v1, v2)));
        st_vol_s(&ctl.next, v3);
        if (v4 != 0) st_vol_s(&ctl.ok, 0u);
      }
      // prefetch the planes of iterations [front + D, front + D + pf_ahead) into L2
      const unsigned front = ld_vol_s(&ctl.front);
      const unsigned want = front + (unsigned)D + (unsigned)cfg.pf_ahead;
      if (cfg.pf_ahead > 0 && lane >= 8 && lane < 17) {
        if (pf_done < front + (unsigned)D) pf_done = front + (unsigned)D;
        for (; pf_done < want; ++pf_done) {
          const unsigned sweep = pf_done / sweep_iters, it = pf_done % sweep_iters;
          const int n = j + (int)sweep * S;
          if (n >= g.tt) break;
          const int rb = n & 1;
          const int P = wrapi(n % g.X - 1 + (int)it, g.X), Pn = wrapi(P + 1, g.X);
          const int a = lane - 8;
          const T* base;
          int plane;
          if (a < 3) { base = rb ? p.E2[a] : p.E[a]; plane = Pn; }
          else if (a < 6) { base = rb ? p.H2[a - 3] : p.H[a - 3]; plane = P; }
          else { base = p.B[a - 6]; plane = P; }
          prefetch_l2_bulk(base + (size_t)plane * g.P + pf_off, pf_bytes);
        }
      }
      pf_done = __shfl_sync(0xffffffffu, pf_done, 8);
      __nanosleep(100);
    }
    return;
  }
  // ================================== publisher warp =============================================
  if (tid >= NTc + 32) {
    if (lane == 0) {
      unsigned last = 0;
      while (true) {
        const unsigned ex = ld_vol_s(&ctl.exit_);
        const unsigned d = ld_vol_s(&ctl.done);
        if (d != last) {
          st_release_u32(my_prog, d);              // fence.acq_rel.gpu + store
          last = d;
        } else if (ex) {
          break;
        } else {
          __nanosleep(50);
        }
      }
    }
    return;
  }

  // ================================= compute warps ===============================================
  // Everything that does not change along the sweep is computed here, once: the loop body below
  // is issue-bound, so it carries running plane/slot indices and forms each global offset with
  // a single 64-bit multiply-add per plane.
  const int Zq = g.Zq, X = g.X;
  const int c = tid / Zq, q = tid - c * Zq;
  const bool active = c < Yt + 2;
  const bool doH = c <= Yt;
  const bool own = c >= 1 && c <= Yt;
  const int y = wrapi(y0 - 1 + (active ? c : 0), g.Y);
  const size_t coff = ((size_t)y * Zq + q) * VW;
  const int slot = psi_slot(g, q);
  const bool has_psi = slot >= 0;
  const bool psiH_thr = has_psi && doH, psiE_thr = has_psi && own;
  const size_t poff = ((size_t)y * g.npg + (has_psi ? slot : 0)) * VW;
  const size_t pplane = (size_t)g.Y * g.npg * VW;
  const size_t gP = (size_t)g.P;
  const int pidx = (c * g.npg + (has_psi ? slot : 0)) * PV;   // float4 index inside a psi slot
  const bool fix_up = cfg.need_zfix && lane == 31 && q + 1 < Zq;
  const bool fix_dn = cfg.need_zfix && lane == 0 && q > 0;
  const bool top = q + 1 == Zq, bottom = q == 0;
  const size_t XY = (size_t)X * g.Y;
  const int nbp = doH ? tid + Zq : tid;            // ring index of the y+1 neighbour
  const int nbm = own ? tid - Zq : tid;            // ring index of the y-1 neighbour
  const int eslot = 3 * NTc, pslot = 4 * npsi;     // float4 per ring slot
  const int tstride = g.Zp / 4;
  const float4* const tq = sT + q * PV;            // CPML table w of this z-group: tq[w*tstride + v]
  // plane source: cheap pre-test so that add_source() is off the common path
  const int sp0 = g.src_pos, sp1 = wrapi(g.src_pos - 1, g.src_axis == 0 ? X : g.Y);
  const bool src_thr = g.src_axis == 1 ? (y == sp0 || y == sp1)
                                       : (g.src_axis == 2 && q == g.src_pos / VW);
  bool ok = true;

  auto load_tab = [&](int which, float (&dst)[VW]) {
#pragma unroll
    for (int v = 0; v < PV; ++v) {
      const float4 r = tq[which * tstride + v];
      dst[4 * v] = r.x; dst[4 * v + 1] = r.y; dst[4 * v + 2] = r.z; dst[4 * v + 3] = r.w;
    }
  };

  unsigned iters_done = 0;                         // cumulative iterations finished (for front)
  for (int n = j; n < g.tt && ok; n += S) {
    const int m = n / S;
    const unsigned base_prev = (unsigned)((j > 0 ? m : m - 1)) * (unsigned)X;
    const unsigned base_mine = (unsigned)m * (unsigned)X;
    const bool has_prev = n > 0, has_next = n + 1 < g.tt && j + 1 < S;
    const int rb = n & 1, wb = rb ^ 1;
    const T* const Er0 = p.Es[rb][0]; const T* const Er1 = p.Es[rb][1]; const T* const Er2 = p.Es[rb][2];
    const T* const Hr0 = p.Hs[rb][0]; const T* const Hr1 = p.Hs[rb][1]; const T* const Hr2 = p.Hs[rb][2];
    const float* const pHr0 = p.psiHs[rb][0]; const float* const pHr1 = p.psiHs[rb][1];
    const int cstart = n % X;
    const int oi = snapshot_index(g, n);
    const float w0 = __ldg(p.wave + 2 * (size_t)n), w1 = __ldg(p.wave + 2 * (size_t)n + 1);

    // Blocks until the loads of iteration `it` are allowed: they read planes up to sweep index
    // it+1 of the previous stage, which therefore must have finished it+2 indices on the tiles
    // t-1, t, t+1 (the k+3 rule); and they must not run more than max_lead indices ahead of the
    // next stage (keeps the planes in flight inside L2).  Spins on shared memory only; the
    // branch on the loaded value orders the following loads behind it (no speculation on GPUs).
    auto wait_deps = [&](int it) -> bool {
      const unsigned need = base_prev + (unsigned)min(it + 2, X);
      const int lead = min(it, X) - 1 - cfg.max_lead;
      const unsigned need_next = base_mine + (unsigned)max(lead, 0);
      const bool chk_a = has_prev, chk_b = has_next && lead > 0;
      if ((!chk_a || ld_vol_s(&ctl.avail) >= need) && (!chk_b || ld_vol_s(&ctl.next) >= need_next))
        return true;
      unsigned long long t0 = 0;
      unsigned spins = 0;
      while (true) {
        const bool a = !chk_a || ld_vol_s(&ctl.avail) >= need;
        const bool b = !chk_b || ld_vol_s(&ctl.next) >= need_next;
        if (a && b) return true;
        if (ld_vol_s(&ctl.ok) == 0) return false;
        if ((++spins & 255u) == 0) {
          const unsigned long long now = globaltimer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 5000000000ull) {
            atomicCAS(status, 0u, 1u + blockIdx.x);
            st_vol_s(&ctl.ok, 0u);
            return false;
          }
        }
      }
    };

    // Async copies consumed by one iteration: E[Pn] -> E slot `se`, H/B/psi[P] -> slot `sh`.
    auto issue = [&](int P, int Pn, int se, int sh, bool first, bool ecoef) {
      if (active) {
        const size_t offP = (size_t)P * gP + coff, offN = (size_t)Pn * gP + coff;
        float4* e = sE + se * eslot + tid;
        cp_async16(e, Er0 + offN);
        cp_async16(e + 2 * NTc, Er2 + offN);
        if (doH) cp_async16(e + NTc, Er1 + offN);
        if (first) {                               // very first plane of the sweep: E[P] too
          float4* e0 = sE + tid;
          cp_async16(e0, Er0 + offP);
          cp_async16(e0 + 2 * NTc, Er2 + offP);
          if (doH) cp_async16(e0 + NTc, Er1 + offP);
        }
        if (doH) {
          float4* h = sH + sh * eslot + tid;
          cp_async16(h, Hr0 + offP);
          cp_async16(h + NTc, Hr1 + offP);
          cp_async16(h + 2 * NTc, Hr2 + offP);
          if (own && ecoef) {
            float4* b = sB + sh * eslot + tid;
            cp_async16(b, p.B[0] + offP);
            cp_async16(b + NTc, p.B[1] + offP);
            cp_async16(b + 2 * NTc, p.B[2] + offP);
          }
          if (has_psi) {
            float4* ps = sP + sh * pslot + pidx;
            const size_t po = (size_t)P * pplane + poff;
#pragma unroll
            for (int v = 0; v < PV; ++v) {
              cp_async16(ps + v, pHr0 + po + 4 * v);
              cp_async16(ps + npsi + v, pHr1 + po + 4 * v);
              if (own && ecoef) {
                cp_async16(ps + 2 * npsi + v, p.psiE[0] + po + 4 * v);
                cp_async16(ps + 3 * npsi + v, p.psiE[1] + po + 4 * v);
              }
            }
          }
        }
      }
    };

    // ---- fill the ring: iterations 0 .. D-1 ------------------------------------------------------
    // A failed wait (timeout / another CTA gave up) becomes a CTA-uniform decision at the next
    // AND-reducing barrier, so all compute threads leave the loops at the same point.
    int PL = wrapi(cstart - 1, X);                 // plane whose H/B/psi the next issue() loads
#pragma unroll
    for (int it = 0; it < D; ++it) {
      ok = ok && wait_deps(it);
      const int PLn = PL + 1 == X ? 0 : PL + 1;
      if (ok && it <= X) issue(PL, PLn, it + 1, it, it == 0, it >= 1);
      cp_async_commit();
      PL = PLn;
    }
    ok = bar_compute_and(NTc, ok);

    float hyp[VW], hzp[VW];
#pragma unroll
    for (int v = 0; v < VW; ++v) { hyp[v] = 0.f; hzp[v] = 0.f; }
    float a0n = 0.f, a1n = 0.f, a2n = 0.f;       // absorber row of the NEXT plane (prefetched)
    int P = wrapi(cstart - 1, X);                  // plane of iteration i (i = 0: prologue plane)
    int se = 0, sh = 0;                            // ring slots of E[P] and H/B/psi[P]

    for (int i = 0; i <= X && ok; ++i) {
      const bool real = i >= 1;
      const int Pn = P + 1 == X ? 0 : P + 1;
      const int sen = se + 1 == NE ? 0 : se + 1;   // slot of E[P+1]
      cp_async_wait<D - 1>();
      bar_compute(NTc);                            // A_i
      if (tid == 0) {
        // every store of iterations < i has been issued by all compute threads
        if (i >= 2) st_vol_s(&ctl.done, base_mine + (unsigned)(i - 1));
        st_vol_s(&ctl.front, iters_done + (unsigned)i);
      }
      ok = wait_deps(i + D);
      {
        const int PLn = PL + 1 == X ? 0 : PL + 1;
        // loads of iteration i+D go to the slots freed by iteration i-1
        if (ok && i + D <= X)
          issue(PL, PLn, se == 0 ? NE - 1 : se - 1, sh == 0 ? NH - 1 : sh - 1, false, true);
        cp_async_commit();
        PL = PLn;
      }

      const float a0 = a0n, a1 = a1n, a2 = a2n;
      if (own && i < X) {
        const size_t xy = (size_t)Pn * g.Y + y;
        a0n = __ldg(p.A + xy); a1n = __ldg(p.A + XY + xy); a2n = __ldg(p.A + 2 * XY + xy);
      }
      const float4* eC = sE + se * eslot;          // E^n[P]
      const float4* eN = sE + sen * eslot;         // E^n[P+1]
      const float4* hO = sH + sh * eslot;          // H^{n-1/2}[P]
      const float4* bC = sB + sh * eslot;          // B[P]
      const float4* pS = sP + sh * pslot + pidx;

      float ex[VW], ey[VW], ez[VW], hx[VW], hy[VW], hz[VW];
      {
        float ez_yp[VW], ex_yp[VW], ey_xp[VW], ez_xp[VW], psx[VW], psy[VW];
        unpack(eC[tid], ex, T()); unpack(eC[NTc + tid], ey, T()); unpack(eC[2 * NTc + tid], ez, T());
        unpack(eC[2 * NTc + nbp], ez_yp, T()); unpack(eC[nbp], ex_yp, T());
        unpack(eN[NTc + tid], ey_xp, T()); unpack(eN[2 * NTc + tid], ez_xp, T());
        unpack(hO[tid], hx, T()); unpack(hO[NTc + tid], hy, T()); unpack(hO[2 * NTc + tid], hz, T());
#pragma unroll
        for (int v = 0; v < VW; ++v) { psx[v] = 0.f; psy[v] = 0.f; }
        if (psiH_thr) {
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
        if (psiH_thr && own && real) {             // new psiH of the owned PML cells
          const size_t po = (size_t)P * pplane + poff;
          float* const w0p = p.psiHs[wb][0] + po;
          float* const w1p = p.psiHs[wb][1] + po;
#pragma unroll
          for (int v = 0; v < VW; v += 4) {
            __stcg(reinterpret_cast<float4*>(w0p + v), make_float4(psx[v], psx[v + 1], psx[v + 2], psx[v + 3]));
            __stcg(reinterpret_cast<float4*>(w1p + v), make_float4(psy[v], psy[v + 1], psy[v + 2], psy[v + 3]));
          }
        }
      }

      const float4 hxv = pack(hx, T()), hyv = pack(hy, T()), hzv = pack(hz, T());
      sX[tid] = hzv;
      sX[NTc + tid] = hxv;
      if (cfg.need_zfix) sX[2 * NTc + tid] = hyv;
      ok = bar_compute_and(NTc, ok);               // B_i (+ uniform failure decision)
      if (!ok) break;
      if (real) {
        float hx_bot = __shfl_up_sync(0xffffffffu, hx[VW - 1], 1);
        float hy_bot = __shfl_up_sync(0xffffffffu, hy[VW - 1], 1);
        if (fix_dn) {
          float tmp[VW];
          unpack(sX[NTc + tid - 1], tmp, T()); hx_bot = tmp[VW - 1];
          unpack(sX[2 * NTc + tid - 1], tmp, T()); hy_bot = tmp[VW - 1];
        }
        if (bottom) { hx_bot = 0.f; hy_bot = 0.f; }
        if (own) {
          const size_t offP = (size_t)P * gP + coff;
          float hz_ym[VW], hx_ym[VW], qsx[VW], qsy[VW], b0[VW], b1[VW], b2[VW];
          unpack(sX[nbm], hz_ym, T());
          unpack(sX[NTc + nbm], hx_ym, T());
#pragma unroll
          for (int v = 0; v < VW; ++v) { qsx[v] = 0.f; qsy[v] = 0.f; }
          if (psiE_thr) {
#pragma unroll
            for (int v = 0; v < PV; ++v) {
              float4 r = pS[2 * npsi + v];
              qsx[4 * v] = r.x; qsx[4 * v + 1] = r.y; qsx[4 * v + 2] = r.z; qsx[4 * v + 3] = r.w;
              r = pS[3 * npsi + v];
              qsy[4 * v] = r.x; qsy[4 * v + 1] = r.y; qsy[4 * v + 2] = r.z; qsy[4 * v + 3] = r.w;
            }
          }
          unpack(bC[tid], b0, T()); unpack(bC[NTc + tid], b1, T()); unpack(bC[2 * NTc + tid], b2, T());
          float ae[VW], be[VW], ike[VW];
          load_tab(0, ae); load_tab(1, be); load_tab(2, ike);
#pragma unroll
          for (int v = 0; v < VW; ++v) {
            const float hxz = (v > 0) ? hx[(v + VW - 1) % VW] : hx_bot;
            const float hyz = (v > 0) ? hy[(v + VW - 1) % VW] : hy_bot;
            e_cell(hx[v], hy[v], hz[v], hxz, hyz, hz_ym[v], hx_ym[v], hyp[v], hzp[v], ae[v], be[v],
                   ike[v], a0, a1, a2, b0[v], b1[v], b2[v], qsx[v], qsy[v], ex[v], ey[v], ez[v]);
          }
          if (g.src_axis == 0 ? (P == sp0 || P == sp1) : src_thr)
            add_source<VW>(g, p.src, w0, w1, P, y, q, ex, ey, ez);
          st16<LD_CG>(p.Hs[wb][0] + offP, hxv);
          st16<LD_CG>(p.Hs[wb][1] + offP, hyv);
          st16<LD_CG>(p.Hs[wb][2] + offP, hzv);
          store_vec<T, LD_CG>(p.Es[wb][0] + offP, ex);
          store_vec<T, LD_CG>(p.Es[wb][1] + offP, ey);
          store_vec<T, LD_CG>(p.Es[wb][2] + offP, ez);
          if (psiE_thr) {
            const size_t po = (size_t)P * pplane + poff;
#pragma unroll
            for (int v = 0; v < VW; v += 4) {
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
      P = Pn;
      se = sen;
      sh = sh + 1 == NH ? 0 : sh + 1;
    }
    cp_async_wait<0>();
    bar_compute(NTc);                              // end of sweep: all stores issued, ring drained
    iters_done += (unsigned)X + 1u;
    if (tid == 0 && ok) st_vol_s(&ctl.done, base_mine + (unsigned)X);
  }
  bar_compute(NTc);
  if (tid == 0) st_vol_s(&ctl.exit_, 1u);
}

template <typename T, int D>
size_t systolic2_smem_bytes(const Geom& g, int compute_threads, int tile_y) {
  constexpr int PV = VecTraits<T>::VW / 4;
  const size_t npsi = (size_t)(tile_y + 2) * g.npg * PV;
  return sizeof(float4) * ((size_t)((D + 2) * 3 + 2 * (D + 1) * 3 + 3) * compute_threads +
                           (size_t)(D + 1) * 4 * npsi + 6 * (size_t)g.Zp / 4);
}

template <typename T, int D>
bool systolic2_configure_d(const Geom& g, int tile_y_req, int stages_req, int threads_req, int sms,
                           int l2_bytes, SystolicCfg* cfg, std::string* why) {
  const int max_threads = threads_req > 0 ? (threads_req < kSys2MaxCompute ? threads_req : kSys2MaxCompute)
                                          : kSys2MaxCompute;
  if (g.Zq * 3 > max_threads) { *why = "z extent too large for one CTA"; return false; }
  int max_tile = max_threads / g.Zq - 2;
  // largest tile whose staging ring fits in shared memory
  while (max_tile >= 1 &&
         systolic2_smem_bytes<T, D>(g, ((max_tile + 2) * g.Zq + 31) / 32 * 32, max_tile) + 64 >
             227 * 1024)
    --max_tile;
  if (max_tile < 1) { *why = "staging ring does not fit in shared memory"; return false; }
  if (tile_y_req > 0 && tile_y_req < max_tile) max_tile = tile_y_req;
  if (max_tile > g.Y) max_tile = g.Y;
  const int ntiles = (g.Y + max_tile - 1) / max_tile;
  const int widest = (g.Y + ntiles - 1) / ntiles;
  const int compute = ((widest + 2) * g.Zq + 31) / 32 * 32;
  cfg->tile_y = widest;
  cfg->ntiles = ntiles;
  cfg->threads = compute + kSys2Service;
  cfg->need_zfix = (32 % g.Zq != 0);
  cfg->smem_bytes = (int)systolic2_smem_bytes<T, D>(g, compute, widest);
  // >= 2D+4 is needed for deadlock freedom (a throttled stage has published i-2 indices while its
  // successor needs i'+D+2 of them to advance); the rest is slack for the
  // polling/publishing latency.
  cfg->max_lead = 2 * D + 8;
  cfg->pf_ahead = 6;
  // tuning knobs (benchmark sweeps only; values below the deadlock bound are clamped)
  if (const char* e = getenv("B200FDTD_MAX_LEAD")) cfg->max_lead = atoi(e);
  if (const char* e = getenv("B200FDTD_PF_AHEAD")) cfg->pf_ahead = atoi(e);
  if (cfg->max_lead < 2 * D + 4) cfg->max_lead = 2 * D + 4;
  if (cfg->pf_ahead < 0) cfg->pf_ahead = 0;
  cfg->trap_on_timeout = 1;
  int occ = 0;
  if (cudaFuncSetAttribute(systolic2_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           cfg->smem_bytes) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, systolic2_kernel<T, D>, cfg->threads,
                                                    cfg->smem_bytes) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    *why = "kernel does not fit on an SM";
    return false;
  }
  const long long capacity = (long long)occ * sms;
  if (ntiles > capacity) { *why = "more y-tiles than co-resident CTAs"; return false; }
  int stages = (int)(capacity / ntiles);
  const long long plane_bytes = g.P * (long long)sizeof(T) * 15;
  const int lag = D + 5;                         // planes a stage trails its predecessor by
  long long by_l2 = (long long)(l2_bytes * 0.6) / (lag * plane_bytes);
  if (by_l2 < 1) by_l2 = 1;
  if (stages > by_l2) stages = (int)by_l2;
  if (stages_req > 0 && stages_req <= capacity / ntiles) stages = stages_req;
  if (stages > g.tt) stages = g.tt > 0 ? g.tt : 1;
  if (stages > g.X) stages = g.X;
  cfg->stages = stages;
  cfg->l2_window_bytes = (long long)stages * lag * plane_bytes;
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
