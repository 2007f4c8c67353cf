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
//  * Two adjacent columns per thread: the y+1 neighbour of the even column and the y-1
//    neighbour of the odd one are the thread's own registers, CPML tables / dependency checks /
//    loop bookkeeping are shared, and only 8 compute warps meet at the barriers.
//
// Register budget: 8 compute + 2 service warps leave ~200 registers per thread.
#pragma once

#include <stdio.h>
#include <stdlib.h>

#include <string>

#include "fdtd_common.cuh"
#include "kernels_systolic.cuh"

namespace b200 {

// Registers are partitioned per SM sub-partition (16 Ki each): with 8 warps per CTA every
// sub-partition holds 2 warps and a thread may use up to 255 registers; a 9th/10th warp would
// cap it at 168 and spill (measured: local-memory traffic halves the throughput).
constexpr int kSys2MaxCompute = 224;   // 7 warps, two columns per thread
constexpr int kSys2Service = 32;       // one service warp: poller + publisher + L2 prefetcher

// No "memory" clobber on the copy/commit: the slot being filled is not read before the
// wait_group + barrier of a LATER iteration (both are compiler barriers), and without the clobber
// ptxas is free to interleave this iteration's shared-memory reads with the copy issue.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;");
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
// Progress counters are polled with a relaxed gpu-scope load (served by L2, the point of
// coherence).  ld.acquire.gpu would add a CCTL.IVALL (L1 invalidate) per poll; the data loads
// that depend on the counter bypass L1 (cp.async.cg) and are issued behind a branch on its value.
__device__ __forceinline__ unsigned ld_relaxed_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
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

// K = adjacent columns per compute thread: 2 (7 compute warps, ~250 registers, fewest
// instructions per cell) or 1 (14 compute warps, <= 128 registers, twice the warps to hide latency).
template <typename T, int D, int K>
__global__ void __launch_bounds__(K == 2 ? kSys2MaxCompute + kSys2Service : 2 * kSys2MaxCompute + kSys2Service)
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
  const int ring = (cfg.tile_y + 2) * g.Zq;       // float4 per component per ring slot
  const int eslot = 3 * ring, pslot = 4 * npsi;   // float4 per E/H/B slot, per psi slot

  float4* const sE = smem;                                   // [NE][3][ring]
  float4* const sH = sE + (size_t)NE * eslot;                // [NH][3][ring]
  float4* const sB = sH + (size_t)NH * eslot;                // [NH][3][ring]
  float4* const sX = sB + (size_t)NH * eslot;                // [2 | 5][NTc] new H at pair boundaries
  float4* const sP = sX + (size_t)(cfg.need_zfix ? 1 + 2 * K : 2) * NTc;   // [NH][4][npsi]
  float4* const sT = sP + (size_t)NH * pslot;                // [6][Zp/4] CPML tables

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
    unsigned published = 0;
    while (true) {
      // publisher duty (lane 0): st.release.gpu = fence + store, whenever `done` advanced
      const unsigned ex = ld_vol_s(&ctl.exit_);
      const unsigned dn = ld_vol_s(&ctl.done);
      if (dn != published) {
        if (lane == 0) st_release_u32(my_prog, dn);
        published = dn;
      } else if (ex) {
        break;
      }
      unsigned v = 0xffffffffu;
      if (lane < 5) v = ld_relaxed_gpu_u32(watch);
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
          // only planes the previous step has already produced (prefetching a plane that is
          // about to be overwritten fetches dead data from HBM; measured in kernels_lean.cuh)
          if (n > 0) {
            const unsigned m = (unsigned)(n / S);
            const unsigned need = (j > 0 ? m : m - 1u) * (unsigned)g.X +
                                  (unsigned)min((int)it + 2, g.X);
            if (min(v0, min(v1, v2)) < need) break;
          }
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
      __nanosleep(cfg.svc_sleep_ns);
    }
    return;
  }

  // ================================= compute warps ===============================================
  // One thread owns the 16-byte z-vector q of TWO adjacent columns (2cp, 2cp+1) of the loaded
  // tile (columns 0 .. Yt+1 = y0-1 .. y0+Yt).  The y+1 neighbour of the even column and the y-1
  // neighbour of the odd column are the thread's own registers; only the pair boundary goes
  // through shared memory.  Everything that does not change along the sweep is computed here,
  // once: the loop body is issue-bound.
  const int Zq = g.Zq, X = g.X;
  const int cp = tid / Zq, q = tid - cp * Zq;
  const int ncols = Yt + 2;
  int f[K], yk[K];
  bool act[K], doH[K], own[K];
  unsigned coff[K];                                // element offset inside a plane (< 2^31)
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int c = K * cp + k;
    act[k] = c < ncols;
    doH[k] = c <= Yt;
    own[k] = c >= 1 && c <= Yt;
    yk[k] = wrapi(y0 - 1 + (act[k] ? c : 0), g.Y);
    f[k] = act[k] ? c * Zq + q : q;                // ring index of the item (in range if idle)
    coff[k] = (unsigned)((yk[k] * Zq + q) * VW);
  }
  const int slot = psi_slot(g, q);
  const bool has_psi = slot >= 0;
  const size_t pplane = (size_t)g.Y * g.npg * VW;
  const size_t gP = (size_t)g.P;
  unsigned ppoff[K];                               // psi offset of the item inside a plane (< 2^31)
#pragma unroll
  for (int k = 0; k < K; ++k)
    ppoff[k] = (unsigned)((yk[k] * g.npg + (has_psi ? slot : 0)) * VW);
  const bool fix_up = cfg.need_zfix && lane == 31 && q + 1 < Zq;
  const bool fix_dn = cfg.need_zfix && lane == 0 && q > 0;
  const bool top = q + 1 == Zq, bottom = q == 0;
  const size_t XY = (size_t)X * g.Y;
  const int tstride = g.Zp / 4;
  const float4* const tq = sT + q * PV;            // CPML table w of this z-group: tq[w*tstride + v]
  // plane source: cheap pre-test so that add_source() is off the common path
  const int sp0 = g.src_pos, sp1 = wrapi(g.src_pos - 1, g.src_axis == 0 ? X : g.Y);
  bool src_thr[K];
#pragma unroll
  for (int k = 0; k < K; ++k)
    src_thr[k] = g.src_axis == 1 ? (yk[k] == sp0 || yk[k] == sp1)
                                 : (g.src_axis == 2 && q == g.src_pos / VW);
  bool ok = true;

  auto load_tab = [&](int which, float (&dst)[VW]) {
#pragma unroll
    for (int v = 0; v < PV; ++v) {
      const float4 r = tq[which * tstride + v];
      dst[4 * v] = r.x; dst[4 * v + 1] = r.y; dst[4 * v + 2] = r.z; dst[4 * v + 3] = r.w;
    }
  };
  auto load_psi = [&](const float4* ps, float (&dst)[VW]) {
#pragma unroll
    for (int v = 0; v < PV; ++v) {
      const float4 r = ps[v];
      dst[4 * v] = r.x; dst[4 * v + 1] = r.y; dst[4 * v + 2] = r.z; dst[4 * v + 3] = r.w;
    }
  };
  auto store_psi = [&](float* dst, const float (&src)[VW]) {
#pragma unroll
    for (int v = 0; v < VW; v += 4)
      __stcg(reinterpret_cast<float4*>(dst + v), make_float4(src[v], src[v + 1], src[v + 2], src[v + 3]));
  };

  unsigned iters_done = 0;                         // cumulative iterations finished (for front)
  for (int n = j; n < g.tt && ok; n += S) {
    const int m = n / S;
    const unsigned base_prev = (unsigned)((j > 0 ? m : m - 1)) * (unsigned)X;
    const unsigned base_mine = (unsigned)m * (unsigned)X;
    const bool has_prev = n > 0, has_next = n + 1 < g.tt && j + 1 < S;
    const int rb = n & 1, wb = rb ^ 1;
    const int cstart = n % X;
    const int oi = snapshot_index(g, n);
    const float w0 = __ldg(p.wave + 2 * (size_t)n), w1 = __ldg(p.wave + 2 * (size_t)n + 1);

    // Blocks until the loads of iteration `it` are allowed: they read planes up to sweep index
    // it+1 of the previous stage, which therefore must have finished it+2 indices on the tiles
    // t-1, t, t+1 (the k+3 rule); and they must not run more than max_lead indices ahead of the
    // next stage (keeps the planes in flight inside L2).  Spins on shared memory only; the
    // branch on the loaded value orders the following loads behind it (no speculation on GPUs).
    auto wait_deps = [&](int it) -> bool {
      const unsigned need = has_prev ? base_prev + (unsigned)min(it + 2, X) : 0u;
      const int lead = min(it, X) - 1 - cfg.max_lead;
      const unsigned need_next = (has_next && lead > 0) ? base_mine + (unsigned)lead : 0u;
      if (ld_vol_s(&ctl.avail) >= need && ld_vol_s(&ctl.next) >= need_next) return true;
      unsigned long long t0 = 0;
      unsigned spins = 0;
      while (true) {
        if (ld_vol_s(&ctl.avail) >= need && ld_vol_s(&ctl.next) >= need_next) return true;
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
      const size_t pP = (size_t)P * gP, pN = (size_t)Pn * gP;
      float4* const eb = sE + se * eslot;
      float4* const hb = sH + sh * eslot;
      float4* const bb = sB + sh * eslot;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (act[k]) {
        const size_t offP = pP + coff[k], offN = pN + coff[k];
        cp_async16(eb + f[k], p.Es[rb][0] + offN);
        cp_async16(eb + 2 * ring + f[k], p.Es[rb][2] + offN);
        if (doH[k]) cp_async16(eb + ring + f[k], p.Es[rb][1] + offN);
        if (first) {                               // very first plane of the sweep: E[P] too
          cp_async16(sE + f[k], p.Es[rb][0] + offP);
          cp_async16(sE + 2 * ring + f[k], p.Es[rb][2] + offP);
          if (doH[k]) cp_async16(sE + ring + f[k], p.Es[rb][1] + offP);
        }
        if (doH[k]) {
          cp_async16(hb + f[k], p.Hs[rb][0] + offP);
          cp_async16(hb + ring + f[k], p.Hs[rb][1] + offP);
          cp_async16(hb + 2 * ring + f[k], p.Hs[rb][2] + offP);
          if (own[k] && ecoef) {
            cp_async16(bb + f[k], p.B[0] + offP);
            cp_async16(bb + ring + f[k], p.B[1] + offP);
            cp_async16(bb + 2 * ring + f[k], p.B[2] + offP);
          }
          if (has_psi) {
            float4* ps = sP + sh * pslot + ((K * cp + k) * g.npg + slot) * PV;
            const size_t po = (size_t)P * pplane + ppoff[k];
#pragma unroll
            for (int v = 0; v < PV; ++v) {
              cp_async16(ps + v, p.psiHs[rb][0] + po + 4 * v);
              cp_async16(ps + npsi + v, p.psiHs[rb][1] + po + 4 * v);
              if (own[k] && ecoef) {
                cp_async16(ps + 2 * npsi + v, p.psiE[0] + po + 4 * v);
                cp_async16(ps + 3 * npsi + v, p.psiE[1] + po + 4 * v);
              }
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
      if (tid == 0) ok = ok && wait_deps(it);
      ok = bar_compute_and(NTc, ok);
      const int PLn = PL + 1 == X ? 0 : PL + 1;
      if (ok && it <= X) issue(PL, PLn, it + 1, it, it == 0, it >= 1);
      cp_async_commit();
      PL = PLn;
    }

    float hyp[K][VW], hzp[K][VW];                  // H^{n+1/2}[P-1] of the thread's own cells
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int v = 0; v < VW; ++v) { hyp[k][v] = 0.f; hzp[k][v] = 0.f; }
    float an[K][3];                                // absorber rows of the NEXT plane
#pragma unroll
    for (int k = 0; k < K; ++k) { an[k][0] = 0.f; an[k][1] = 0.f; an[k][2] = 0.f; }
    int P = wrapi(cstart - 1, X);                  // plane of iteration i (i = 0: prologue plane)
    int se = 0, sh = 0;                            // ring slots of E[P] and H/B/psi[P]

#pragma unroll 2
    for (int i = 0; i <= X && ok; ++i) {
      const bool real = i >= 1;
      const int Pn = P + 1 == X ? 0 : P + 1;
      const int sen = se + 1 == NE ? 0 : se + 1;   // slot of E[P+1]
      const size_t pP = (size_t)P * gP, psiP = (size_t)P * pplane;
      cp_async_wait<D - 1>();
      // One thread checks the dependencies of the loads issued below; barrier A_i carries the
      // verdict to everybody (and makes the landed ring slot visible).
      if (tid == 0) ok = wait_deps(i + D);
      ok = bar_compute_and(NTc, ok);               // A_i
      if (!ok) break;
      if (tid == 0) {
        // every store of iterations < i has been issued by all compute threads
        if (i >= 2) st_vol_s(&ctl.done, base_mine + (unsigned)(i - 1));
        st_vol_s(&ctl.front, iters_done + (unsigned)i);
      }
      {
        const int PLn = PL + 1 == X ? 0 : PL + 1;
        // loads of iteration i+D go to the slots freed by iteration i-1
        if (ok && i + D <= X)
          issue(PL, PLn, se == 0 ? NE - 1 : se - 1, sh == 0 ? NH - 1 : sh - 1, false, true);
        cp_async_commit();
        PL = PLn;
      }

      float a[K][3];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        a[k][0] = an[k][0]; a[k][1] = an[k][1]; a[k][2] = an[k][2];
        if (own[k] && i < X) {
          const size_t xy = (size_t)Pn * g.Y + yk[k];
          an[k][0] = __ldg(p.A + xy); an[k][1] = __ldg(p.A + XY + xy); an[k][2] = __ldg(p.A + 2 * XY + xy);
        }
      }
      const float4* const eC = sE + se * eslot;    // E^n[P]
      const float4* const eN = sE + sen * eslot;   // E^n[P+1]
      const float4* const hO = sH + sh * eslot;    // H^{n-1/2}[P]
      const float4* const bC = sB + sh * eslot;    // B[P]
      const float4* const pS = sP + sh * pslot + (K * cp * g.npg + (has_psi ? slot : 0)) * PV;
      const int pstep = g.npg * PV;                // item 1's psi vectors follow item 0's column

      float ex[K][VW], ey[K][VW], ez[K][VW], hx[K][VW], hy[K][VW], hz[K][VW];
      {
        float ah[VW], bh[VW], ikh[VW];
        load_tab(3, ah); load_tab(4, bh); load_tab(5, ikh);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          unpack(eC[f[k]], ex[k], T()); unpack(eC[ring + f[k]], ey[k], T());
          unpack(eC[2 * ring + f[k]], ez[k], T());
          unpack(hO[f[k]], hx[k], T()); unpack(hO[ring + f[k]], hy[k], T());
          unpack(hO[2 * ring + f[k]], hz[k], T());
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
          float ez_yp[VW], ex_yp[VW], ey_xp[VW], ez_xp[VW], psx[VW], psy[VW];
          if (k + 1 < K) {                         // next column is this thread's own
#pragma unroll
            for (int v = 0; v < VW; ++v) { ez_yp[v] = ez[(k + 1) % K][v]; ex_yp[v] = ex[(k + 1) % K][v]; }
          } else {
            const int nb = doH[K - 1] ? f[K - 1] + Zq : f[K - 1];
            unpack(eC[2 * ring + nb], ez_yp, T()); unpack(eC[nb], ex_yp, T());
          }
          unpack(eN[ring + f[k]], ey_xp, T()); unpack(eN[2 * ring + f[k]], ez_xp, T());
#pragma unroll
          for (int v = 0; v < VW; ++v) { psx[v] = 0.f; psy[v] = 0.f; }
          if (has_psi && doH[k]) { load_psi(pS + k * pstep, psx); load_psi(pS + k * pstep + npsi, psy); }
          float ex_top = __shfl_down_sync(0xffffffffu, ex[k][0], 1);
          float ey_top = __shfl_down_sync(0xffffffffu, ey[k][0], 1);
          if (fix_up) {
            float tmp[VW];
            unpack(eC[f[k] + 1], tmp, T()); ex_top = tmp[0];
            unpack(eC[ring + f[k] + 1], tmp, T()); ey_top = tmp[0];
          }
          if (top) { ex_top = 0.f; ey_top = 0.f; }
#pragma unroll
          for (int v = 0; v < VW; ++v) {
            const float exz = (v + 1 < VW) ? ex[k][(v + 1) % VW] : ex_top;
            const float eyz = (v + 1 < VW) ? ey[k][(v + 1) % VW] : ey_top;
            h_cell(ex[k][v], ey[k][v], ez[k][v], exz, eyz, ez_yp[v], ex_yp[v], ey_xp[v], ez_xp[v],
                   ah[v], bh[v], ikh[v], g.dt, psx[v], psy[v], hx[k][v], hy[k][v], hz[k][v]);
            hx[k][v] = round_store<T>(hx[k][v]); hy[k][v] = round_store<T>(hy[k][v]);
            hz[k][v] = round_store<T>(hz[k][v]);
          }
          if (has_psi && own[k] && real) {         // new psiH of the owned PML cells
            const size_t po = psiP + ppoff[k];
            store_psi(p.psiHs[wb][0] + po, psx);
            store_psi(p.psiHs[wb][1] + po, psy);
          }
        }
      }

      // pair boundary: the odd column's new H is the y-1 neighbour of the next thread's even one
      sX[tid] = pack(hz[K - 1], T());
      sX[NTc + tid] = pack(hx[K - 1], T());
      if (cfg.need_zfix) {                         // cross-warp z-1 neighbours need Hx, Hy of all
        sX[2 * NTc + tid] = pack(hy[K - 1], T());
#pragma unroll
        for (int k = 0; k + 1 < K; ++k) {
          sX[(3 + 2 * k) * NTc + tid] = pack(hx[k], T());
          sX[(4 + 2 * k) * NTc + tid] = pack(hy[k], T());
        }
      }
      bar_compute(NTc);                            // B_i
      if (real) {
        float ae[VW], be[VW], ike[VW];
        load_tab(0, ae); load_tab(1, be); load_tab(2, ike);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          float hx_bot = __shfl_up_sync(0xffffffffu, hx[k][VW - 1], 1);
          float hy_bot = __shfl_up_sync(0xffffffffu, hy[k][VW - 1], 1);
          if (fix_dn) {
            float tmp[VW];
            unpack(sX[(k == K - 1 ? 1 : 3 + 2 * k) * NTc + tid - 1], tmp, T()); hx_bot = tmp[VW - 1];
            unpack(sX[(k == K - 1 ? 2 : 4 + 2 * k) * NTc + tid - 1], tmp, T()); hy_bot = tmp[VW - 1];
          }
          if (bottom) { hx_bot = 0.f; hy_bot = 0.f; }
          if (own[k]) {
            const size_t offP = pP + coff[k];
            float hz_ym[VW], hx_ym[VW], qsx[VW], qsy[VW], b0[VW], b1[VW], b2[VW];
            if (k > 0) {                           // previous column is this thread's own
#pragma unroll
              for (int v = 0; v < VW; ++v) { hz_ym[v] = hz[(k + K - 1) % K][v]; hx_ym[v] = hx[(k + K - 1) % K][v]; }
            } else {                               // previous thread's last column
              unpack(sX[tid - Zq], hz_ym, T());
              unpack(sX[NTc + tid - Zq], hx_ym, T());
            }
#pragma unroll
            for (int v = 0; v < VW; ++v) { qsx[v] = 0.f; qsy[v] = 0.f; }
            if (has_psi) { load_psi(pS + k * pstep + 2 * npsi, qsx); load_psi(pS + k * pstep + 3 * npsi, qsy); }
            unpack(bC[f[k]], b0, T()); unpack(bC[ring + f[k]], b1, T()); unpack(bC[2 * ring + f[k]], b2, T());
#pragma unroll
            for (int v = 0; v < VW; ++v) {
              const float hxz = (v > 0) ? hx[k][(v + VW - 1) % VW] : hx_bot;
              const float hyz = (v > 0) ? hy[k][(v + VW - 1) % VW] : hy_bot;
              e_cell(hx[k][v], hy[k][v], hz[k][v], hxz, hyz, hz_ym[v], hx_ym[v], hyp[k][v], hzp[k][v],
                     ae[v], be[v], ike[v], a[k][0], a[k][1], a[k][2], b0[v], b1[v], b2[v], qsx[v],
                     qsy[v], ex[k][v], ey[k][v], ez[k][v]);
            }
            if (g.src_axis == 0 ? (P == sp0 || P == sp1) : src_thr[k])
              add_source<VW>(g, p.src, w0, w1, P, yk[k], q, ex[k], ey[k], ez[k]);
            store_vec<T, LD_CG>(p.Hs[wb][0] + offP, hx[k]);
            store_vec<T, LD_CG>(p.Hs[wb][1] + offP, hy[k]);
            store_vec<T, LD_CG>(p.Hs[wb][2] + offP, hz[k]);
            store_vec<T, LD_CG>(p.Es[wb][0] + offP, ex[k]);
            store_vec<T, LD_CG>(p.Es[wb][1] + offP, ey[k]);
            store_vec<T, LD_CG>(p.Es[wb][2] + offP, ez[k]);
            if (has_psi) {
              const size_t po = psiP + ppoff[k];
              store_psi(p.psiE[0] + po, qsx);
              store_psi(p.psiE[1] + po, qsy);
            }
            if (oi >= 0) {
#pragma unroll
              for (int v = 0; v < VW; ++v) {
                ex[k][v] = round_store<T>(ex[k][v]); ey[k][v] = round_store<T>(ey[k][v]);
                ez[k][v] = round_store<T>(ez[k][v]);
              }
              write_snapshot<VW>(g, p.out, oi, P, yk[k], q, ex[k], ey[k], ez[k], p.proj);
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int v = 0; v < VW; ++v) { hyp[k][v] = hy[k][v]; hzp[k][v] = hz[k][v]; }
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

// compute threads of a CTA that owns `tile_y` columns: one thread per z-vector of a column pair
inline int systolic2_compute_threads(const Geom& g, int tile_y, int K = 2) {
  return ((tile_y + 2 + K - 1) / K * g.Zq + 31) / 32 * 32;
}

template <typename T, int D>
size_t systolic2_smem_bytes(const Geom& g, int tile_y, int K = 2) {
  constexpr int PV = VecTraits<T>::VW / 4;
  const size_t npsi = (size_t)(tile_y + 2) * g.npg * PV;
  const size_t ring = (size_t)(tile_y + 2) * g.Zq;
  const bool zfix = 32 % g.Zq != 0;
  return sizeof(float4) * ((size_t)((D + 2) * 3 + 2 * (D + 1) * 3) * ring +
                           (size_t)(zfix ? 1 + 2 * K : 2) * systolic2_compute_threads(g, tile_y, K) +
                           (size_t)(D + 1) * 4 * npsi + 6 * (size_t)g.Zp / 4);
}

template <typename T, int D>
bool systolic2_configure_d(const Geom& g, int tile_y_req, int stages_req, int threads_req, int sms,
                           int l2_bytes, int K, SystolicCfg* cfg, std::string* why) {
  const int cap = K == 2 ? kSys2MaxCompute : 2 * kSys2MaxCompute;
  const int max_threads = threads_req > 0 ? (threads_req < cap ? threads_req : cap) : cap;
  if (g.Zq * 2 > max_threads) { *why = "z extent too large for one CTA"; return false; }
  int max_tile = K * (max_threads / g.Zq) - 2;
  // largest tile whose staging ring fits in shared memory
  while (max_tile >= 1 && (systolic2_compute_threads(g, max_tile, K) > max_threads ||
                           systolic2_smem_bytes<T, D>(g, max_tile, K) + 64 > 227 * 1024))
    --max_tile;
  if (max_tile < 1) { *why = "staging ring does not fit in shared memory"; return false; }
  if (tile_y_req > 0 && tile_y_req < max_tile) max_tile = tile_y_req;
  if (max_tile > g.Y) max_tile = g.Y;
  const int ntiles = (g.Y + max_tile - 1) / max_tile;
  const int widest = (g.Y + ntiles - 1) / ntiles;
  const int compute = systolic2_compute_threads(g, widest, K);
  cfg->tile_y = widest;
  cfg->ntiles = ntiles;
  cfg->cols = K;
  cfg->threads = compute + kSys2Service;
  cfg->need_zfix = (32 % g.Zq != 0);
  cfg->smem_bytes = (int)systolic2_smem_bytes<T, D>(g, widest, K);
  // >= 2D+4 is needed for deadlock freedom (a throttled stage has published i-2 indices while its
  // successor needs i'+D+2 of them to advance); the rest is slack for the
  // polling/publishing latency.
  cfg->max_lead = 2 * D + 8;
  cfg->pf_ahead = 6;
  cfg->svc_sleep_ns = 400;
  // tuning knobs (benchmark sweeps only; values below the deadlock bound are clamped)
  if (const char* e = getenv("B200FDTD_MAX_LEAD")) cfg->max_lead = atoi(e);
  if (const char* e = getenv("B200FDTD_PF_AHEAD")) cfg->pf_ahead = atoi(e);
  if (const char* e = getenv("B200FDTD_SVC_SLEEP")) cfg->svc_sleep_ns = atoi(e);
  if (cfg->max_lead < 2 * D + 4) cfg->max_lead = 2 * D + 4;
  if (cfg->pf_ahead < 0) cfg->pf_ahead = 0;
  cfg->trap_on_timeout = 1;
  int occ = 0;
  const void* fn = K == 2 ? (const void*)systolic2_kernel<T, D, 2> : (const void*)systolic2_kernel<T, D, 1>;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg->smem_bytes) !=
          cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, cfg->threads, cfg->smem_bytes) !=
          cudaSuccess || occ < 1) {
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
  const void* fn = cfg.cols == 2 ? (const void*)systolic2_kernel<T, D, 2>
                                 : (const void*)systolic2_kernel<T, D, 1>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       cfg.smem_bytes);
  if (e != cudaSuccess) return (int)e;
  Geom gg = g;
  Ptrs<T> pp = p;
  SystolicCfg cc = cfg;
  void* args[] = {&gg, &pp, &cc, &sync};
  e = cudaLaunchCooperativeKernel(fn,
                                  dim3(cfg.stages * cfg.ntiles), dim3(cfg.threads), args,
                                  cfg.smem_bytes, st);
  if (e != cudaSuccess) return (int)e;
  systolic_check_kernel<<<1, 1, 0, st>>>(sync + (size_t)cfg.stages * cfg.ntiles * kSysFlagStride);
  return (int)cudaGetLastError();
}

}  // namespace b200
