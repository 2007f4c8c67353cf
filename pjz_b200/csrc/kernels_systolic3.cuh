// Systolic kernel, TMA generation ("systolic_tma"): same decomposition and dependency protocol
// as kernels_systolic.cuh / kernels_systolic2.cuh, with the operand staging moved off the SM's
// load/store path onto the TMA unit.
//
// Why: in the cp.async version every compute thread issues 13-15 LDGSTS per plane (8 issue
// cycles each on the LSU path, plus 64-bit address arithmetic and predicates); together with the
// LDS/STG traffic the MIO path, not HBM or L2, bounded the kernel (~3.5 SM-cycles per
// cell-update regardless of how the warps were arranged).  Here:
//
//  * PRODUCER = lane 0 of the service warp.  A tile's columns y0-1 .. y0+Yt are CONTIGUOUS in
//    global memory for every component plane, and the shared-memory ring slot uses the same
//    order, so one `cp.async.bulk.shared.global` (1-D bulk copy, no tensor map) moves a whole
//    component plane of the tile: 13 bulk copies per plane (E x3, H x3, B x3, psi x4; the
//    wrapped halo column of the first/last tile is a separate copy).  Completion is signalled
//    through an mbarrier (`complete_tx`), one "full" barrier per ring stage.
//  * CONSUMERS = 7 compute warps, two adjacent columns per thread as before.  They wait on the
//    full barrier of the stage (mbarrier.try_wait.parity), compute, and each warp arrives on the
//    "empty" barrier of the stage when it is done reading it.  They issue no loads, run no
//    dependency logic and meet at ONE named barrier per plane (the H exchange).
//  * The service warp also owns the inter-CTA protocol: it waits for `empty`, publishes this
//    CTA's progress (st.release.gpu), polls the predecessor/successor counters (ld.relaxed.gpu)
//    until the next group's dependencies hold, fences generic->async proxy, issues the group,
//    and prefetches upcoming planes into L2 (cp.async.bulk.prefetch.L2).
//
// Group G (cumulative over sweeps) = the copies iteration G consumes: E[P_G + 1] -> E slot
// (G+1) % NE, H/B/psi[P_G] -> stage G % NH (+ E[P_G] -> E slot G % NE for the first iteration of
// a sweep).  Group G may be issued once every compute warp has finished iteration G-D-1.
#pragma once

#include <stdio.h>
#include <stdlib.h>

#include <string>

#include "fdtd_common.cuh"
#include "kernels_systolic.cuh"
#include "kernels_systolic2.cuh"

namespace b200 {

constexpr int kSys3MaxCompute = 224;   // 7 warps (8 warps per CTA => 255-register budget)
constexpr int kSys3Service = 32;

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async;" ::: "memory");
}

// Spin on an mbarrier phase with a watchdog: a dependency that never arrives becomes a trap
// (sticky CUDA error), never a hang.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity,
                                          unsigned* status) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0 = 0;
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      const unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 5000000000ull) {
        atomicCAS(status, 0u, 1u + blockIdx.x);
        __trap();
      }
    }
  }
}

template <typename T, int D>
__global__ void __launch_bounds__(kSys3MaxCompute + kSys3Service)
systolic3_kernel(const Geom g, const Ptrs<T> p, const SystolicCfg cfg, unsigned* sync) {
  constexpr int VW = VecTraits<T>::VW;
  constexpr int NE = D + 2, NH = D + 1;
  constexpr int PV = VW / 4;                     // float4 per psi vector
  extern __shared__ float4 smem[];
  __shared__ __align__(8) unsigned long long full_bar[NH], empty_bar[NH];
  const int NTc = blockDim.x - kSys3Service;     // compute threads
  const int NWc = NTc / 32;                      // compute warps
  const int tid = threadIdx.x, lane = tid & 31;
  const int Zq = g.Zq, X = g.X;
  const int S = cfg.stages, NT = cfg.ntiles;
  const int t = blockIdx.x % NT, j = blockIdx.x / NT;
  const int y0 = (int)((long long)t * g.Y / NT);
  const int Yt = (int)((long long)(t + 1) * g.Y / NT) - y0;
  const int ncols = Yt + 2;
  const int npsi = (cfg.tile_y + 2) * g.npg * PV; // float4 per psi array per slot
  const int ring = (cfg.tile_y + 2) * Zq;         // float4 per component per ring slot
  const int eslot = 3 * ring, pslot = 4 * npsi;   // float4 per E/H/B slot, per psi slot

  float4* const sE = smem;                                   // [NE][3][ring]
  float4* const sH = sE + (size_t)NE * eslot;                // [NH][3][ring]
  float4* const sB = sH + (size_t)NH * eslot;                // [NH][3][ring]
  const int xrows = cfg.need_zfix ? 5 : 2;
  float4* const sX0 = sB + (size_t)NH * eslot;               // [2][2 | 5][NTc] new H at the pair
  float4* const sP = sX0 + (size_t)2 * xrows * NTc;          //   boundaries, double-buffered
                                                             // sP: [NH][4][npsi]
  float4* const sT = sP + (size_t)NH * pslot;                // [6][Zp/4] CPML tables

  unsigned* const status = sync + (size_t)S * NT * kSysFlagStride;
  unsigned* const my_prog = sync + ((size_t)j * NT + t) * kSysFlagStride;
  const unsigned sweep_iters = (unsigned)X + 1u;
  // number of sweeps (stages) this CTA row performs
  const unsigned nsweeps = j < g.tt ? (unsigned)((g.tt - 1 - j) / S + 1) : 0u;
  const unsigned total_iters = nsweeps * sweep_iters;

  if (tid == 0) {
    for (int s = 0; s < NH; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NWc); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 6 * g.Zp / 4; i += blockDim.x)
    sT[i] = __ldg(reinterpret_cast<const float4*>(p.tab) + i);
  __syncthreads();

  // ============================ service warp: producer + protocol ================================
  if (tid >= NTc) {
    const int jp = (j + S - 1) % S, jn = (j + 1) % S;
    const unsigned* watch = sync + ((size_t)jp * NT + wrapi(t - 1 + (lane < 3 ? lane : 1), NT)) *
                                       kSysFlagStride;
    if (lane == 3) watch = sync + ((size_t)jn * NT + t) * kSysFlagStride;
    if (lane == 4) watch = status;
    // tile geometry: columns c = 0 .. ncols-1 map to y = y0-1+c; the interior run [ca, cb) needs
    // no wrap, the (at most two) wrapped halo columns are copied separately.
    const int ca = y0 == 0 ? 1 : 0, cb = (y0 + Yt == g.Y) ? ncols - 1 : ncols;
    const unsigned colB = (unsigned)(g.Zp * (int)sizeof(T));          // bytes per column
    const unsigned colP = (unsigned)(g.npg * VW * (int)sizeof(float)); // psi bytes per column
    const unsigned grpB = (unsigned)ncols * (9u * colB + 4u * colP);   // bytes of a normal group
    const int ylo = max(y0 - 1, 0), yhi = min(y0 + Yt, g.Y - 1);
    const unsigned pf_bytes = (unsigned)((yhi - ylo + 1) * g.Zp * (int)sizeof(T));
    const size_t pf_off = (size_t)ylo * g.Zp;
    unsigned pf_done = 0;
    unsigned avail = 0, next = 0;                  // cached counters
    unsigned published = 0;

    // Non-blocking event loop.  Gi = next group to issue, Gp = next iteration whose completion
    // is to be published.  Priority: (a) issue a group as soon as its ring stage is free and its
    // dependencies hold (this is what keeps the compute warps fed), (b) publish finished
    // iterations, (c) poll the neighbours' counters.  Never blocking in one duty while another
    // is pending is also what makes short domains (X < 3*stages) deadlock-free: a CTA keeps
    // publishing its finished planes while it waits for its own predecessor.
    unsigned Gi = 0, Gp = 0, prog_new = 0, naps = 0;
    unsigned long long idle_since = 0;
    while (Gp < total_iters) {
      bool progress = false;
      // ---- (a) issue group Gi: needs iteration Gi-D-1 finished (Gi <= Gp + D) and its deps
      const bool ring_free = Gi < total_iters && Gi <= Gp + (unsigned)D;
      if (ring_free) {
        const unsigned G = Gi;
        const unsigned sw = G / sweep_iters, it = G % sweep_iters;
        const int n = j + (int)sw * S;
        const unsigned base_prev = (j > 0 ? sw : sw - 1u) * (unsigned)X;
        const unsigned base_mine = sw * (unsigned)X;
        const unsigned need = n > 0 ? base_prev + (unsigned)min((int)it + 2, X) : 0u;
        const int lead = min((int)it, X) - 1 - cfg.max_lead;
        const unsigned need_next =
            (n + 1 < g.tt && j + 1 < S && lead > 0) ? base_mine + (unsigned)lead : 0u;
        if (avail >= need && next >= need_next) {
          // lanes 0..15 each own one array (E0-2 @P+1, H0-2, B0-2, psi x4 @P, and E0-2 @P for
          // the first iteration of a sweep) and issue its bulk copies.
          const int rb = n & 1;
          const int P = wrapi(n % X - 1 + (int)it, X), Pn = P + 1 == X ? 0 : P + 1;
          const size_t pP = (size_t)P * g.P, pN = (size_t)Pn * g.P;
          unsigned long long* bar = &full_bar[G % NH];
          const bool first = it == 0;
          if (lane == 0) {
            fence_proxy_async();                   // other CTAs' generic-proxy stores -> TMA reads
            mbar_expect_tx(bar, grpB + (first ? (unsigned)ncols * 3u * colB : 0u));
          }
          __syncwarp();
          const char* src = nullptr;               // start of the array's plane
          float4* dst = nullptr;                   // ring row 0 of the array
          unsigned cbytes = colB;                  // bytes per column
          int cunits = Zq;                         // float4 per column in the ring
          const int a = lane;
          if (a < 3) {
            src = reinterpret_cast<const char*>(p.Es[rb][a] + pN);
            dst = sE + ((G + 1u) % NE) * eslot + a * ring;
          } else if (a < 6) {
            src = reinterpret_cast<const char*>(p.Hs[rb][a - 3] + pP);
            dst = sH + (G % NH) * eslot + (a - 3) * ring;
          } else if (a < 9) {
            src = reinterpret_cast<const char*>(p.B[a - 6] + pP);
            dst = sB + (G % NH) * eslot + (a - 6) * ring;
          } else if (a < 13) {
            if (g.npg > 0) {
              const size_t po = (size_t)P * g.Y * g.npg * VW;
              const float* base = a < 11 ? p.psiHs[rb][a - 9] : p.psiE[a - 11];
              src = reinterpret_cast<const char*>(base + po);
              dst = sP + (G % NH) * pslot + (a - 9) * npsi;
              cbytes = colP;
              cunits = g.npg * PV;
            }
          } else if (a < 16 && first) {
            src = reinterpret_cast<const char*>(p.Es[rb][a - 13] + pP);
            dst = sE + (G % NE) * eslot + (a - 13) * ring;
          }
          if (src != nullptr) {
            // interior run of columns [ca, cb), then the (at most two) wrapped halo columns
            tma_load_1d(dst + ca * cunits, src + (size_t)(y0 - 1 + ca) * cbytes,
                        (unsigned)(cb - ca) * cbytes, bar);
            if (ca == 1) tma_load_1d(dst, src + (size_t)(g.Y - 1) * cbytes, cbytes, bar);
            if (cb == ncols - 1) tma_load_1d(dst + (ncols - 1) * cunits, src, cbytes, bar);
          }
          ++Gi;
          progress = true;
          // L2 prefetch of the planes of groups [Gi, Gi + pf_ahead)
          if (cfg.pf_ahead > 0 && lane >= 16 && lane < 25) {
            if (pf_done < Gi) pf_done = Gi;
            for (; pf_done < Gi + (unsigned)cfg.pf_ahead && pf_done < total_iters; ++pf_done) {
              const unsigned sw2 = pf_done / sweep_iters, it2 = pf_done % sweep_iters;
              const int n2 = j + (int)sw2 * S, rb2 = n2 & 1;
              const int P2 = wrapi(n2 % X - 1 + (int)it2, X), Pn2 = P2 + 1 == X ? 0 : P2 + 1;
              const int b = lane - 16;
              const T* base;
              int plane;
              if (b < 3) { base = p.Es[rb2][b]; plane = Pn2; }
              else if (b < 6) { base = p.Hs[rb2][b - 3]; plane = P2; }
              else { base = p.B[b - 6]; plane = P2; }
              prefetch_l2_bulk(base + (size_t)plane * g.P + pf_off, pf_bytes);
            }
          }
          __syncwarp();
          idle_since = 0;
          continue;                                // try to issue the next group right away
        }
      }
      // ---- (b) note finished iterations (cheap: frees their ring stage for the next issue)
      if (mbar_try_wait(&empty_bar[Gp % NH], (Gp / NH) & 1u)) {
        do {
          const unsigned sw = Gp / sweep_iters, it = Gp % sweep_iters;
          if (it >= 1) prog_new = sw * (unsigned)X + it;  // cumulative finished sweep indices
          ++Gp;
          // (testing phase Gp/NH of a stage is alias-free: its previous phase, iteration Gp-NH,
          // is already known to be complete)
        } while (Gp < total_iters && mbar_try_wait(&empty_bar[Gp % NH], (Gp / NH) & 1u));
        idle_since = 0;
        continue;                                  // give (a) the first chance
      }
      // ---- (c) publish (st.release.gpu = fence + store, ~1 us): only when nothing can be issued
      if (prog_new != published) {
        if (lane == 0) st_release_u32(my_prog, prog_new);
        published = prog_new;
        idle_since = 0;
        continue;
      }
      // ---- (d) blocked on dependencies: refresh the neighbours' counters; otherwise the compute
      //          warps are simply busy -> back off briefly
      if (ring_free) {
        unsigned v = 0xffffffffu;
        if (lane < 5) v = ld_relaxed_gpu_u32(watch);
        const unsigned v0 = __shfl_sync(0xffffffffu, v, 0), v1 = __shfl_sync(0xffffffffu, v, 1),
                       v2 = __shfl_sync(0xffffffffu, v, 2), v3 = __shfl_sync(0xffffffffu, v, 3),
                       v4 = __shfl_sync(0xffffffffu, v, 4);
        const unsigned na = min(v0, min(v1, v2));
        if (na != avail || v3 != next) progress = true;
        avail = na;
        next = v3;
        if (v4 != 0) __trap();                     // another CTA gave up
      } else {
        __nanosleep(32);
        if ((++naps & 4095u) != 0) continue;       // look at the clock only now and then
      }
      // ---- watchdog: nothing moved for 5 s -> report and trap (never hang)
      if (progress) {
        idle_since = 0;
      } else {
        const unsigned long long now = globaltimer_ns();
        if (idle_since == 0) idle_since = now;
        else if (now - idle_since > 5000000000ull) {
          atomicCAS(status, 0u, 1u + blockIdx.x);
          __trap();
        }
      }
    }
    if (prog_new != published && lane == 0) st_release_u32(my_prog, prog_new);   // final count
    return;
  }

  // ================================= compute warps ===============================================
  // One thread owns the 16-byte z-vector q of TWO adjacent columns (2cp, 2cp+1) of the loaded
  // tile.  The y+1 neighbour of the even column and the y-1 neighbour of the odd column are the
  // thread's own registers; only the pair boundary goes through shared memory.
  const int cp = tid / Zq, q = tid - cp * Zq;
  int f[2], yk[2];
  bool act[2], doH[2], own[2];
  unsigned coff[2];                                // element offset inside a plane (< 2^31)
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int c = 2 * cp + k;
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
  unsigned ppoff[2];                               // psi offset of the item inside a plane (< 2^31)
#pragma unroll
  for (int k = 0; k < 2; ++k)
    ppoff[k] = (unsigned)((yk[k] * g.npg + (has_psi ? slot : 0)) * VW);
  const bool fix_up = cfg.need_zfix && lane == 31 && q + 1 < Zq;
  const bool fix_dn = cfg.need_zfix && lane == 0 && q > 0;
  const bool top = q + 1 == Zq, bottom = q == 0;
  const size_t XY = (size_t)X * g.Y;
  const int tstride = g.Zp / 4;
  const float4* const tq = sT + q * PV;            // CPML table w of this z-group: tq[w*tstride + v]
  const int sp0 = g.src_pos, sp1 = wrapi(g.src_pos - 1, g.src_axis == 0 ? X : g.Y);
  bool src_thr[2];
#pragma unroll
  for (int k = 0; k < 2; ++k)
    src_thr[k] = g.src_axis == 1 ? (yk[k] == sp0 || yk[k] == sp1)
                                 : (g.src_axis == 2 && q == g.src_pos / VW);

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

  unsigned G = 0;                                  // cumulative iteration counter
  unsigned se = 0, sh = 0;                         // G % NE, G % NH
  unsigned ph = 0;                                 // (G / NH) & 1
  unsigned xb = 0;                                 // exchange-buffer parity (toggles per exchange)
  for (int n = j; n < g.tt; n += S) {
    const int rb = n & 1, wb = rb ^ 1;
    const int cstart = n % X;
    const int oi = snapshot_index(g, n);
    const float w0 = __ldg(p.wave + 2 * (size_t)n), w1 = __ldg(p.wave + 2 * (size_t)n + 1);

    float hyp[2][VW], hzp[2][VW];                  // H^{n+1/2}[P-1] of the thread's own cells
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int v = 0; v < VW; ++v) { hyp[k][v] = 0.f; hzp[k][v] = 0.f; }
    float an[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};   // absorber rows of the NEXT plane
    int P = wrapi(cstart - 1, X);                  // plane of iteration i (i = 0: prologue plane)

    for (int i = 0; i <= X; ++i) {
      const bool real = i >= 1;
      const int Pn = P + 1 == X ? 0 : P + 1;
      const unsigned sen = se + 1 == NE ? 0 : se + 1;   // slot of E[P+1]
      const size_t pP = (size_t)P * gP, psiP = (size_t)P * pplane;

      float a[2][3];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        a[k][0] = an[k][0]; a[k][1] = an[k][1]; a[k][2] = an[k][2];
        if (own[k] && i < X) {
          const size_t xy = (size_t)Pn * g.Y + yk[k];
          an[k][0] = __ldg(p.A + xy); an[k][1] = __ldg(p.A + XY + xy); an[k][2] = __ldg(p.A + 2 * XY + xy);
        }
      }
      mbar_wait(&full_bar[sh], ph, status);        // group G has landed

      const float4* const eC = sE + se * eslot;    // E^n[P]
      const float4* const eN = sE + sen * eslot;   // E^n[P+1]
      const float4* const hO = sH + sh * eslot;    // H^{n-1/2}[P]
      const float4* const bC = sB + sh * eslot;    // B[P]
      const float4* const pS = sP + sh * pslot + (2 * cp * g.npg + (has_psi ? slot : 0)) * PV;
      const int pstep = g.npg * PV;                // item 1's psi vectors follow item 0's column

      float ex[2][VW], ey[2][VW], ez[2][VW], hx[2][VW], hy[2][VW], hz[2][VW];
      {
        float ah[VW], bh[VW], ikh[VW];
        load_tab(3, ah); load_tab(4, bh); load_tab(5, ikh);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          unpack(eC[f[k]], ex[k], T()); unpack(eC[ring + f[k]], ey[k], T());
          unpack(eC[2 * ring + f[k]], ez[k], T());
          unpack(hO[f[k]], hx[k], T()); unpack(hO[ring + f[k]], hy[k], T());
          unpack(hO[2 * ring + f[k]], hz[k], T());
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          float ez_yp[VW], ex_yp[VW], ey_xp[VW], ez_xp[VW], psx[VW], psy[VW];
          if (k == 0) {
#pragma unroll
            for (int v = 0; v < VW; ++v) { ez_yp[v] = ez[1][v]; ex_yp[v] = ex[1][v]; }
          } else {
            const int nb = doH[1] ? f[1] + Zq : f[1];
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

      // pair boundary: the odd column's new H is the y-1 neighbour of the next thread's even one.
      // There is no other CTA-wide barrier in the loop, so the exchange buffer alternates by
      // iteration parity: a warp can be at most one barrier ahead of the slowest one.
      float4* const sX = sX0 + (size_t)xb * xrows * NTc;
      if (real) {
        xb ^= 1u;
        // E-phase operands that do not depend on the exchange are read before the barrier
        float b0[2][VW], b1[2][VW], b2[2][VW];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          unpack(bC[f[k]], b0[k], T()); unpack(bC[ring + f[k]], b1[k], T());
          unpack(bC[2 * ring + f[k]], b2[k], T());
        }
        sX[tid] = pack(hz[1], T());
        sX[NTc + tid] = pack(hx[1], T());
        if (cfg.need_zfix) {                       // cross-warp z-1 neighbours need Hx, Hy of both
          sX[2 * NTc + tid] = pack(hy[1], T());
          sX[3 * NTc + tid] = pack(hx[0], T());
          sX[4 * NTc + tid] = pack(hy[0], T());
        }
        bar_compute(NTc);                          // the one CTA-wide barrier per plane
        float ae[VW], be[VW], ike[VW];
        load_tab(0, ae); load_tab(1, be); load_tab(2, ike);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          float hx_bot = __shfl_up_sync(0xffffffffu, hx[k][VW - 1], 1);
          float hy_bot = __shfl_up_sync(0xffffffffu, hy[k][VW - 1], 1);
          if (fix_dn) {
            float tmp[VW];
            unpack(sX[(k == 1 ? NTc : 3 * NTc) + tid - 1], tmp, T()); hx_bot = tmp[VW - 1];
            unpack(sX[(k == 1 ? 2 * NTc : 4 * NTc) + tid - 1], tmp, T()); hy_bot = tmp[VW - 1];
          }
          if (bottom) { hx_bot = 0.f; hy_bot = 0.f; }
          if (own[k]) {
            const size_t offP = pP + coff[k];
            float hz_ym[VW], hx_ym[VW], qsx[VW], qsy[VW];
            if (k == 1) {
#pragma unroll
              for (int v = 0; v < VW; ++v) { hz_ym[v] = hz[0][v]; hx_ym[v] = hx[0][v]; }
            } else {                               // even column >= 2: previous thread's odd column
              unpack(sX[tid - Zq], hz_ym, T());
              unpack(sX[NTc + tid - Zq], hx_ym, T());
            }
#pragma unroll
            for (int v = 0; v < VW; ++v) { qsx[v] = 0.f; qsy[v] = 0.f; }
            if (has_psi) { load_psi(pS + k * pstep + 2 * npsi, qsx); load_psi(pS + k * pstep + 3 * npsi, qsy); }
#pragma unroll
            for (int v = 0; v < VW; ++v) {
              const float hxz = (v > 0) ? hx[k][(v + VW - 1) % VW] : hx_bot;
              const float hyz = (v > 0) ? hy[k][(v + VW - 1) % VW] : hy_bot;
              e_cell(hx[k][v], hy[k][v], hz[k][v], hxz, hyz, hz_ym[v], hx_ym[v], hyp[k][v], hzp[k][v],
                     ae[v], be[v], ike[v], a[k][0], a[k][1], a[k][2], b0[k][v], b1[k][v], b2[k][v],
                     qsx[v], qsy[v], ex[k][v], ey[k][v], ez[k][v]);
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
      // this warp is done with ring stage `sh` (and E slot `se`) and has issued its stores
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[sh]);
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int v = 0; v < VW; ++v) { hyp[k][v] = hy[k][v]; hzp[k][v] = hz[k][v]; }
      P = Pn;
      se = sen;
      if (++sh == NH) { sh = 0; ph ^= 1u; }
      ++G;
    }
  }
}

// same ring as the cp.async kernel plus the second H-exchange buffer
template <typename T, int D>
size_t systolic3_smem_bytes(const Geom& g, int tile_y) {
  const bool zfix = 32 % g.Zq != 0;
  return systolic2_smem_bytes<T, D>(g, tile_y) +
         sizeof(float4) * (size_t)(zfix ? 5 : 2) * systolic2_compute_threads(g, tile_y);
}

template <typename T, int D>
bool systolic3_configure_d(const Geom& g, int tile_y_req, int stages_req, int threads_req, int sms,
                           int l2_bytes, SystolicCfg* cfg, std::string* why) {
  const int max_threads = threads_req > 0 ? (threads_req < kSys3MaxCompute ? threads_req : kSys3MaxCompute)
                                          : kSys3MaxCompute;
  if (g.Zq * 2 > max_threads) { *why = "z extent too large for one CTA"; return false; }
  int max_tile = 2 * (max_threads / g.Zq) - 2;
  while (max_tile >= 1 && (systolic2_compute_threads(g, max_tile) > max_threads ||
                           systolic3_smem_bytes<T, D>(g, max_tile) + 128 > 227 * 1024))
    --max_tile;
  if (max_tile < 1) { *why = "staging ring does not fit in shared memory"; return false; }
  if (tile_y_req > 0 && tile_y_req < max_tile) max_tile = tile_y_req;
  if (max_tile > g.Y) max_tile = g.Y;
  const int ntiles = (g.Y + max_tile - 1) / max_tile;
  const int widest = (g.Y + ntiles - 1) / ntiles;
  const int compute = systolic2_compute_threads(g, widest);
  cfg->tile_y = widest;
  cfg->ntiles = ntiles;
  cfg->threads = compute + kSys3Service;
  cfg->need_zfix = (32 % g.Zq != 0);
  cfg->smem_bytes = (int)systolic3_smem_bytes<T, D>(g, widest);
  cfg->max_lead = 2 * D + 8;
  cfg->pf_ahead = 6;
  cfg->svc_sleep_ns = 0;
  cfg->cols = 2;
  if (const char* e = getenv("B200FDTD_MAX_LEAD")) cfg->max_lead = atoi(e);
  if (const char* e = getenv("B200FDTD_PF_AHEAD")) cfg->pf_ahead = atoi(e);
  if (cfg->max_lead < 2 * D + 4) cfg->max_lead = 2 * D + 4;
  if (cfg->pf_ahead < 0) cfg->pf_ahead = 0;
  cfg->trap_on_timeout = 1;
  int occ = 0;
  if (cudaFuncSetAttribute(systolic3_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           cfg->smem_bytes) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, systolic3_kernel<T, D>, cfg->threads,
                                                    cfg->smem_bytes) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    *why = "kernel does not fit on an SM";
    return false;
  }
  const long long capacity = (long long)occ * sms;
  if (ntiles > capacity) { *why = "more y-tiles than co-resident CTAs"; return false; }
  int stages = (int)(capacity / ntiles);
  const long long plane_bytes = g.P * (long long)sizeof(T) * 15;
  const int lag = D + 5;
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
int systolic3_launch_d(const Geom& g, const Ptrs<T>& p, const SystolicCfg& cfg, unsigned* sync,
                       cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(systolic3_kernel<T, D>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, cfg.smem_bytes);
  if (e != cudaSuccess) return (int)e;
  Geom gg = g;
  Ptrs<T> pp = p;
  SystolicCfg cc = cfg;
  void* args[] = {&gg, &pp, &cc, &sync};
  e = cudaLaunchCooperativeKernel((const void*)systolic3_kernel<T, D>,
                                  dim3(cfg.stages * cfg.ntiles), dim3(cfg.threads), args,
                                  cfg.smem_bytes, st);
  if (e != cudaSuccess) return (int)e;
  systolic_check_kernel<<<1, 1, 0, st>>>(sync + (size_t)cfg.stages * cfg.ntiles * kSysFlagStride);
  return (int)cudaGetLastError();
}

}  // namespace b200
