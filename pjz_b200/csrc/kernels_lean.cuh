// Systolic kernel, warp-autonomous generation ("systolic_lean"): the same stage/tile
// decomposition, ping-pong buffers and inter-CTA progress protocol as kernels_systolic.cuh and
// kernels_systolic2.cuh (read those headers first), specialised for the flagship geometry
// -- fp32 storage with a z-column of exactly 32 16-byte vectors (125 <= Z <= 128) -- so that
// ONE WARP = ONE PAIR OF ADJACENT y-COLUMNS and the plane loop contains no CTA-wide barrier.
//
//  * Lane q of warp w owns z-vector q of the tile-local columns 2w and 2w+1 (column 0 is the
//    y0-1 halo).  z+-1 neighbours are warp shuffles, the y neighbour inside the pair is the
//    thread's own registers, x-1 is carried in registers along the sweep.
//  * Every lane stages ITS OWN operand vectors (E^n[P+1] of its two columns and of the column
//    after the pair, H^{n-1/2}[P], psiH[P]) with cp.async.cg into a per-warp ring one plane
//    ahead and is the only reader of what it copied: cp.async.wait_group is all the
//    synchronisation the ring needs.  B[P], psiE[P] and the absorber row travel the same way
//    (plain loads issued at the top of the iteration were measured to stall the first
//    shared-memory read of the H half-step: they end up on the same scoreboard).  The copies of
//    one iteration go out in two bursts -- E[P+2] at the top, H/B/psi[P+1] after the H half-step
//    -- and form one commit group: a single burst of ~30 LDGSTS queued up in the load/store pipe
//    and held the address registers the next instructions wanted (long-scoreboard stalls).
//  * The only data that crosses warps is the freshly formed (Hz, Hx) of the pair's second
//    column -- the y-1 neighbour of the next warp's first column.  It goes through a 2-deep
//    per-warp shared-memory slot guarded by two monotonic counters (produced / consumed):
//    a warp waits for its NEIGHBOUR only, never for the whole CTA.
//  * Inter-CTA dependencies are checked by each warp against the shared-memory mirror the
//    service warp maintains (as in systolic_async); every warp reports its own finished sweep
//    index and the service warp publishes the minimum with st.release.gpu.
//
// All loop predicates are warp-uniform, array strides are powers of two, addresses are 32-bit
// vector indices.  Measured (ncu source page, cfg2): 744 hot-path instructions per warp and plane
// (a thread handles 2 columns x 4 z-cells: 93 per cell) with 7 compute warps / 240 registers,
// 1044 with 8 compute warps / 160 registers (the compiler re-materialises uniform values under
// the 168-register cap of 9 warps), ~1000 in systolic_async for the same 8 cells per thread.
// No spills in either build.
#pragma once

#include <stdio.h>
#include <stdlib.h>

#include <string>

#include "fdtd_common.cuh"
#include "kernels_systolic.cuh"
#include "kernels_systolic2.cuh"

namespace b200 {

constexpr int kLeanMaxWarps = 8;       // compute warps per CTA (+1 service warp: 224 registers)
constexpr int kLeanXR = 2;             // depth of the boundary-H exchange ring
constexpr int kLeanERows = 8;          // 512-byte rows per E slot   (3 slots: P, P+1, in flight)
constexpr int kLeanHRows = 12;         // 512-byte rows per H/B slot (2 slots: P, in flight)
constexpr int kSlabMaxStages = 36;     // y-slab sessions: most pipeline stages (courier counter table)
constexpr int kSlabCouriers = 8;       // y-slab sessions: courier CTAs (SMs kept free of tiles)

struct LeanCtl {
  unsigned avail;      // min over the three predecessor counters (raw, cumulative)
  unsigned next;       // successor stage's counter on this tile
  unsigned ok;         // 0 once any CTA gave up
  unsigned front;      // cumulative iteration index warp 0 has reached (L2 prefetch cursor)
  unsigned exited;     // compute warps that have left the time loop
  unsigned cour;       // slab edge tiles: planes of the step two back that the courier has carried off
  unsigned wdone[kLeanMaxWarps + 1];   // per warp: cumulative finished sweep indices
  unsigned hcnt[kLeanMaxWarps + 1];    // per warp: iterations whose boundary H is in the slot
  unsigned rcnt[kLeanMaxWarps + 1];    // per warp: iterations the next warp has consumed
};

__device__ __forceinline__ float4 lds16(const float4* p) { return *p; }

// The 128-byte line at p is dead: drop it from L2 without writing it back.
__device__ __forceinline__ void discard_l2_line(const void* p) {
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
}

// ---- y-slab sessions: halo exchange from inside the kernel (peer-mapped stores over NVLink) -----
// One GPU owns the columns [g.ylo, g.yhi) of its local array; columns ylo-1 and yhi are ghosts.
// The neighbouring GPUs run the same kernel on the same local geometry and behave exactly like
// the tiles t-1 / t+1 of the periodic single-GPU run: the warp that owns the slab's last column
// stores its new (E, H, psiH) also into the HIGH neighbour's low ghost column, the warp that owns
// the first column stores (Ex, Ez) into the LOW neighbour's high ghost column, and a COURIER CTA
// forwards the edge tiles' progress counters into the neighbours' mirror slots with
// st.release.sys (see the courier block in the kernel).  Consumers read ghost data and mirror
// counters from their OWN memory, so the k+3 rule and the WAR argument of DESIGN.md 4.5 carry over
// with "tile t-1 / t+1" = the neighbour's edge tile.  All workspaces have the same layout, so a
// peer address is the local address plus one byte offset per side.
struct SlabPeers {
  long long delta_lo, delta_hi;    // neighbour's workspace base minus mine (bytes); 0 = myself
  int enabled;
  int edge_lead;                   // extra planes of max_lead for the tiles next to a slab edge
  int diag;                        // B200FDTD_SLAB_DIAG = 4: print the courier's push latencies
};

__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
template <typename P>
__device__ __forceinline__ P* peer_ptr(P* local, long long delta) {
  return reinterpret_cast<P*>(reinterpret_cast<char*>(local) + delta);
}

__device__ __forceinline__ void f4_to_arr(const float4 r, float (&v)[4]) {
  v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}
__device__ __forceinline__ float4 arr_to_f4(const float (&v)[4]) {
  return make_float4(v[0], v[1], v[2], v[3]);
}

// STATS = per-warp wait-time accounting printed at the end (debug builds of the plan only).
// SLAB = y-slab session: tiles cover [g.ylo, g.yhi), edge tiles exchange with the neighbour GPUs.
template <bool STATS, bool SLAB>
__global__ void __launch_bounds__(32 * (kLeanMaxWarps + 1), 1)
lean_kernel(const Geom g, const Ptrs<float> p, const SystolicCfg cfg, unsigned* sync,
            const SlabPeers peers) {
  constexpr int VW = 4;
  constexpr int ZQ = 32;
  extern __shared__ float4 smem[];
  __shared__ LeanCtl ctl;
  const int tid = threadIdx.x, lane = tid & 31;
  const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int NW = (int)(blockDim.x >> 5) - 1;       // compute warps
  const int S = cfg.stages, NT = cfg.ntiles;
  // block index -> (stage, tile): stage-major (consecutive blocks = the stages of one tile; round 2:
  // cfg2 95.8 -> 97.5 Gcell/s, cfg3 94.3 -> 96.0, cfg4 unchanged, 4096 x 512 slab 86.8 -> 89.9, the
  // board draws a little less power) or tile-major (B200FDTD_LEAN_MAP=0: round 1's order)
  const bool stage_major = cfg.block_order == 2 && blockIdx.x < (unsigned)(S * NT);
  const int t = stage_major ? (int)blockIdx.x / S : (int)(blockIdx.x % NT);
  const int j = stage_major ? (int)blockIdx.x % S : (int)(blockIdx.x / NT);
  const int ybase = SLAB ? g.ylo : 0, yspan = SLAB ? g.yhi - g.ylo : g.Y;
  const int y0 = ybase + (int)((long long)t * yspan / NT);
  const int Yt = ybase + (int)((long long)(t + 1) * yspan / NT) - y0;
  const int X = g.X, Y = g.Y;
  const int psi_row = g.npg;                       // float4 per psi row of a slot
  const int eslot_f4 = kLeanERows * ZQ;
  const int hslot_f4 = kLeanHRows * ZQ + 8 * psi_row + 4;   // + 8 psi rows, 2 absorber rows, 2 z-source rows
  const int warp_f4 = 3 * eslot_f4 + 2 * hslot_f4 + kLeanXR * 2 * ZQ;
  const int NWt = min(NW, (Yt + 2) / 2);           // warps with work on THIS tile (balanced tiles)

  unsigned* const status = sync + (size_t)S * NT * kSysFlagStride;
  unsigned* const my_prog = sync + ((size_t)j * NT + t) * kSysFlagStride;
  unsigned* const mirror_lo = status + kSysFlagStride;            // [S]: low neighbour's last tile
  unsigned* const mirror_hi = mirror_lo + (size_t)S * kSysFlagStride;   // high neighbour's first tile
  unsigned* const cour_lo = mirror_hi + (size_t)S * kSysFlagStride;     // [S]: courier progress, low side
  unsigned* const cour_hi = cour_lo + (size_t)S * kSysFlagStride;       // ... high side
  const bool edge_lo = SLAB && t == 0, edge_hi = SLAB && t == NT - 1;
  // A neighbour GPU's counter arrives ~6-9 us (three planes) later than a local tile's, so the tiles
  // an edge's delay reaches within one round (one tile further per stage) may run further ahead of
  // their successor stage before they are throttled; the L2 window grows for those tiles only.
  const int max_lead = cfg.max_lead + ((SLAB && (t < S || t >= NT - S)) ? peers.edge_lead : 0);

  if (tid < (int)(sizeof(LeanCtl) / sizeof(unsigned))) reinterpret_cast<unsigned*>(&ctl)[tid] = 0u;
  __syncthreads();
  if (tid == 0) ctl.ok = 1u;
  __syncthreads();

  // ==================================== courier CTA ===============================================
  // Slab sessions: up to eight extra CTAs (block indices >= S*NT, on SMs no tile uses) carry the halo
  // to the neighbour GPUs.  One warp per edge counter (side x stage): it watches the counter the edge
  // tile publishes (ld.acquire.gpu), copies the planes that counter newly covers -- the slab's
  // last owned column (E, H, psiH) into the HIGH neighbour's low ghost column, the first owned
  // column (Ex, Ez) into the LOW neighbour's high ghost column, read from the local L2, stored
  // through the peer mapping -- and then forwards the counter into the neighbour's mirror slot
  // with st.release.sys, which orders the warp's own peer stores before it.
  // Why not from the tiles themselves: a system-scope release costs ~3 us (5 400-6 000 cycles
  // measured against a plane time of 2.4 us) and stalls the memory pipeline of the SM that issues it
  // (1024 x 512 x 128 slab, one GPU wrapped onto itself: 31 Gcell/s with a system fence in the storing
  // warps, 79 with the pushes in the edge tiles' service warp, 67-72 with a tenth warp, 106 with
  // counters forwarded from here), and peer stores issued by a compute warp throttle that warp
  // across NVLink (two GPUs: 80 Gcell/s per GPU against 106 wrapped).  Here both sit on an idle SM
  // and only add latency, which the pipeline's slack (max_lead) absorbs.
  if (SLAB && blockIdx.x >= (unsigned)(S * NT)) {
    __shared__ unsigned cour_last[2 * kSlabMaxStages];       // count already carried, per counter
    const int nwarps = (int)(blockDim.x >> 5), wid = tid >> 5;
    const int NC = (int)gridDim.x - S * NT, ci = (int)blockIdx.x - S * NT;   // courier CTAs, mine
    for (int c = tid; c < 2 * S; c += (int)blockDim.x) cour_last[c] = 0u;
    __syncthreads();
    const unsigned PVn = (unsigned)Y * ZQ, PPn = (unsigned)Y * g.npg;
    unsigned spins = 0, n_push = 0;
    long long t_push = 0;
    bool busy = true, give_up = false;
    while (busy && !give_up) {
      busy = false;
      // counter c = side * S + stage; dealt across the courier CTAs first (a system-scope fence
      // waits for every peer store in flight on its SM, whoever issued it: 8 GPUs, all counters on
      // one SM: 34-43 k cycles per push on some ranks against 12 k with two GPUs), then across warps
      for (int c = ci + wid * NC; c < 2 * S; c += nwarps * NC) {
        const unsigned last_c = cour_last[c];
        const int side = c / S, jj = c % S;
        const int left = g.tt - g.n0 - jj;                   // steps n0+jj, n0+jj+S, ... < tt
        const unsigned final_count = left > 0 ? (unsigned)((left + S - 1) / S) * (unsigned)X : 0u;
        if (last_c >= final_count) continue;
        busy = true;
        const unsigned v = ld_acquire_u32(sync + ((size_t)jj * NT + (side == 0 ? 0 : NT - 1)) * kSysFlagStride);
        if (v == last_c) continue;
        const long long c0 = clock64();
        const long long delta = side == 0 ? peers.delta_lo : peers.delta_hi;
        const unsigned ysrc = (unsigned)(side == 0 ? g.ylo : g.yhi - 1);       // my edge column
        const unsigned ydst = (unsigned)(side == 0 ? g.yhi : g.ylo - 1);       // the neighbour's ghost
        for (unsigned cnt = last_c + 1u; cnt <= v; ++cnt) {
          const unsigned m = (cnt - 1u) / (unsigned)X, i = (cnt - 1u) % (unsigned)X + 1u;
          const int n = g.n0 + jj + (int)m * S;
          const unsigned P = (unsigned)wrapi(wrapi(n % X - 1, X) + (int)i, X);   // plane of sweep index i
          const int wb = (n & 1) ^ 1;
          const unsigned so = P * PVn + ysrc * ZQ + lane, dof = P * PVn + ydst * ZQ + lane;
          if (side == 0) {                                   // first owned column -> (Ex, Ez)
            const float4 a0 = __ldcg(reinterpret_cast<const float4*>(p.Es[wb][0]) + so);
            const float4 a2 = __ldcg(reinterpret_cast<const float4*>(p.Es[wb][2]) + so);
            __stcg(peer_ptr(reinterpret_cast<float4*>(p.Es[wb][0]) + dof, delta), a0);
            __stcg(peer_ptr(reinterpret_cast<float4*>(p.Es[wb][2]) + dof, delta), a2);
          } else {                                           // last owned column -> E, H, psiH
            float4 r[6];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
              r[cc] = __ldcg(reinterpret_cast<const float4*>(p.Es[wb][cc]) + so);
              r[3 + cc] = __ldcg(reinterpret_cast<const float4*>(p.Hs[wb][cc]) + so);
            }
            float4 ps[2];
            const bool pl = lane < g.npg;
            const unsigned pso = P * PPn + ysrc * g.npg + lane, pdo = P * PPn + ydst * g.npg + lane;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc)
              if (pl) ps[cc] = __ldcg(reinterpret_cast<const float4*>(p.psiHs[wb][cc]) + pso);
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
              __stcg(peer_ptr(reinterpret_cast<float4*>(p.Es[wb][cc]) + dof, delta), r[cc]);
              __stcg(peer_ptr(reinterpret_cast<float4*>(p.Hs[wb][cc]) + dof, delta), r[3 + cc]);
            }
#pragma unroll
            for (int cc = 0; cc < 2; ++cc)
              if (pl) __stcg(peer_ptr(reinterpret_cast<float4*>(p.psiHs[wb][cc]) + pdo, delta), ps[cc]);
          }
        }
        __syncwarp();
        if (lane == 0) {
          // the edge column's planes up to v have been read: the step after next may overwrite them
          st_release_u32((side == 0 ? cour_lo : cour_hi) + (size_t)jj * kSysFlagStride, v);
          unsigned* const mirror = side == 0 ? peer_ptr(mirror_hi + (size_t)jj * kSysFlagStride, peers.delta_lo)
                                             : peer_ptr(mirror_lo + (size_t)jj * kSysFlagStride, peers.delta_hi);
          st_release_sys_u32(mirror, v);
        }
        __syncwarp();
        t_push += clock64() - c0; ++n_push;
        if (lane == 0) cour_last[c] = v;
        __syncwarp();
      }
      if ((++spins & 4095u) == 0 && ld_relaxed_gpu_u32(status) != 0) give_up = true;   // somebody gave up
    }
    if ((peers.diag & 4) && lane == 0 && wid < 2)
      printf("courier warp %d: %u pushes, %.0f cycles each\n", wid, n_push,
             (double)t_push / (n_push ? n_push : 1));
    return;
  }

  // =================================== service warp ==============================================
  if (w == NW) {
    const int jp = (j + S - 1) % S, jn = (j + 1) % S;
    const unsigned* watch = sync + ((size_t)jp * NT + wrapi(t - 1 + (lane < 3 ? lane : 1), NT)) *
                                       kSysFlagStride;
    if constexpr (SLAB) {                          // beyond the slab: the neighbour GPU's edge tile
      if (lane == 0 && t == 0) watch = mirror_lo + (size_t)jp * kSysFlagStride;
      if (lane == 2 && t == NT - 1) watch = mirror_hi + (size_t)jp * kSysFlagStride;
    }
    if (lane == 3) watch = sync + ((size_t)jn * NT + t) * kSysFlagStride;
    if (lane == 4) watch = status;
    // slab edge tiles: the courier's progress on the step two before this CTA's (always stage j - 2)
    const int jq = (j + 2 * S - 2) % S;
    if (lane == 5) watch = cour_lo + (size_t)jq * kSysFlagStride;
    if (lane == 6) watch = cour_hi + (size_t)jq * kSysFlagStride;
    const bool watch_cour = (lane == 5 && edge_lo) || (lane == 6 && edge_hi);
    // L2 prefetch duty: lanes 8..16 own one array each (E0..2, H0..2 of the read set, B0..2).
    const int ylo = max(y0 - 1, 0), yhi = min(y0 + Yt, Y - 1);
    const unsigned pf_bytes = (unsigned)((yhi - ylo + 1) * g.Zp * (int)sizeof(float));
    const size_t pf_off = (size_t)ylo * g.Zp;
    unsigned pf_done = 0;
    const unsigned sweep_iters = (unsigned)X + 1u;
    unsigned published = 0;
    while (true) {
      const unsigned ex = ld_vol_s(&ctl.exited);
      unsigned dn = lane < NWt ? ld_vol_s(&ctl.wdone[lane]) : 0xffffffffu;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dn = min(dn, __shfl_xor_sync(0xffffffffu, dn, o));
      if (dn != published) {
        if (lane == 0) {
          st_release_u32(my_prog, dn);
        }
        published = dn;
      } else if (ex == (unsigned)NW) {
        break;
      }
      unsigned v = 0xffffffffu;
      if (SLAB && ((lane == 0 && t == 0) || (lane == 2 && t == NT - 1)))
        v = ld_relaxed_sys_u32(watch);               // a mirror slot: written by the neighbour GPU's courier
      else if (lane < 5 || watch_cour) v = ld_relaxed_gpu_u32(watch);
      const unsigned v0 = __shfl_sync(0xffffffffu, v, 0), v1 = __shfl_sync(0xffffffffu, v, 1),
                     v2 = __shfl_sync(0xffffffffu, v, 2), v3 = __shfl_sync(0xffffffffu, v, 3),
                     v4 = __shfl_sync(0xffffffffu, v, 4);
      if (lane == 0) {
        st_vol_s(&ctl.avail, min(v0, min(v1, v2)));
        st_vol_s(&ctl.next, v3);
        if (v4 != 0) st_vol_s(&ctl.ok, 0u);
      }
      if constexpr (SLAB) {
        const unsigned v5 = __shfl_sync(0xffffffffu, v, 5), v6 = __shfl_sync(0xffffffffu, v, 6);
        if (lane == 0) st_vol_s(&ctl.cour, min(v5, v6));
      }
      const unsigned front = ld_vol_s(&ctl.front);
      const unsigned want = front + 1u + (unsigned)cfg.pf_ahead;
      if (cfg.pf_ahead > 0 && lane >= 8 && lane < 17) {
        if (pf_done < front + 1u) pf_done = front + 1u;
        for (; pf_done < want; ++pf_done) {
          const unsigned sweep = pf_done / sweep_iters, it = pf_done % sweep_iters;
          const int n = g.n0 + j + (int)sweep * S;
          if (n >= g.tt) break;
          // Only planes the previous step has already produced: prefetching a plane that is
          // about to be overwritten would fetch dead data from HBM.  (A follower finds them in
          // L2 anyway; the prefetch matters for the stage that leads the window.)
          if (n > g.n0) {
            const unsigned m = (unsigned)((n - g.n0) / S);
            const unsigned need = (j > 0 ? m : m - 1u) * (unsigned)X + (unsigned)min((int)it + 2, X);
            if (min(v0, min(v1, v2)) < need) break;
          }
          const int rb = n & 1;
          const int P = wrapi(n % X - 1 + (int)it, X), Pn = wrapi(P + 1, X);
          const int a = lane - 8;
          const float* base;
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
  if (w >= NWt) {                                  // narrower tile: this warp has no column
    if (lane == 0) atomicAdd(&ctl.exited, 1u);
    return;
  }
  const int q = lane;
  const int cA = 2 * w, cB = 2 * w + 1;            // tile-local columns; 0 is the y0-1 halo
  const bool ownA = cA >= 1;                       // (cA <= Yt by construction of NWt)
  const bool doHB = cB <= Yt;                      // column B forms H and is owned
  const int yA = wrapi(y0 - 1 + cA, Y), yB = wrapi(y0 - 1 + cB, Y);
  const int yC = wrapi(y0 - 1 + (doHB ? cB + 1 : cB), Y);
  const unsigned PVn = (unsigned)Y * ZQ;           // vectors per x-plane
  const unsigned tvA = (unsigned)yA * ZQ + q, tvB = (unsigned)yB * ZQ + q, tvC = (unsigned)yC * ZQ + q;
  const int slot = psi_slot(g, q);
  const bool has_psi = slot >= 0;
  const unsigned PPn = (unsigned)Y * g.npg;        // psi vectors per x-plane
  const unsigned pvA = (unsigned)yA * g.npg + (has_psi ? slot : 0);
  const unsigned pvB = (unsigned)yB * g.npg + (has_psi ? slot : 0);
  const bool top = q + 1 == ZQ, bottom = q == 0;
  const bool discA = cA >= 2 && cA <= Yt - 1, discB = cB >= 2 && cB <= Yt - 1;

  float4* const wbase = smem + (size_t)w * warp_f4;
  float4* const hbase = wbase + 3 * eslot_f4;
  float4* const xmine = hbase + 2 * hslot_f4 + q;                // boundary-H slots this warp writes
  const float4* const xprev = xmine - warp_f4;                   // ... and those of warp w-1

  // CPML tables of this z-group live in registers for the whole run
  float ae[VW], be[VW], ike[VW], ah[VW], bh[VW], ikh[VW];
  f4_to_arr(__ldg(reinterpret_cast<const float4*>(p.tab + 0 * g.Zp) + q), ae);
  f4_to_arr(__ldg(reinterpret_cast<const float4*>(p.tab + 1 * g.Zp) + q), be);
  f4_to_arr(__ldg(reinterpret_cast<const float4*>(p.tab + 2 * g.Zp) + q), ike);
  f4_to_arr(__ldg(reinterpret_cast<const float4*>(p.tab + 3 * g.Zp) + q), ah);
  f4_to_arr(__ldg(reinterpret_cast<const float4*>(p.tab + 4 * g.Zp) + q), bh);
  f4_to_arr(__ldg(reinterpret_cast<const float4*>(p.tab + 5 * g.Zp) + q), ikh);

  // plane source: cheap warp-uniform pre-tests so that add_source() stays off the common path
  const int sp0 = g.src_pos, sp1 = wrapi(g.src_pos - 1, g.src_axis == 0 ? X : Y);
  const bool srcA = g.src_axis == 1 ? (yA == sp0 || yA == sp1) : g.src_axis == 2;
  const bool srcB = g.src_axis == 1 ? (yB == sp0 || yB == sp1) : g.src_axis == 2;
  const float dt = g.dt;
  const float4* const A4 = reinterpret_cast<const float4*>(p.A4);
  const float4* const S4 = reinterpret_cast<const float4*>(p.S4);
  // z-plane source: staged with the coefficients (one packed 16-byte row per column); only the
  // lane holding z-group src_pos / 4 is hit, at element src_pos % 4
  const bool zsrc = g.src_axis == 2;
  const bool zhit = zsrc && q == g.src_pos / VW;
  const int zidx = g.src_pos % VW;

  long long st_cp = 0, st_avail = 0, st_next = 0, st_rc = 0, st_hc = 0;
  const long long st_begin = STATS ? clock64() : 0;
  bool ok = true;
  unsigned kk = 0;                                 // cumulative iteration count (never reset)
  unsigned iters_done = 0;

  // Spin (all lanes, warp-uniform verdict) until cond() holds; false = give up.  The slow path
  // backs off with nanosleep so that a waiting warp leaves the issue slots (and the power
  // budget) to the warp it shares its scheduler with.
  auto spin = [&](auto cond) -> bool {
    if (__all_sync(0xffffffffu, cond())) return true;
    unsigned long long t0 = 0;
    unsigned spins = 0, ns = (unsigned)cfg.spin_ns0;
    while (true) {
      __nanosleep(ns);
      if (__all_sync(0xffffffffu, cond())) return true;
      if (__any_sync(0xffffffffu, ld_vol_s(&ctl.ok) == 0)) return false;
      if (ns < (unsigned)cfg.spin_ns_max) ns += ns;
      if ((++spins & 255u) == 0) {
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 5000000000ull) {
          if (lane == 0) { atomicCAS(status, 0u, 1u + blockIdx.x); st_vol_s(&ctl.ok, 0u); }
          return false;
        }
      }
    }
  };

  for (int n = g.n0 + j; n < g.tt && ok; n += S) {
    const int m = (n - g.n0) / S;
    const unsigned base_prev = (unsigned)((j > 0 ? m : m - 1)) * (unsigned)X;
    const unsigned base_mine = (unsigned)m * (unsigned)X;
    const bool has_prev = n > g.n0, has_next = n + 1 < g.tt && j + 1 < S;
    const int rb = n & 1, wb = rb ^ 1;
    const int cstart = n % X;
    const int oi = snapshot_index(g, n);
    const float w0 = __ldg(p.wave + 2 * (size_t)n), w1 = __ldg(p.wave + 2 * (size_t)n + 1);
    // z-plane source on one column: same operation order as add_source() (channel 0, then 1)
    auto zsource = [&](const float4 s, float (&ex)[VW], float (&ey)[VW]) {
      auto one = [&](float& e0, float& e1) {
        const float t0 = fmaf(w1, s.z, fmaf(w0, s.x, e0)), t1 = fmaf(w1, s.w, fmaf(w0, s.y, e1));
        e0 = zhit ? t0 : e0;
        e1 = zhit ? t1 : e1;
      };
      switch (zidx) {                                // warp-uniform
        case 0: one(ex[0], ey[0]); break;
        case 1: one(ex[1], ey[1]); break;
        case 2: one(ex[2], ey[2]); break;
        default: one(ex[3], ey[3]); break;
      }
    };
    const float4* const rEx = reinterpret_cast<const float4*>(p.Es[rb][0]);
    const float4* const rEy = reinterpret_cast<const float4*>(p.Es[rb][1]);
    const float4* const rEz = reinterpret_cast<const float4*>(p.Es[rb][2]);
    const float4* const rHx = reinterpret_cast<const float4*>(p.Hs[rb][0]);
    const float4* const rHy = reinterpret_cast<const float4*>(p.Hs[rb][1]);
    const float4* const rHz = reinterpret_cast<const float4*>(p.Hs[rb][2]);
    float4* const wEx = reinterpret_cast<float4*>(p.Es[wb][0]);
    float4* const wEy = reinterpret_cast<float4*>(p.Es[wb][1]);
    float4* const wEz = reinterpret_cast<float4*>(p.Es[wb][2]);
    float4* const wHx = reinterpret_cast<float4*>(p.Hs[wb][0]);
    float4* const wHy = reinterpret_cast<float4*>(p.Hs[wb][1]);
    float4* const wHz = reinterpret_cast<float4*>(p.Hs[wb][2]);
    const float4* const rPx = reinterpret_cast<const float4*>(p.psiHs[rb][0]);
    const float4* const rPy = reinterpret_cast<const float4*>(p.psiHs[rb][1]);
    float4* const wPx = reinterpret_cast<float4*>(p.psiHs[wb][0]);
    float4* const wPy = reinterpret_cast<float4*>(p.psiHs[wb][1]);
    float4* const ePx = reinterpret_cast<float4*>(p.psiE[0]);
    float4* const ePy = reinterpret_cast<float4*>(p.psiE[1]);
    const float4* const Bx = reinterpret_cast<const float4*>(p.B[0]);
    const float4* const By = reinterpret_cast<const float4*>(p.B[1]);
    const float4* const Bz = reinterpret_cast<const float4*>(p.B[2]);

    // The loads of iteration `it` read planes up to sweep index it+1 of the previous stage, which
    // must therefore have finished it+2 indices on tiles t-1, t, t+1 (the k+3 rule), and must not
    // run more than max_lead indices ahead of the next stage (keeps the window inside L2).
    auto wait_deps = [&](int it) -> bool {
      const unsigned need = has_prev ? base_prev + (unsigned)min(it + 2, X) : 0u;
      const int lead = min(it, X) - 1 - max_lead;
      const unsigned need_next = (has_next && lead > 0) ? base_mine + (unsigned)lead : 0u;
      if constexpr (SLAB) {
        // An edge tile overwrites (iteration it - 1, same buffer set) the plane that the step two
        // back wrote at its sweep index it + 1: the courier must have carried it off first.
        if ((edge_lo || edge_hi) && n - 2 >= g.n0) {
          const unsigned need_c = (unsigned)((n - 2 - g.n0) / S) * (unsigned)X + (unsigned)min(it + 1, X);
          if (!spin([&]() { return ld_vol_s(&ctl.cour) >= need_c; })) return false;
        }
      }
      if constexpr (STATS) {
        const long long c0 = clock64();
        const bool r0 = spin([&]() { return ld_vol_s(&ctl.avail) >= need; });
        const long long c1 = clock64();
        const bool r1 = r0 && spin([&]() { return ld_vol_s(&ctl.next) >= need_next; });
        st_avail += c1 - c0; st_next += clock64() - c1;
        return r1;
      }
      return spin([&]() { return ld_vol_s(&ctl.avail) >= need && ld_vol_s(&ctl.next) >= need_next; });
    };

    // Async copies of iteration `it` (plane PL): E[PL+1] -> E slot se; H, B, psi, absorber row
    // of PL -> H/B slot sh.  `ecoef` = the iteration performs an E half-step (it >= 1).
    // Straight-line on purpose: columns that a warp does not own (the halo column of warp 0, the
    // E-only column of the last warp) and the prologue plane are copied all the same -- the
    // addresses are valid, the values unused -- because predicating two thirds of these copies
    // on warp-uniform flags cost more instructions than the copies themselves.
    auto issue_e = [&](int PL, int PLn, float4* se, float4* se_first) {
      const unsigned vN = (unsigned)PLn * PVn, vP = (unsigned)PL * PVn;
      float4* const d = se + q;
      cp_async16(d + 0 * ZQ, rEx + (vN + tvA));
      cp_async16(d + 3 * ZQ, rEz + (vN + tvA));
      cp_async16(d + 6 * ZQ, rEy + (vN + tvA));
      cp_async16(d + 1 * ZQ, rEx + (vN + tvB));
      cp_async16(d + 4 * ZQ, rEz + (vN + tvB));
      cp_async16(d + 7 * ZQ, rEy + (vN + tvB));
      cp_async16(d + 2 * ZQ, rEx + (vN + tvC));
      cp_async16(d + 5 * ZQ, rEz + (vN + tvC));
      if (se_first) {                                // very first plane of the sweep: E[PL] too
        float4* const f = se_first + q;
        cp_async16(f + 0 * ZQ, rEx + (vP + tvA));
        cp_async16(f + 3 * ZQ, rEz + (vP + tvA));
        cp_async16(f + 6 * ZQ, rEy + (vP + tvA));
        cp_async16(f + 1 * ZQ, rEx + (vP + tvB));
        cp_async16(f + 4 * ZQ, rEz + (vP + tvB));
        cp_async16(f + 7 * ZQ, rEy + (vP + tvB));
        cp_async16(f + 2 * ZQ, rEx + (vP + tvC));
        cp_async16(f + 5 * ZQ, rEz + (vP + tvC));
      }
    };
    // The H/B/psi/absorber half of the copies is issued after the H half-step, not right behind
    // issue_e: the slot it fills (hnext) is idle for the whole iteration, and two shorter bursts
    // queue up less in the load/store pipe than one of ~30 (ncu: the first instructions that reuse
    // a copy's address registers sat on the long scoreboard; measured +1.5 % on cfg2).
    auto issue_h = [&](int PL, float4* sh) {
      const unsigned vP = (unsigned)PL * PVn;
      float4* const h = sh + q;
      cp_async16(h + 0 * ZQ, rHx + (vP + tvA));
      cp_async16(h + 2 * ZQ, rHy + (vP + tvA));
      cp_async16(h + 4 * ZQ, rHz + (vP + tvA));
      cp_async16(h + 1 * ZQ, rHx + (vP + tvB));
      cp_async16(h + 3 * ZQ, rHy + (vP + tvB));
      cp_async16(h + 5 * ZQ, rHz + (vP + tvB));
      cp_async16(h + 6 * ZQ, Bx + (vP + tvA));
      cp_async16(h + 8 * ZQ, By + (vP + tvA));
      cp_async16(h + 10 * ZQ, Bz + (vP + tvA));
      cp_async16(h + 7 * ZQ, Bx + (vP + tvB));
      cp_async16(h + 9 * ZQ, By + (vP + tvB));
      cp_async16(h + 11 * ZQ, Bz + (vP + tvB));
      if (lane < 2)
        cp_async16(sh + kLeanHRows * ZQ + 8 * psi_row + lane,
                   A4 + ((unsigned)PL * (unsigned)Y + (lane == 0 ? yA : yB)));
      if (zsrc && lane >= 2 && lane < 4)
        cp_async16(sh + kLeanHRows * ZQ + 8 * psi_row + lane,
                   S4 + ((unsigned)PL * (unsigned)Y + (lane == 2 ? yA : yB)));
      if (has_psi) {
        float4* const ps = sh + kLeanHRows * ZQ + slot;
        const unsigned pp = (unsigned)PL * PPn;
        cp_async16(ps, rPx + (pp + pvA));
        cp_async16(ps + 2 * psi_row, rPy + (pp + pvA));
        cp_async16(ps + psi_row, rPx + (pp + pvB));
        cp_async16(ps + 3 * psi_row, rPy + (pp + pvB));
        cp_async16(ps + 4 * psi_row, ePx + (pp + pvA));
        cp_async16(ps + 6 * psi_row, ePy + (pp + pvA));
        cp_async16(ps + 5 * psi_row, ePx + (pp + pvB));
        cp_async16(ps + 7 * psi_row, ePy + (pp + pvB));
      }
    };
    auto issue = [&](int PL, int PLn, float4* se, float4* sh, float4* se_first, bool /*ecoef*/) {
      issue_e(PL, PLn, se, se_first);
      issue_h(PL, sh);
    };

    int P = wrapi(cstart - 1, X);                  // plane of iteration i (i = 0: prologue plane)
    float4* sprev = wbase;                         // slot holding E[P]
    float4* scur = wbase + eslot_f4;               // slot holding E[P+1]
    float4* snext = wbase + 2 * eslot_f4;          // slot being filled with E[P+2]
    float4* hcur = hbase;                          // slot holding H, B, psi, absorber row of P
    float4* hnext = hbase + hslot_f4;              // ... being filled for P+1
    ok = wait_deps(0);
    if (ok) issue(P, P + 1 == X ? 0 : P + 1, scur, hcur, sprev, false);
    cp_async_commit();

    float hypA[VW], hzpA[VW], hypB[VW], hzpB[VW];  // H^{n+1/2}[P-1] of the thread's own cells
#pragma unroll
    for (int v = 0; v < VW; ++v) { hypA[v] = 0.f; hzpA[v] = 0.f; hypB[v] = 0.f; hzpB[v] = 0.f; }

#pragma unroll 1
    for (int i = 0; i <= X && ok; ++i) {
      const bool real = i >= 1;
      const int Pn = P + 1 == X ? 0 : P + 1;
      const unsigned vP = (unsigned)P * PVn;
      if constexpr (STATS) {
        const long long c0 = clock64();
        cp_async_wait<0>();
        st_cp += clock64() - c0;
      }
      cp_async_wait<0>();                          // this lane's copies of iteration i have landed
      if (i < X) {
        ok = wait_deps(i + 1);
        if (!ok) break;
        issue_e(Pn, Pn + 1 == X ? 0 : Pn + 1, snext, nullptr);
      }
      __syncwarp();                                // the absorber rows were copied by lanes 0, 1
      if (w == 0 && lane == 0) st_vol_s(&ctl.front, iters_done + (unsigned)i);
      // The field vectors that have just landed (E^n[P+1], H^{n-1/2}[P]) have now been read by
      // their only reader -- columns 2 .. Yt-1 are loaded by no other tile, and every plane but
      // the first one of the sweep is loaded exactly once -- and will be overwritten two steps
      // later: drop the dirty lines from L2 instead of letting them be written back to HBM.
      if (cfg.discard && i >= 1 && (q & 7) == 0) {
        const unsigned vN = (unsigned)Pn * PVn;
        if (discA) {                                 // (Ex, Ez of column A: see the E half-step)
          discard_l2_line(rEy + (vN + tvA));
          discard_l2_line(rHx + (vP + tvA)); discard_l2_line(rHy + (vP + tvA));
          discard_l2_line(rHz + (vP + tvA));
        }
        if (discB) {
          discard_l2_line(rEx + (vN + tvB)); discard_l2_line(rEy + (vN + tvB));
          discard_l2_line(rEz + (vN + tvB));
          discard_l2_line(rHx + (vP + tvB)); discard_l2_line(rHy + (vP + tvB));
          discard_l2_line(rHz + (vP + tvB));
        }
      }

      const unsigned pP = (unsigned)P * PPn;

      // ---------------------------------- H half-step ---------------------------------------------
      const float4* const ep = sprev + q;
      const float4* const ec = scur + q;
      const float4* const hc = hcur + q;
      const float4* const pc = hcur + kLeanHRows * ZQ + (has_psi ? slot : 0);
      float exA[VW], eyA[VW], ezA[VW], exB[VW], eyB[VW], ezB[VW];
      float hxA[VW], hyA[VW], hzA[VW], hxB[VW], hyB[VW], hzB[VW];
      float psxA[VW], psyA[VW], psxB[VW], psyB[VW];
      {
        float exC[VW], ezC[VW], eyxA[VW], ezxA[VW], eyxB[VW], ezxB[VW];
        f4_to_arr(lds16(ep + 0 * ZQ), exA); f4_to_arr(lds16(ep + 3 * ZQ), ezA);
        f4_to_arr(lds16(ep + 6 * ZQ), eyA);
        f4_to_arr(lds16(ep + 1 * ZQ), exB); f4_to_arr(lds16(ep + 4 * ZQ), ezB);
        f4_to_arr(lds16(ep + 7 * ZQ), eyB);
        f4_to_arr(lds16(ep + 2 * ZQ), exC); f4_to_arr(lds16(ep + 5 * ZQ), ezC);
        f4_to_arr(lds16(ec + 6 * ZQ), eyxA); f4_to_arr(lds16(ec + 3 * ZQ), ezxA);
        f4_to_arr(lds16(ec + 7 * ZQ), eyxB); f4_to_arr(lds16(ec + 4 * ZQ), ezxB);
        f4_to_arr(lds16(hc + 0 * ZQ), hxA); f4_to_arr(lds16(hc + 2 * ZQ), hyA);
        f4_to_arr(lds16(hc + 4 * ZQ), hzA);
        f4_to_arr(lds16(hc + 1 * ZQ), hxB); f4_to_arr(lds16(hc + 3 * ZQ), hyB);
        f4_to_arr(lds16(hc + 5 * ZQ), hzB);
#pragma unroll
        for (int v = 0; v < VW; ++v) { psxA[v] = 0.f; psyA[v] = 0.f; psxB[v] = 0.f; psyB[v] = 0.f; }
        if (has_psi) {
          f4_to_arr(lds16(pc), psxA); f4_to_arr(lds16(pc + 2 * psi_row), psyA);
          f4_to_arr(lds16(pc + psi_row), psxB); f4_to_arr(lds16(pc + 3 * psi_row), psyB);
        }
        float exA_top = __shfl_down_sync(0xffffffffu, exA[0], 1);
        float eyA_top = __shfl_down_sync(0xffffffffu, eyA[0], 1);
        float exB_top = __shfl_down_sync(0xffffffffu, exB[0], 1);
        float eyB_top = __shfl_down_sync(0xffffffffu, eyB[0], 1);
        if (top) { exA_top = 0.f; eyA_top = 0.f; exB_top = 0.f; eyB_top = 0.f; }
#pragma unroll
        for (int v = 0; v < VW; ++v) {
          const float exz = (v + 1 < VW) ? exA[(v + 1) % VW] : exA_top;
          const float eyz = (v + 1 < VW) ? eyA[(v + 1) % VW] : eyA_top;
          h_cell(exA[v], eyA[v], ezA[v], exz, eyz, ezB[v], exB[v], eyxA[v], ezxA[v], ah[v], bh[v],
                 ikh[v], dt, psxA[v], psyA[v], hxA[v], hyA[v], hzA[v]);
        }
#pragma unroll
        for (int v = 0; v < VW; ++v) {
          const float exz = (v + 1 < VW) ? exB[(v + 1) % VW] : exB_top;
          const float eyz = (v + 1 < VW) ? eyB[(v + 1) % VW] : eyB_top;
          h_cell(exB[v], eyB[v], ezB[v], exz, eyz, ezC[v], exC[v], eyxB[v], ezxB[v], ah[v], bh[v],
                 ikh[v], dt, psxB[v], psyB[v], hxB[v], hyB[v], hzB[v]);
        }
      }
      if (i < X) issue_h(Pn, hnext);               // second half of the copies for iteration i + 1
      cp_async_commit();
      // boundary H for the next warp: wait until it has consumed the slot's previous content
      {
        float4* const xs = xmine + (kk & (kLeanXR - 1)) * 2 * ZQ;
        if (kk >= (unsigned)kLeanXR && w + 1 < NWt) {
          const unsigned need = kk + 1u - (unsigned)kLeanXR;
          const long long c0 = STATS ? clock64() : 0;
          ok = spin([&]() { return ld_vol_s(&ctl.rcnt[w]) >= need; });
          if constexpr (STATS) st_rc += clock64() - c0;
          if (!ok) break;
        }
        xs[0] = arr_to_f4(hzB);
        xs[ZQ] = arr_to_f4(hxB);
        __syncwarp();
        if (lane == 0) st_vol_s(&ctl.hcnt[w], kk + 1u);
      }

      // ---------------------------------- E half-step ---------------------------------------------
      if (real) {
        float hzmA[VW], hxmA[VW];                  // (Hz, Hx) of the column before the pair
        if (w > 0) {
          const unsigned need = kk + 1u;
          const long long c0 = STATS ? clock64() : 0;
          ok = spin([&]() { return ld_vol_s(&ctl.hcnt[w - 1]) >= need; });
          if constexpr (STATS) st_hc += clock64() - c0;
          if (!ok) break;
          const float4* const xs = xprev + (kk & (kLeanXR - 1)) * 2 * ZQ;
          f4_to_arr(lds16(xs), hzmA); f4_to_arr(lds16(xs + ZQ), hxmA);
          __syncwarp();
          if (lane == 0) st_vol_s(&ctl.rcnt[w - 1], kk + 1u);
          // warp w-1 has finished the H half-step of this iteration, i.e. its copy of this
          // column's (Ex, Ez)[P+1] -- the only other reader -- has landed
          if (cfg.discard && discA && (q & 7) == 0) {
            const unsigned vN = (unsigned)Pn * PVn;
            discard_l2_line(rEx + (vN + tvA)); discard_l2_line(rEz + (vN + tvA));
          }
        } else {
#pragma unroll
          for (int v = 0; v < VW; ++v) { hzmA[v] = 0.f; hxmA[v] = 0.f; }
        }
        float hxA_bot = __shfl_up_sync(0xffffffffu, hxA[VW - 1], 1);
        float hyA_bot = __shfl_up_sync(0xffffffffu, hyA[VW - 1], 1);
        float hxB_bot = __shfl_up_sync(0xffffffffu, hxB[VW - 1], 1);
        float hyB_bot = __shfl_up_sync(0xffffffffu, hyB[VW - 1], 1);
        if (bottom) { hxA_bot = 0.f; hyA_bot = 0.f; hxB_bot = 0.f; hyB_bot = 0.f; }
        if (ownA) {
          float b0[VW], b1[VW], b2[VW], qsx[VW], qsy[VW];
          f4_to_arr(lds16(hc + 6 * ZQ), b0); f4_to_arr(lds16(hc + 8 * ZQ), b1);
          f4_to_arr(lds16(hc + 10 * ZQ), b2);
          const float4 aA = lds16(hcur + kLeanHRows * ZQ + 8 * psi_row);
#pragma unroll
          for (int v = 0; v < VW; ++v) { qsx[v] = 0.f; qsy[v] = 0.f; }
          if (has_psi) { f4_to_arr(lds16(pc + 4 * psi_row), qsx); f4_to_arr(lds16(pc + 6 * psi_row), qsy); }
#pragma unroll
          for (int v = 0; v < VW; ++v) {
            const float hxz = (v > 0) ? hxA[(v + VW - 1) % VW] : hxA_bot;
            const float hyz = (v > 0) ? hyA[(v + VW - 1) % VW] : hyA_bot;
            e_cell(hxA[v], hyA[v], hzA[v], hxz, hyz, hzmA[v], hxmA[v], hypA[v], hzpA[v], ae[v], be[v],
                   ike[v], aA.x, aA.y, aA.z, b0[v], b1[v], b2[v], qsx[v], qsy[v], exA[v], eyA[v], ezA[v]);
          }
          if (zsrc) zsource(lds16(hcur + kLeanHRows * ZQ + 8 * psi_row + 2), exA, eyA);
          else if (g.src_axis == 0 ? (P == sp0 || P == sp1) : srcA)
            add_source<VW>(g, p.src, w0, w1, P, yA, q, exA, eyA, ezA);
          const unsigned o = vP + tvA;
          __stcg(wHx + o, arr_to_f4(hxA)); __stcg(wHy + o, arr_to_f4(hyA)); __stcg(wHz + o, arr_to_f4(hzA));
          __stcg(wEx + o, arr_to_f4(exA)); __stcg(wEy + o, arr_to_f4(eyA)); __stcg(wEz + o, arr_to_f4(ezA));
          if (has_psi) {
            __stcg(wPx + (pP + pvA), arr_to_f4(psxA)); __stcg(wPy + (pP + pvA), arr_to_f4(psyA));
            __stcg(ePx + (pP + pvA), arr_to_f4(qsx)); __stcg(ePy + (pP + pvA), arr_to_f4(qsy));
          }
          if (oi >= 0) write_snapshot<VW>(g, p.out, oi, P, yA, q, exA, eyA, ezA, p.proj);
        }
        if (doHB) {
          float b0[VW], b1[VW], b2[VW], qsx[VW], qsy[VW];
          f4_to_arr(lds16(hc + 7 * ZQ), b0); f4_to_arr(lds16(hc + 9 * ZQ), b1);
          f4_to_arr(lds16(hc + 11 * ZQ), b2);
          const float4 aB = lds16(hcur + kLeanHRows * ZQ + 8 * psi_row + 1);
#pragma unroll
          for (int v = 0; v < VW; ++v) { qsx[v] = 0.f; qsy[v] = 0.f; }
          if (has_psi) { f4_to_arr(lds16(pc + 5 * psi_row), qsx); f4_to_arr(lds16(pc + 7 * psi_row), qsy); }
#pragma unroll
          for (int v = 0; v < VW; ++v) {
            const float hxz = (v > 0) ? hxB[(v + VW - 1) % VW] : hxB_bot;
            const float hyz = (v > 0) ? hyB[(v + VW - 1) % VW] : hyB_bot;
            e_cell(hxB[v], hyB[v], hzB[v], hxz, hyz, hzA[v], hxA[v], hypB[v], hzpB[v], ae[v], be[v],
                   ike[v], aB.x, aB.y, aB.z, b0[v], b1[v], b2[v], qsx[v], qsy[v], exB[v], eyB[v], ezB[v]);
          }
          if (zsrc) zsource(lds16(hcur + kLeanHRows * ZQ + 8 * psi_row + 3), exB, eyB);
          else if (g.src_axis == 0 ? (P == sp0 || P == sp1) : srcB)
            add_source<VW>(g, p.src, w0, w1, P, yB, q, exB, eyB, ezB);
          const unsigned o = vP + tvB;
          __stcg(wHx + o, arr_to_f4(hxB)); __stcg(wHy + o, arr_to_f4(hyB)); __stcg(wHz + o, arr_to_f4(hzB));
          __stcg(wEx + o, arr_to_f4(exB)); __stcg(wEy + o, arr_to_f4(eyB)); __stcg(wEz + o, arr_to_f4(ezB));
          if (has_psi) {
            __stcg(wPx + (pP + pvB), arr_to_f4(psxB)); __stcg(wPy + (pP + pvB), arr_to_f4(psyB));
            __stcg(ePx + (pP + pvB), arr_to_f4(qsx)); __stcg(ePy + (pP + pvB), arr_to_f4(qsy));
          }
          if (oi >= 0) write_snapshot<VW>(g, p.out, oi, P, yB, q, exB, eyB, ezB, p.proj);
        }
        // every store of sweep indices <= i has been issued by this warp
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          st_vol_s(&ctl.wdone[w], base_mine + (unsigned)i);
        }
      } else if (w > 0 && lane == 0) {
        st_vol_s(&ctl.rcnt[w - 1], kk + 1u);       // prologue plane: nothing to consume
      }
#pragma unroll
      for (int v = 0; v < VW; ++v) { hypA[v] = hyA[v]; hzpA[v] = hzA[v]; hypB[v] = hyB[v]; hzpB[v] = hzB[v]; }
      P = Pn;
      float4* const tmp = sprev; sprev = scur; scur = snext; snext = tmp;
      float4* const tmh = hcur; hcur = hnext; hnext = tmh;
      ++kk;
    }
    cp_async_wait<0>();
    iters_done += (unsigned)X + 1u;
  }
  cp_async_wait<0>();
  __syncwarp();
  if constexpr (STATS) {
    // every CTA, first and last warp: where the warp's time went, and on which SM (a CTA that never
    // waits is what the others wait for)
    if (lane == 0 && (w == 0 || w == NWt - 1)) {
      const double tot = (double)(clock64() - st_begin);
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      const double waits = (double)(st_cp + st_avail + st_next + st_rc + st_hc);
      printf("leanstats j %d t %d w %d sm %u iters %u cyc/iter %.0f busy/iter %.0f  cp %.3f avail %.3f next %.3f rcnt %.3f hcnt %.3f\n",
             j, t, w, smid, kk, tot / (kk ? kk : 1), (tot - waits) / (kk ? kk : 1), st_cp / tot,
             st_avail / tot, st_next / tot, st_rc / tot, st_hc / tot);
    }
  }
  if (lane == 0) atomicAdd(&ctl.exited, 1u);
}

inline const void* lean_fn(bool stats, bool slab) {
  if (slab) return (const void*)lean_kernel<false, true>;
  return stats ? (const void*)lean_kernel<true, false> : (const void*)lean_kernel<false, false>;
}

// Compute warps for a tile of `tile_y` owned columns: columns 0 .. tile_y form H, two per warp.
inline int lean_warps(int tile_y) { return (tile_y + 2) / 2; }

inline size_t lean_smem_bytes(const Geom& g, int tile_y) {
  const size_t eslot_f4 = (size_t)kLeanERows * 32;
  const size_t hslot_f4 = (size_t)kLeanHRows * 32 + 8 * (size_t)g.npg + 4;
  const size_t warp_f4 = 3 * eslot_f4 + 2 * hslot_f4 + (size_t)kLeanXR * 2 * 32;
  return sizeof(float4) * warp_f4 * lean_warps(tile_y);
}

inline bool lean_configure(const Geom& g, bool reduced, int tile_y_req, int stages_req, int sms,
                           int l2_bytes, SystolicCfg* cfg, std::string* why) {
  if (reduced) { *why = "fp32 storage only"; return false; }
  if (g.Zq != 32) { *why = "needs a z-column of exactly 32 vectors (125 <= Z <= 128)"; return false; }
  if (g.N / 4 * 3 >= (1ll << 32)) { *why = "domain too large for 32-bit vector indices"; return false; }
  int max_tile = 2 * kLeanMaxWarps - 1;          // 2*NW - 1 owned columns fill NW warps exactly
  while (max_tile >= 1 && lean_smem_bytes(g, max_tile) + 256 > 227 * 1024) --max_tile;
  if (max_tile < 1) { *why = "staging ring does not fit in shared memory"; return false; }
  if (tile_y_req > 0 && tile_y_req < max_tile) max_tile = tile_y_req;
  const int Yspan = g.yhi - g.ylo;               // the whole domain, or a slab's owned columns
  if (Yspan < 1) { *why = "empty column range"; return false; }
  if (max_tile > Yspan) max_tile = Yspan;
  const int ntiles = (Yspan + max_tile - 1) / max_tile;
  const int widest = (Yspan + ntiles - 1) / ntiles;
  cfg->tile_y = widest;
  cfg->ntiles = ntiles;
  cfg->cols = 2;
  cfg->threads = 32 * (lean_warps(widest) + 1);
  cfg->smem_bytes = (int)lean_smem_bytes(g, widest);
  cfg->max_lead = 10;
  // L2 prefetch by the service warp: OFF.  Measured on cfg2 (round 2, bench.py, 20 000 steps):
  // 103.4 Gcell/s with 6 planes of cp.async.bulk.prefetch.L2 ahead of the ring, 102.3 with 12,
  // 119.5 with none -- the ring's own copies run a plane (2.7 us) ahead, which already covers
  // the HBM latency, and the prefetches compete with them for L2 bandwidth and power.
  cfg->pf_ahead = 0;
  cfg->svc_sleep_ns = 200;
  cfg->spin_ns_max = 400;
  cfg->discard = 1;
  if (const char* e = getenv("B200FDTD_LEAN_DISCARD")) cfg->discard = atoi(e);
  if (const char* e = getenv("B200FDTD_SPIN_NS")) cfg->spin_ns_max = atoi(e);
  if (const char* e = getenv("B200FDTD_MAX_LEAD")) cfg->max_lead = atoi(e);
  if (const char* e = getenv("B200FDTD_PF_AHEAD")) cfg->pf_ahead = atoi(e);
  if (const char* e = getenv("B200FDTD_SVC_SLEEP")) cfg->svc_sleep_ns = atoi(e);
  if (cfg->max_lead < 6) cfg->max_lead = 6;
  if (cfg->pf_ahead < 0) cfg->pf_ahead = 0;
  cfg->trap_on_timeout = 1;
  int occ = 0;
  // back-off of a waiting warp: first sleep 200 ns, doubling up to 400 (round 2: 20 -> 160 before;
  // cfg2 96.2 -> 97.2 Gcell/s on a slower box, fp16 126.5 -> 129.0: half the polls, less power)
  cfg->spin_ns0 = 200;
  if (const char* e = getenv("B200FDTD_SPIN_NS0")) cfg->spin_ns0 = atoi(e) < 1 ? 1 : atoi(e);
  cfg->block_order = 2;                          // stage-major (tile-major: 1)
  if (const char* e = getenv("B200FDTD_LEAN_MAP")) cfg->block_order = atoi(e) != 0 ? 2 : 1;
  const bool slab = Yspan != g.Y;
  const void* fn = lean_fn(getenv("B200FDTD_LEAN_STATS") != nullptr, slab);
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg->smem_bytes) !=
          cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, cfg->threads, cfg->smem_bytes) !=
          cudaSuccess || occ < 1) {
    cudaGetLastError();
    *why = "kernel does not fit on an SM";
    return false;
  }
  // a slab's courier CTAs: up to kSlabCouriers SMs are kept free of tiles
  const long long capacity = (long long)occ * sms - (slab ? (sms >= 64 ? kSlabCouriers : 1) : 0);
  if (ntiles > capacity) { *why = "more y-tiles than co-resident CTAs"; return false; }
  int stages = (int)(capacity / ntiles);
  const long long plane_bytes = g.P * 4ll * 15;
  const int lag = 6;                             // planes a stage trails its predecessor by
  long long by_l2 = (long long)(l2_bytes * 0.8) / (lag * plane_bytes);
  if (by_l2 < 1) by_l2 = 1;
  if (stages > by_l2) stages = (int)by_l2;
  if (stages_req > 0 && stages_req <= capacity / ntiles) stages = stages_req;
  if (stages > g.tt) stages = g.tt > 0 ? g.tt : 1;
  if (stages > g.X) stages = g.X;
  if (slab && stages > kSlabMaxStages) stages = kSlabMaxStages;   // (the courier's counter table)
  cfg->stages = stages;
  cfg->l2_window_bytes = (long long)stages * lag * plane_bytes;
  return true;
}

inline int lean_launch(const Geom& g, const Ptrs<float>& p, const SystolicCfg& cfg, unsigned* sync,
                       cudaStream_t st, const SlabPeers* peers = nullptr) {
  const bool slab = peers != nullptr && peers->enabled;
  const void* fn = lean_fn(getenv("B200FDTD_LEAN_STATS") != nullptr, slab);
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       cfg.smem_bytes);
  if (e != cudaSuccess) return (int)e;
  Geom gg = g;
  Ptrs<float> pp = p;
  SystolicCfg cc = cfg;
  SlabPeers sp;
  sp.delta_lo = 0; sp.delta_hi = 0; sp.enabled = 0; sp.edge_lead = 0; sp.diag = 0;
  if (slab) {
    sp = *peers;
    sp.diag = 0;
    // 2 GPUs, 4096 x 512 x 128 per GPU, Gcell/s (wrapped single-GPU slab: 84-85 from 8 upwards):
    // 0: 110, 8: 156, 12: 161, 16: 165, 24: 168 -- NVLink adds ~3 us to the counter's flight
    sp.edge_lead = 24;
    if (const char* e = getenv("B200FDTD_SLAB_EDGE_LEAD")) sp.edge_lead = atoi(e) < 0 ? 0 : atoi(e);
    if (const char* e = getenv("B200FDTD_SLAB_DIAG")) sp.diag = atoi(e) & 4;
  }
  void* args[] = {&gg, &pp, &cc, &sync, &sp};   // (cc, sp finalised above)
  int couriers = 0;
  if (slab) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    couriers = sms >= 64 ? kSlabCouriers : 1;
    if (couriers > 2 * cfg.stages) couriers = 2 * cfg.stages;
  }
  e = cudaLaunchCooperativeKernel(fn, dim3(cfg.stages * cfg.ntiles + couriers), dim3(cfg.threads),
                                  args, cfg.smem_bytes, st);
  if (e != cudaSuccess) return (int)e;
  systolic_check_kernel<<<1, 1, 0, st>>>(sync + (size_t)cfg.stages * cfg.ntiles * kSysFlagStride);
  return (int)cudaGetLastError();
}

}  // namespace b200
