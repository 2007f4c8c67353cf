// "systolic_lean", one-column-per-warp variant: identical decomposition, protocol and per-lane
// staging as kernels_lean.cuh (read that header first), with ONE y-column per warp instead of a
// pair.  The L2 window caps the columns in flight at ~14 per SM whatever the tiling (stages x Y x
// 6 planes x 60 B x Z <= 0.8 x L2), and the two-column kernel is bound by dependent-issue latency
// at 2 warps per scheduler: splitting the same 14-15 columns over 15 warps doubles the warps each
// scheduler can pick from at about the same instruction count per column.
//
//  * warp w = tile-local column w (0 = the y0-1 halo, which only forms H); lane q = z-vector q;
//  * per-warp ring: E 3 slots x 5 rows (Ex, Ez, Ey of the column; Ex, Ez of the next column),
//    H 2 slots x 3 rows (+ psiH), and ONE slot of B / psiE / absorber row that is refilled at
//    the END of an iteration -- after the E half-step has read it -- for the next iteration
//    (two cp.async groups per iteration, wait_group 1 at both use points): 14.6 KB per warp;
//  * the boundary (Hz, Hx) exchange now happens at every column, still neighbour-to-neighbour.
#pragma once

#include "kernels_lean.cuh"

namespace b200 {

constexpr int kLean1MaxWarps = 15;     // compute warps per CTA (+1 service warp = 512 threads)
constexpr int kLean1ERows = 5;
constexpr int kLean1HRows = 3;

struct Lean1Ctl {
  unsigned avail, next, ok, front, exited;
  unsigned wdone[kLean1MaxWarps + 1];
  unsigned hcnt[kLean1MaxWarps + 1];
  unsigned rcnt[kLean1MaxWarps + 1];
};

template <bool STATS>
__global__ void __launch_bounds__(32 * (kLean1MaxWarps + 1), 1)
lean1_kernel(const Geom g, const Ptrs<float> p, const SystolicCfg cfg, unsigned* sync) {
  constexpr int VW = 4;
  constexpr int ZQ = 32;
  extern __shared__ float4 smem[];
  __shared__ Lean1Ctl ctl;
  const int tid = threadIdx.x, lane = tid & 31;
  const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int NW = (int)(blockDim.x >> 5) - 1;       // compute warps
  const int S = cfg.stages, NT = cfg.ntiles;
  const int t = blockIdx.x % NT, j = blockIdx.x / NT;
  const int y0 = (int)((long long)t * g.Y / NT);
  const int Yt = (int)((long long)(t + 1) * g.Y / NT) - y0;
  const int X = g.X, Y = g.Y;
  const int psi_row = g.npg;                       // float4 per psi row
  const int eslot_f4 = kLean1ERows * ZQ;
  const int hslot_f4 = kLean1HRows * ZQ + 2 * psi_row;            // H rows + psiH x, y
  const int bslot_f4 = 3 * ZQ + 2 * psi_row + 1;                  // B rows + psiE x, y + absorber row
  const int warp_f4 = 3 * eslot_f4 + 2 * hslot_f4 + bslot_f4 + kLeanXR * 2 * ZQ;
  const int NWt = min(NW, Yt + 1);                 // warps with a column on THIS tile

  unsigned* const status = sync + (size_t)S * NT * kSysFlagStride;
  unsigned* const my_prog = sync + ((size_t)j * NT + t) * kSysFlagStride;

  if (tid < (int)(sizeof(Lean1Ctl) / sizeof(unsigned))) reinterpret_cast<unsigned*>(&ctl)[tid] = 0u;
  __syncthreads();
  if (tid == 0) ctl.ok = 1u;
  __syncthreads();

  // =================================== service warp ==============================================
  if (w == NW) {
    const int jp = (j + S - 1) % S, jn = (j + 1) % S;
    const unsigned* watch = sync + ((size_t)jp * NT + wrapi(t - 1 + (lane < 3 ? lane : 1), NT)) *
                                       kSysFlagStride;
    if (lane == 3) watch = sync + ((size_t)jn * NT + t) * kSysFlagStride;
    if (lane == 4) watch = status;
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
        if (lane == 0) st_release_u32(my_prog, dn);
        published = dn;
      } else if (ex == (unsigned)NW) {
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
      const unsigned front = ld_vol_s(&ctl.front);
      const unsigned want = front + 1u + (unsigned)cfg.pf_ahead;
      if (cfg.pf_ahead > 0 && lane >= 8 && lane < 17) {
        if (pf_done < front + 1u) pf_done = front + 1u;
        for (; pf_done < want; ++pf_done) {
          const unsigned sweep = pf_done / sweep_iters, it = pf_done % sweep_iters;
          const int n = g.n0 + j + (int)sweep * S;
          if (n >= g.tt) break;
          if (n > g.n0) {                          // only planes the previous step has produced
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
  if (w >= NWt) {
    if (lane == 0) atomicAdd(&ctl.exited, 1u);
    return;
  }
  const int q = lane;
  const int cA = w;                                // tile-local column; 0 is the y0-1 halo
  const bool own = cA >= 1;
  const int yA = wrapi(y0 - 1 + cA, Y), yC = wrapi(y0 + cA, Y);
  const unsigned PVn = (unsigned)Y * ZQ;
  const unsigned tvA = (unsigned)yA * ZQ + q, tvC = (unsigned)yC * ZQ + q;
  const int slot = psi_slot(g, q);
  const bool has_psi = slot >= 0;
  const unsigned PPn = (unsigned)Y * g.npg;
  const unsigned pvA = (unsigned)yA * g.npg + (has_psi ? slot : 0);
  const bool top = q + 1 == ZQ, bottom = q == 0;
  const bool disc = cA >= 2 && cA <= Yt - 1;

  float4* const wbase = smem + (size_t)w * warp_f4;
  float4* const hbase = wbase + 3 * eslot_f4;
  float4* const bslot = hbase + 2 * hslot_f4;
  float4* const xmine = bslot + bslot_f4 + q;                    // boundary-H slots this warp writes
  const float4* const xprev = xmine - warp_f4;                   // ... and those of warp w-1

  float ae[VW], be[VW], ike[VW], ah[VW], bh[VW], ikh[VW];
  f4_to_arr(__ldg(reinterpret_cast<const float4*>(p.tab + 0 * g.Zp) + q), ae);
  f4_to_arr(__ldg(reinterpret_cast<const float4*>(p.tab + 1 * g.Zp) + q), be);
  f4_to_arr(__ldg(reinterpret_cast<const float4*>(p.tab + 2 * g.Zp) + q), ike);
  f4_to_arr(__ldg(reinterpret_cast<const float4*>(p.tab + 3 * g.Zp) + q), ah);
  f4_to_arr(__ldg(reinterpret_cast<const float4*>(p.tab + 4 * g.Zp) + q), bh);
  f4_to_arr(__ldg(reinterpret_cast<const float4*>(p.tab + 5 * g.Zp) + q), ikh);

  const int sp0 = g.src_pos, sp1 = wrapi(g.src_pos - 1, g.src_axis == 0 ? X : Y);
  const bool srcA = g.src_axis == 1 ? (yA == sp0 || yA == sp1) : g.src_axis == 2;
  const float dt = g.dt;
  const float4* const A4 = reinterpret_cast<const float4*>(p.A4);

  long long st_cp = 0, st_avail = 0, st_next = 0, st_rc = 0, st_hc = 0;
  const long long st_begin = STATS ? clock64() : 0;
  bool ok = true;
  unsigned kk = 0;
  unsigned iters_done = 0;

  auto spin = [&](auto cond) -> bool {
    if (__all_sync(0xffffffffu, cond())) return true;
    unsigned long long t0 = 0;
    unsigned spins = 0, ns = 20;
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

    auto wait_deps = [&](int it) -> bool {
      const unsigned need = has_prev ? base_prev + (unsigned)min(it + 2, X) : 0u;
      const int lead = min(it, X) - 1 - cfg.max_lead;
      const unsigned need_next = (has_next && lead > 0) ? base_mine + (unsigned)lead : 0u;
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

    // Group "top" of iteration `it` (plane PL): E[PL+1] -> E slot se, H / psiH[PL] -> H slot sh.
    auto issue_top = [&](int PL, int PLn, float4* se, float4* sh, float4* se_first) {
      const unsigned vN = (unsigned)PLn * PVn, vP = (unsigned)PL * PVn;
      float4* const d = se + q;
      float4* const h = sh + q;
      cp_async16(d + 0 * ZQ, rEx + (vN + tvA));
      cp_async16(d + 1 * ZQ, rEz + (vN + tvA));
      cp_async16(d + 2 * ZQ, rEy + (vN + tvA));
      cp_async16(d + 3 * ZQ, rEx + (vN + tvC));
      cp_async16(d + 4 * ZQ, rEz + (vN + tvC));
      cp_async16(h + 0 * ZQ, rHx + (vP + tvA));
      cp_async16(h + 1 * ZQ, rHy + (vP + tvA));
      cp_async16(h + 2 * ZQ, rHz + (vP + tvA));
      if (has_psi) {
        float4* const ps = sh + kLean1HRows * ZQ + slot;
        const unsigned pp = (unsigned)PL * PPn;
        cp_async16(ps, rPx + (pp + pvA));
        cp_async16(ps + psi_row, rPy + (pp + pvA));
      }
      if (se_first) {                              // very first plane of the sweep: E[PL] too
        float4* const f = se_first + q;
        cp_async16(f + 0 * ZQ, rEx + (vP + tvA));
        cp_async16(f + 1 * ZQ, rEz + (vP + tvA));
        cp_async16(f + 2 * ZQ, rEy + (vP + tvA));
        cp_async16(f + 3 * ZQ, rEx + (vP + tvC));
        cp_async16(f + 4 * ZQ, rEz + (vP + tvC));
      }
    };
    // Group "end": B, psiE and the absorber row of plane PL into the single coefficient slot
    // (issued once the E half-step of the previous plane has read the slot).
    auto issue_end = [&](int PL) {
      if (!own) return;
      const unsigned vP = (unsigned)PL * PVn;
      float4* const b = bslot + q;
      cp_async16(b + 0 * ZQ, Bx + (vP + tvA));
      cp_async16(b + 1 * ZQ, By + (vP + tvA));
      cp_async16(b + 2 * ZQ, Bz + (vP + tvA));
      if (lane == 0) cp_async16(bslot + 3 * ZQ + 2 * psi_row, A4 + ((unsigned)PL * (unsigned)Y + yA));
      if (has_psi) {
        float4* const ps = bslot + 3 * ZQ + slot;
        const unsigned pp = (unsigned)PL * PPn;
        cp_async16(ps, ePx + (pp + pvA));
        cp_async16(ps + psi_row, ePy + (pp + pvA));
      }
    };

    int P = wrapi(cstart - 1, X);
    float4* sprev = wbase;                         // E[P]
    float4* scur = wbase + eslot_f4;               // E[P+1]
    float4* snext = wbase + 2 * eslot_f4;          // being filled with E[P+2]
    float4* hcur = hbase;                          // H, psiH of P
    float4* hnext = hbase + hslot_f4;              // ... being filled for P+1
    ok = wait_deps(0);
    if (ok) issue_top(P, P + 1 == X ? 0 : P + 1, scur, hcur, sprev);
    cp_async_commit();                             // group top(0)
    cp_async_commit();                             // group end(-1): empty (iteration 0 has no E half-step)

    float hyp[VW], hzp[VW];
#pragma unroll
    for (int v = 0; v < VW; ++v) { hyp[v] = 0.f; hzp[v] = 0.f; }

    for (int i = 0; i <= X && ok; ++i) {
      const bool real = i >= 1;
      const int Pn = P + 1 == X ? 0 : P + 1;
      const unsigned vP = (unsigned)P * PVn;
      if constexpr (STATS) {
        const long long c0 = clock64();
        cp_async_wait<1>();
        st_cp += clock64() - c0;
      }
      cp_async_wait<1>();                          // group top(i) has landed (end(i-1) may be in flight)
      if (i < X) {
        ok = wait_deps(i + 1);
        if (!ok) break;
        issue_top(Pn, Pn + 1 == X ? 0 : Pn + 1, snext, hnext, nullptr);
      }
      cp_async_commit();                           // group top(i+1)
      if (w == 0 && lane == 0) st_vol_s(&ctl.front, iters_done + (unsigned)i);
      if (cfg.discard && i >= 1 && disc && (q & 7) == 0) {
        const unsigned vN = (unsigned)Pn * PVn;    // (Ex, Ez: after the neighbour's copy, below)
        discard_l2_line(rEy + (vN + tvA));
        discard_l2_line(rHx + (vP + tvA)); discard_l2_line(rHy + (vP + tvA));
        discard_l2_line(rHz + (vP + tvA));
      }
      const unsigned pP = (unsigned)P * PPn;

      // ---------------------------------- H half-step ---------------------------------------------
      const float4* const ep = sprev + q;
      const float4* const ec = scur + q;
      const float4* const hc = hcur + q;
      float ex[VW], ey[VW], ez[VW], hx[VW], hy[VW], hz[VW], psx[VW], psy[VW];
      {
        float exC[VW], ezC[VW], eyx[VW], ezx[VW];
        f4_to_arr(lds16(ep + 0 * ZQ), ex); f4_to_arr(lds16(ep + 1 * ZQ), ez);
        f4_to_arr(lds16(ep + 2 * ZQ), ey);
        f4_to_arr(lds16(ep + 3 * ZQ), exC); f4_to_arr(lds16(ep + 4 * ZQ), ezC);
        f4_to_arr(lds16(ec + 2 * ZQ), eyx); f4_to_arr(lds16(ec + 1 * ZQ), ezx);
        f4_to_arr(lds16(hc + 0 * ZQ), hx); f4_to_arr(lds16(hc + 1 * ZQ), hy);
        f4_to_arr(lds16(hc + 2 * ZQ), hz);
#pragma unroll
        for (int v = 0; v < VW; ++v) { psx[v] = 0.f; psy[v] = 0.f; }
        if (has_psi) {
          const float4* const ps = hcur + kLean1HRows * ZQ + slot;
          f4_to_arr(lds16(ps), psx); f4_to_arr(lds16(ps + psi_row), psy);
        }
        float ex_top = __shfl_down_sync(0xffffffffu, ex[0], 1);
        float ey_top = __shfl_down_sync(0xffffffffu, ey[0], 1);
        if (top) { ex_top = 0.f; ey_top = 0.f; }
#pragma unroll
        for (int v = 0; v < VW; ++v) {
          const float exz = (v + 1 < VW) ? ex[(v + 1) % VW] : ex_top;
          const float eyz = (v + 1 < VW) ? ey[(v + 1) % VW] : ey_top;
          h_cell(ex[v], ey[v], ez[v], exz, eyz, ezC[v], exC[v], eyx[v], ezx[v], ah[v], bh[v],
                 ikh[v], dt, psx[v], psy[v], hx[v], hy[v], hz[v]);
        }
      }
      {
        float4* const xs = xmine + (kk & (kLeanXR - 1)) * 2 * ZQ;
        if (kk >= (unsigned)kLeanXR && w + 1 < NWt) {
          const unsigned need = kk + 1u - (unsigned)kLeanXR;
          const long long c0 = STATS ? clock64() : 0;
          ok = spin([&]() { return ld_vol_s(&ctl.rcnt[w]) >= need; });
          if constexpr (STATS) st_rc += clock64() - c0;
          if (!ok) break;
        }
        xs[0] = arr_to_f4(hz);
        xs[ZQ] = arr_to_f4(hx);
        __syncwarp();
        if (lane == 0) st_vol_s(&ctl.hcnt[w], kk + 1u);
      }

      // ---------------------------------- E half-step ---------------------------------------------
      if (real && own) {
        float hzm[VW], hxm[VW];
        {
          const unsigned need = kk + 1u;
          const long long c0 = STATS ? clock64() : 0;
          ok = spin([&]() { return ld_vol_s(&ctl.hcnt[w - 1]) >= need; });
          if constexpr (STATS) st_hc += clock64() - c0;
          if (!ok) break;
          const float4* const xs = xprev + (kk & (kLeanXR - 1)) * 2 * ZQ;
          f4_to_arr(lds16(xs), hzm); f4_to_arr(lds16(xs + ZQ), hxm);
          __syncwarp();
          if (lane == 0) st_vol_s(&ctl.rcnt[w - 1], kk + 1u);
          if (cfg.discard && disc && (q & 7) == 0) {   // warp w-1's copy of (Ex, Ez)[P+1] has landed
            const unsigned vN = (unsigned)Pn * PVn;
            discard_l2_line(rEx + (vN + tvA)); discard_l2_line(rEz + (vN + tvA));
          }
        }
        float hx_bot = __shfl_up_sync(0xffffffffu, hx[VW - 1], 1);
        float hy_bot = __shfl_up_sync(0xffffffffu, hy[VW - 1], 1);
        if (bottom) { hx_bot = 0.f; hy_bot = 0.f; }
        cp_async_wait<1>();                        // group end(i-1): B, psiE, absorber row of P
        __syncwarp();                              // (the absorber row was copied by lane 0)
        float b0[VW], b1[VW], b2[VW], qsx[VW], qsy[VW];
        const float4* const bc = bslot + q;
        f4_to_arr(lds16(bc + 0 * ZQ), b0); f4_to_arr(lds16(bc + 1 * ZQ), b1);
        f4_to_arr(lds16(bc + 2 * ZQ), b2);
        const float4 aA = lds16(bslot + 3 * ZQ + 2 * psi_row);
#pragma unroll
        for (int v = 0; v < VW; ++v) { qsx[v] = 0.f; qsy[v] = 0.f; }
        if (has_psi) {
          const float4* const ps = bslot + 3 * ZQ + slot;
          f4_to_arr(lds16(ps), qsx); f4_to_arr(lds16(ps + psi_row), qsy);
        }
#pragma unroll
        for (int v = 0; v < VW; ++v) {
          const float hxz = (v > 0) ? hx[(v + VW - 1) % VW] : hx_bot;
          const float hyz = (v > 0) ? hy[(v + VW - 1) % VW] : hy_bot;
          e_cell(hx[v], hy[v], hz[v], hxz, hyz, hzm[v], hxm[v], hyp[v], hzp[v], ae[v], be[v],
                 ike[v], aA.x, aA.y, aA.z, b0[v], b1[v], b2[v], qsx[v], qsy[v], ex[v], ey[v], ez[v]);
        }
        if (g.src_axis == 0 ? (P == sp0 || P == sp1) : srcA)
          add_source<VW>(g, p.src, w0, w1, P, yA, q, ex, ey, ez);
        const unsigned o = vP + tvA;
        __stcg(wHx + o, arr_to_f4(hx)); __stcg(wHy + o, arr_to_f4(hy)); __stcg(wHz + o, arr_to_f4(hz));
        __stcg(wEx + o, arr_to_f4(ex)); __stcg(wEy + o, arr_to_f4(ey)); __stcg(wEz + o, arr_to_f4(ez));
        if (has_psi) {
          __stcg(wPx + (pP + pvA), arr_to_f4(psx)); __stcg(wPy + (pP + pvA), arr_to_f4(psy));
          __stcg(ePx + (pP + pvA), arr_to_f4(qsx)); __stcg(ePy + (pP + pvA), arr_to_f4(qsy));
        }
        if (oi >= 0) write_snapshot<VW>(g, p.out, oi, P, yA, q, ex, ey, ez, p.proj);
      } else if (own && !real && lane == 0) {
        st_vol_s(&ctl.rcnt[w - 1], kk + 1u);       // prologue plane: nothing to consume
      }
      // the coefficient slot has been read: refill it for the next plane
      __syncwarp();
      if (i < X) issue_end(Pn);
      cp_async_commit();                           // group end(i)
      if (real && lane == 0) {
        __threadfence_block();
        st_vol_s(&ctl.wdone[w], base_mine + (unsigned)i);
      }
#pragma unroll
      for (int v = 0; v < VW; ++v) { hyp[v] = hy[v]; hzp[v] = hz[v]; }
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
    if (lane == 0 && (t == 0 || t == NT / 2)) {
      const double tot = (double)(clock64() - st_begin);
      printf("leanstats j %d t %d w %d iters %u cyc/iter %.0f  cp %.3f avail %.3f next %.3f rcnt %.3f hcnt %.3f\n",
             j, t, w, kk, tot / (kk ? kk : 1), st_cp / tot, st_avail / tot, st_next / tot,
             st_rc / tot, st_hc / tot);
    }
  }
  if (lane == 0) atomicAdd(&ctl.exited, 1u);
}

inline size_t lean1_smem_bytes(const Geom& g, int tile_y) {
  const size_t eslot = (size_t)kLean1ERows * 32;
  const size_t hslot = (size_t)kLean1HRows * 32 + 2 * (size_t)g.npg;
  const size_t bslot = 3 * 32 + 2 * (size_t)g.npg + 1;
  const size_t warp_f4 = 3 * eslot + 2 * hslot + bslot + (size_t)kLeanXR * 2 * 32;
  return sizeof(float4) * warp_f4 * (size_t)(tile_y + 1);
}

inline const void* lean1_fn(bool stats) {
  return stats ? (const void*)lean1_kernel<true> : (const void*)lean1_kernel<false>;
}

inline bool lean1_configure(const Geom& g, bool reduced, int tile_y_req, int stages_req, int sms,
                            int l2_bytes, SystolicCfg* cfg, std::string* why) {
  if (reduced) { *why = "fp32 storage only"; return false; }
  if (g.Zq != 32) { *why = "needs a z-column of exactly 32 vectors (125 <= Z <= 128)"; return false; }
  if (g.N / 4 * 3 >= (1ll << 32)) { *why = "domain too large for 32-bit vector indices"; return false; }
  int max_tile = kLean1MaxWarps - 1;             // columns 0..tile form H: tile + 1 warps
  while (max_tile >= 1 && lean1_smem_bytes(g, max_tile) + 512 > 227 * 1024) --max_tile;
  if (max_tile < 1) { *why = "staging ring does not fit in shared memory"; return false; }
  if (tile_y_req > 0 && tile_y_req < max_tile) max_tile = tile_y_req;
  if (max_tile > g.Y) max_tile = g.Y;
  const int ntiles = (g.Y + max_tile - 1) / max_tile;
  const int widest = (g.Y + ntiles - 1) / ntiles;
  cfg->tile_y = widest;
  cfg->ntiles = ntiles;
  cfg->cols = 1;
  cfg->threads = 32 * (widest + 2);
  cfg->need_zfix = 0;
  cfg->unroll = 1;
  cfg->smem_bytes = (int)lean1_smem_bytes(g, widest);
  cfg->max_lead = 10;
  cfg->pf_ahead = 6;
  cfg->svc_sleep_ns = 200;
  cfg->spin_ns_max = 160;
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
  const void* fn = lean1_fn(getenv("B200FDTD_LEAN_STATS") != nullptr);
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
  const long long plane_bytes = g.P * 4ll * 15;
  const int lag = 6;
  long long by_l2 = (long long)(l2_bytes * 0.8) / (lag * plane_bytes);
  if (by_l2 < 1) by_l2 = 1;
  if (stages > by_l2) stages = (int)by_l2;
  if (stages_req > 0 && stages_req <= capacity / ntiles) stages = stages_req;
  if (stages > g.tt) stages = g.tt > 0 ? g.tt : 1;
  if (stages > g.X) stages = g.X;
  cfg->stages = stages;
  cfg->l2_window_bytes = (long long)stages * lag * plane_bytes;
  return true;
}

inline int lean1_launch(const Geom& g, const Ptrs<float>& p, const SystolicCfg& cfg, unsigned* sync,
                        cudaStream_t st) {
  const void* fn = lean1_fn(getenv("B200FDTD_LEAN_STATS") != nullptr);
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       cfg.smem_bytes);
  if (e != cudaSuccess) return (int)e;
  Geom gg = g;
  Ptrs<float> pp = p;
  SystolicCfg cc = cfg;
  void* args[] = {&gg, &pp, &cc, &sync};
  e = cudaLaunchCooperativeKernel(fn, dim3(cfg.stages * cfg.ntiles), dim3(cfg.threads), args,
                                  cfg.smem_bytes, st);
  if (e != cudaSuccess) return (int)e;
  systolic_check_kernel<<<1, 1, 0, st>>>(sync + (size_t)cfg.stages * cfg.ntiles * kSysFlagStride);
  return (int)cudaGetLastError();
}

}  // namespace b200
