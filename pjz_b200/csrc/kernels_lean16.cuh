// "systolic_lean", sub-warp generation: the warp-autonomous kernel of kernels_lean.cuh (read that
// header first: same stage/tile decomposition, ping-pong buffers, progress protocol, service warp
// and L2 discard rule) for z-columns of AT MOST 16 16-byte vectors -- everything fdtd-z itself
// accepts: fp16 storage with Z <= 128 (pjz's default: use_reduced_precision=True,
// Z = 128 - sum(pml_widths) = 96, /root/reference/src/pjz/_field.py:36,56-58) and fp32 storage
// with Z <= 64.
//
//  * A z-column occupies LPC = 16 or 8 lanes (one 16-byte vector per lane), so a warp holds
//    G = 32 / LPC sub-groups; a thread owns K adjacent columns (K = 1 for fp16 storage, where a
//    vector is 8 cells; K = 2 for fp32, 4 cells per vector): 8 cells per thread and plane in both
//    cases, the same work per thread as the fp32 kernel of kernels_lean.cuh.  A warp therefore
//    owns CW = G * K adjacent columns, tile-local CW*w .. CW*w + CW - 1 (column 0 = y0-1 halo).
//  * z+-1 neighbours are shuffles of width LPC; x-1 is carried in registers along the sweep; the
//    y neighbour inside a thread's K columns is its own registers.
//  * Each lane stages the operands of ITS columns (E^n[P+1], H^{n-1/2}[P], B[P], psi[P]) into the
//    per-warp cp.async ring.  The y+1 neighbour of a sub-group's last column is the next
//    sub-group's ring row (read after cp.async.wait_group + __syncwarp); the column after the
//    warp's last one is copied once, Ex by sub-group 0 and Ez by sub-group 1.
//  * The y-1 neighbour of a sub-group's first column (new Hz, Hx, already rounded to the storage
//    type) goes through the 2-deep boundary slot: sub-group h reads what sub-group h-1 wrote
//    (program order + __syncwarp); sub-group 0 reads the last sub-group of warp w-1
//    (produced / consumed counters, as in kernels_lean.cuh).
//  * Lanes q >= Zq (columns shorter than LPC vectors) run along on the addresses of lane Zq-1
//    and store nothing; columns beyond the tile's E-only column Yt+1 are clamped onto it.
//  * CPML tables: registers for fp32 (24 values), shared memory for fp16 (48 values would not fit
//    next to the 8-cell working set under the 168-register cap of 12 warps).
//  * L2 discard needs whole 128-byte lines per column (Zq a multiple of 8) and is opt-in here
//    (B200FDTD_LEAN_DISCARD=1): it halves HBM traffic but costs 4-5 % throughput at these sizes.
#pragma once

#include "kernels_lean.cuh"

namespace b200 {

constexpr int kL16ZR = 16;             // most 16-byte vectors per z-column this kernel takes

template <typename T, int LPC, int K>
struct L16 {
  static constexpr int VW = VecTraits<T>::VW;
  static constexpr int PV = VW / 4;              // float4 per psi / table vector
  static constexpr int G = 32 / LPC;             // sub-groups (column slices) per warp
  static constexpr int CW = G * K;               // columns per warp
  static constexpr int ERows = 3 * CW + 2;       // Ex[0..CW], Ez[0..CW], Ey[0..CW-1]
  static constexpr int HRows = 6 * CW;           // Hx, Hy, Hz, Bx, By, Bz [0..CW-1]
  static constexpr int XF4 = G * 2 * LPC;        // float4 per boundary-slot ring entry (= 64)
  static constexpr int MaxWarps = K == 2 ? 8 : 11;   // compute warps per CTA (+1 service warp): 12 warps = 168 registers; a 13th rounds up to 16 warps = 128
  static size_t eslot_f4() { return (size_t)ERows * LPC; }
  static size_t hslot_f4(int npg) { return (size_t)HRows * LPC + 4 * CW * (size_t)npg * PV + 2 * CW; }
  static size_t warp_f4(int npg) { return 3 * eslot_f4() + 2 * hslot_f4(npg) + (size_t)kLeanXR * XF4; }
};

struct Lean16Ctl {
  unsigned avail, next, ok, front, exited;   // as LeanCtl
  unsigned wdone[12];
  unsigned hcnt[12];
  unsigned rcnt[12];
};

template <typename T, int LPC, int K, bool STATS = false>
__global__ void __launch_bounds__(32 * (L16<T, LPC, K>::MaxWarps + 1), 1)
lean16_kernel(const Geom g, const Ptrs<T> p, const SystolicCfg cfg, unsigned* sync) {
  using C = L16<T, LPC, K>;
  constexpr int VW = C::VW, PV = C::PV, G = C::G, CW = C::CW;
  extern __shared__ float4 smem[];
  __shared__ Lean16Ctl ctl;
  __shared__ float4 stab[6][PV * LPC];             // CPML tables, [table][part * LPC + lane]
  const int tid = threadIdx.x, lane = tid & 31;
  const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int NW = (int)(blockDim.x >> 5) - 1;       // compute warps
  const int S = cfg.stages, NT = cfg.ntiles;
  const bool stage_major = cfg.block_order == 2;        // block order, see kernels_lean.cuh
  const int t = stage_major ? (int)blockIdx.x / S : (int)(blockIdx.x % NT);
  const int j = stage_major ? (int)blockIdx.x % S : (int)(blockIdx.x / NT);
  const int y0 = (int)((long long)t * g.Y / NT);
  const int Yt = (int)((long long)(t + 1) * g.Y / NT) - y0;
  const int X = g.X, Y = g.Y, Zq = g.Zq, npg = g.npg;
  const int psi_row = npg * PV;                    // float4 per psi row of a slot
  const int eslot_f4 = C::ERows * LPC;
  const int hslot_f4 = C::HRows * LPC + 4 * CW * psi_row + 2 * CW;   // + psi rows, absorber, z-source rows
  const int warp_f4 = 3 * eslot_f4 + 2 * hslot_f4 + kLeanXR * C::XF4;
  const int NWt = min(NW, (Yt + CW) / CW);         // warps with an H-forming column on THIS tile

  unsigned* const status = sync + (size_t)S * NT * kSysFlagStride;
  unsigned* const my_prog = sync + ((size_t)j * NT + t) * kSysFlagStride;

  if (tid < (int)(sizeof(Lean16Ctl) / sizeof(unsigned))) reinterpret_cast<unsigned*>(&ctl)[tid] = 0u;
  for (int i = tid; i < 6 * PV * LPC; i += (int)blockDim.x) {
    const int k = i / (PV * LPC), r = i % (PV * LPC), part = r / LPC, ql = r % LPC;
    stab[k][r] = ql < Zq ? __ldg(reinterpret_cast<const float4*>(p.tab + (size_t)k * g.Zp) + ql * PV + part)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
  }
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
    // L2 prefetch duty: lanes 8..16 own one array each (E0..2, H0..2 of the read set, B0..2).
    const int ylo = max(y0 - 1, 0), yhi = min(y0 + Yt, Y - 1);
    const unsigned pf_bytes = (unsigned)((yhi - ylo + 1) * g.Zp * (int)sizeof(T));
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
          // only planes the previous step has already produced (see kernels_lean.cuh)
          if (n > g.n0) {
            const unsigned m = (unsigned)((n - g.n0) / S);
            const unsigned need = (j > 0 ? m : m - 1u) * (unsigned)X + (unsigned)min((int)it + 2, X);
            if (min(v0, min(v1, v2)) < need) break;
          }
          const int rb = n & 1;
          const int P = wrapi(n % X - 1 + (int)it, X), Pn = wrapi(P + 1, X);
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
  if (w >= NWt) {                                  // narrower tile: this warp has no column
    if (lane == 0) atomicAdd(&ctl.exited, 1u);
    return;
  }
  const int h = lane / LPC, q = lane % LPC;        // sub-group (column slice), z-vector
  const bool live = q < Zq;
  const int qc = live ? q : Zq - 1;                // addresses of a lane beyond the column's top
  const int slot = psi_slot(g, qc);
  const bool has_psi = live && slot >= 0;
  const unsigned PVn = (unsigned)Y * Zq;           // vectors per x-plane
  const unsigned PPn = (unsigned)Y * psi_row;      // psi float4 per x-plane
  const int sp0 = g.src_pos, sp1 = wrapi(g.src_pos - 1, g.src_axis == 0 ? X : Y);
  int yk[K];                                       // tile-local column CW*w + K*h + k; 0 = y0-1 halo
  unsigned tv[K], pv[K];
  bool own[K], disc[K], srcY[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int c = CW * w + K * h + k;
    yk[k] = wrapi(y0 - 1 + min(c, Yt + 1), Y);
    own[k] = live && c >= 1 && c <= Yt;            // E-updated and stored by this lane
    disc[k] = cfg.discard && live && (q & 7) == 0 && c >= 2 && c <= Yt - 1;
    srcY[k] = g.src_axis == 1 && (yk[k] == sp0 || yk[k] == sp1);
    tv[k] = (unsigned)yk[k] * Zq + qc;
    pv[k] = ((unsigned)yk[k] * npg + (has_psi ? slot : 0)) * PV;
  }
  const int yC = wrapi(y0 - 1 + min(CW * w + CW, Yt + 1), Y);     // the column after the warp's last
  const unsigned tvC = (unsigned)yC * Zq + qc;
  const int ya = (K == 2 && (q & 1)) ? yk[K - 1] : yk[0];   // column whose absorber / z-source row lane q copies
  const bool top = q + 1 >= Zq, bottom = q == 0;

  float4* const wbase = smem + (size_t)w * warp_f4;
  float4* const hbase = wbase + 3 * eslot_f4;
  float4* const xbase = hbase + 2 * hslot_f4;                     // boundary-H slots of this warp
  float4* const xmine = xbase + h * 2 * LPC + q;
  // (Hz, Hx) of the column before this sub-group's first: sub-group h-1 of this warp, or the last
  // sub-group of warp w-1 (column 0 has no predecessor and is never E-updated)
  const float4* const xread = h > 0 ? xbase + (h - 1) * 2 * LPC + q
                                    : (w > 0 ? xbase - warp_f4 + (G - 1) * 2 * LPC + q : xbase + q);

  // CPML tables: registers when they are 4 values each
  float tabr[6][4];
  if constexpr (VW == 4) {
#pragma unroll
    for (int k = 0; k < 6; ++k) f4_to_arr(stab[k][q], tabr[k]);
  }
  auto load_tab = [&](int k, float (&v)[VW]) {
    if constexpr (VW == 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = tabr[k][i];
    } else {
#pragma unroll
      for (int i = 0; i < PV; ++i) {
        const float4 r = stab[k][i * LPC + q];
        v[4 * i] = r.x; v[4 * i + 1] = r.y; v[4 * i + 2] = r.z; v[4 * i + 3] = r.w;
      }
    }
  };
  // psi vectors of a column in a ring row: part i of PML group `slot` at [i * npg + slot]
  auto load_psi = [&](const float4* ps, float (&v)[VW]) {
#pragma unroll
    for (int i = 0; i < PV; ++i) {
      const float4 r = ps[i * npg];
      v[4 * i] = r.x; v[4 * i + 1] = r.y; v[4 * i + 2] = r.z; v[4 * i + 3] = r.w;
    }
  };
  auto store_psi = [&](float4* dst, const float (&v)[VW]) {
#pragma unroll
    for (int i = 0; i < PV; ++i)
      __stcg(dst + i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
  };

  const float dt = g.dt;
  const float4* const A4 = reinterpret_cast<const float4*>(p.A4);
  const float4* const S4 = reinterpret_cast<const float4*>(p.S4);
  const bool zsrc = g.src_axis == 2;
  const bool zhit = zsrc && q == g.src_pos / VW;
  const int zidx = g.src_pos % VW;

  bool ok = true;
  unsigned kk = 0;                                 // cumulative iteration count (never reset)
  unsigned iters_done = 0;
  long long st_cp = 0, st_avail = 0, st_next = 0, st_rc = 0, st_hc = 0;   // STATS: cycles waited
  const long long st_begin = STATS ? clock64() : 0;

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
    const float4* const rEx = reinterpret_cast<const float4*>(p.Es[rb][0]);
    const float4* const rEy = reinterpret_cast<const float4*>(p.Es[rb][1]);
    const float4* const rEz = reinterpret_cast<const float4*>(p.Es[rb][2]);
    const float4* const rEc = h ? rEz : rEx;       // this lane's share of the column after the warp's
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

    // the k+3 rule and the max_lead throttle, exactly as in kernels_lean.cuh
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

    // Async copies of one iteration (plane PL): E[PL+1] -> E slot se; H, B, psi, absorber row of
    // PL -> H/B slot sh.  Straight-line: halo / E-only / clamped columns are copied all the same.
    auto issue_e = [&](int PL, int PLn, float4* se, float4* se_first) {
      const unsigned vN = (unsigned)PLn * PVn, vP = (unsigned)PL * PVn;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float4* const d = se + (K * h + k) * LPC + q;
        cp_async16(d, rEx + (vN + tv[k]));
        cp_async16(d + (CW + 1) * LPC, rEz + (vN + tv[k]));
        cp_async16(d + 2 * (CW + 1) * LPC, rEy + (vN + tv[k]));
      }
      if (G == 2 || h < 2) cp_async16(se + (h ? 2 * CW + 1 : CW) * LPC + q, rEc + (vN + tvC));
      if (se_first) {                                // very first plane of the sweep: E[PL] too
#pragma unroll
        for (int k = 0; k < K; ++k) {
          float4* const f = se_first + (K * h + k) * LPC + q;
          cp_async16(f, rEx + (vP + tv[k]));
          cp_async16(f + (CW + 1) * LPC, rEz + (vP + tv[k]));
          cp_async16(f + 2 * (CW + 1) * LPC, rEy + (vP + tv[k]));
        }
        if (G == 2 || h < 2) cp_async16(se_first + (h ? 2 * CW + 1 : CW) * LPC + q, rEc + (vP + tvC));
      }
    };
    // H, B, psi, absorber / z-source rows: issued after the H half-step (the slot they fill is
    // idle for the whole iteration; two shorter bursts instead of one, see kernels_lean.cuh)
    auto issue_h = [&](int PL, float4* sh) {
      const unsigned vP = (unsigned)PL * PVn;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float4* const hh = sh + (K * h + k) * LPC + q;
        cp_async16(hh + 0 * CW * LPC, rHx + (vP + tv[k]));
        cp_async16(hh + 1 * CW * LPC, rHy + (vP + tv[k]));
        cp_async16(hh + 2 * CW * LPC, rHz + (vP + tv[k]));
        cp_async16(hh + 3 * CW * LPC, Bx + (vP + tv[k]));
        cp_async16(hh + 4 * CW * LPC, By + (vP + tv[k]));
        cp_async16(hh + 5 * CW * LPC, Bz + (vP + tv[k]));
      }
      float4* const tail = sh + C::HRows * LPC + 4 * CW * psi_row;
      if (q < K) cp_async16(tail + K * h + q, A4 + ((unsigned)PL * (unsigned)Y + ya));
      if (zsrc && q >= K && q < 2 * K)
        cp_async16(tail + CW + K * h + (q - K), S4 + ((unsigned)PL * (unsigned)Y + ya));
      if (has_psi) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          float4* const ps = sh + C::HRows * LPC + (K * h + k) * psi_row + slot;
          const unsigned pp = (unsigned)PL * PPn + pv[k];
#pragma unroll
          for (int i = 0; i < PV; ++i) {
            cp_async16(ps + 0 * CW * psi_row + i * npg, rPx + (pp + i));
            cp_async16(ps + 1 * CW * psi_row + i * npg, rPy + (pp + i));
            cp_async16(ps + 2 * CW * psi_row + i * npg, ePx + (pp + i));
            cp_async16(ps + 3 * CW * psi_row + i * npg, ePy + (pp + i));
          }
        }
      }
    };
    auto issue = [&](int PL, int PLn, float4* se, float4* sh, float4* se_first) {
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
    __syncwarp();                                  // the previous sweep's reads of the ring are done
    if (ok) issue(P, P + 1 == X ? 0 : P + 1, scur, hcur, sprev);
    cp_async_commit();

    float hyp[K][VW], hzp[K][VW];                  // H^{n+1/2}[P-1] of the thread's own cells
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int v = 0; v < VW; ++v) { hyp[k][v] = 0.f; hzp[k][v] = 0.f; }

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
      __syncwarp();                                // ... and so have those of the other lanes
      if (w == 0 && lane == 0) st_vol_s(&ctl.front, iters_done + (unsigned)i);
      // Dead lines (see kernels_lean.cuh): Ey and H of every column and (Ex, Ez) of every column
      // but the warp's first have been read by their only reader; (Ex, Ez) of the warp's first
      // column wait for warp w-1 (E half-step).
      if (i >= 1) {
        const unsigned vN = (unsigned)Pn * PVn;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          if (disc[k]) {
            discard_l2_line(rEy + (vN + tv[k]));
            discard_l2_line(rHx + (vP + tv[k])); discard_l2_line(rHy + (vP + tv[k]));
            discard_l2_line(rHz + (vP + tv[k]));
            if (k > 0 || h > 0) { discard_l2_line(rEx + (vN + tv[k])); discard_l2_line(rEz + (vN + tv[k])); }
          }
        }
      }

      // ---------------------------------- H half-step ---------------------------------------------
      const float4* const ep = sprev + K * h * LPC + q;
      const float4* const ec = scur + K * h * LPC + q;
      const float4* const hc = hcur + K * h * LPC + q;
      const float4* const pc = hcur + C::HRows * LPC + K * h * psi_row + (has_psi ? slot : 0);
      const float4* const tail = hcur + C::HRows * LPC + 4 * CW * psi_row;
      float ex[K][VW], ey[K][VW], ez[K][VW], hx[K][VW], hy[K][VW], hz[K][VW], psx[K][VW], psy[K][VW];
      float4 phx[K], phy[K], phz[K];               // the new H in storage format (boundary slot, stores)
      {
        float exn[VW], ezn[VW], ah[VW], bh[VW], ikh[VW];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          unpack(lds16(ep + k * LPC), ex[k], T());
          unpack(lds16(ep + (CW + 1 + k) * LPC), ez[k], T());
          unpack(lds16(ep + (2 * (CW + 1) + k) * LPC), ey[k], T());
          unpack(lds16(hc + (0 * CW + k) * LPC), hx[k], T());
          unpack(lds16(hc + (1 * CW + k) * LPC), hy[k], T());
          unpack(lds16(hc + (2 * CW + k) * LPC), hz[k], T());
#pragma unroll
          for (int v = 0; v < VW; ++v) { psx[k][v] = 0.f; psy[k][v] = 0.f; }
          if (has_psi) {
            load_psi(pc + (0 * CW + k) * psi_row, psx[k]);
            load_psi(pc + (1 * CW + k) * psi_row, psy[k]);
          }
        }
        unpack(lds16(ep + K * LPC), exn, T());                 // the column after this thread's last
        unpack(lds16(ep + (CW + 1 + K) * LPC), ezn, T());
        load_tab(3, ah); load_tab(4, bh); load_tab(5, ikh);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          float eyx[VW], ezx[VW];
          unpack(lds16(ec + (2 * (CW + 1) + k) * LPC), eyx, T());
          unpack(lds16(ec + (CW + 1 + k) * LPC), ezx, T());
          float ex_top = __shfl_down_sync(0xffffffffu, ex[k][0], 1, LPC);
          float ey_top = __shfl_down_sync(0xffffffffu, ey[k][0], 1, LPC);
          if (top) { ex_top = 0.f; ey_top = 0.f; }
#pragma unroll
          for (int v = 0; v < VW; ++v) {
            const float exz = (v + 1 < VW) ? ex[k][(v + 1) % VW] : ex_top;
            const float eyz = (v + 1 < VW) ? ey[k][(v + 1) % VW] : ey_top;
            const float ez_yp = (k + 1 < K) ? ez[(k + 1) % K][v] : ezn[v];
            const float ex_yp = (k + 1 < K) ? ex[(k + 1) % K][v] : exn[v];
            h_cell(ex[k][v], ey[k][v], ez[k][v], exz, eyz, ez_yp, ex_yp, eyx[v], ezx[v], ah[v], bh[v],
                   ikh[v], dt, psx[k][v], psy[k][v], hx[k][v], hy[k][v], hz[k][v]);
          }
          // round to the storage type by packing (one F2FP per pair of halves) and keep the packed
          // vectors for the boundary slot and the stores; the E half-step uses the rounded values
          phx[k] = pack(hx[k], T()); phy[k] = pack(hy[k], T()); phz[k] = pack(hz[k], T());
          if constexpr (sizeof(T) == 2) {
            unpack(phx[k], hx[k], T()); unpack(phy[k], hy[k], T()); unpack(phz[k], hz[k], T());
          }
        }
      }
      if (i < X) issue_h(Pn, hnext);               // second half of the copies for iteration i + 1
      cp_async_commit();
      // boundary H for the next column: the last sub-group waits until warp w+1 has consumed the slot
      {
        float4* const xs = xmine + (kk & (kLeanXR - 1)) * C::XF4;
        if (kk >= (unsigned)kLeanXR && w + 1 < NWt) {
          const unsigned need = kk + 1u - (unsigned)kLeanXR;
          const long long c0 = STATS ? clock64() : 0;
          ok = spin([&]() { return ld_vol_s(&ctl.rcnt[w]) >= need; });
          if constexpr (STATS) st_rc += clock64() - c0;
          if (!ok) break;
        }
        xs[0] = phz[K - 1];
        xs[LPC] = phx[K - 1];
        __syncwarp();
        if (lane == 0) st_vol_s(&ctl.hcnt[w], kk + 1u);
      }

      // ---------------------------------- E half-step ---------------------------------------------
      if (real) {
        if (w > 0) {
          const unsigned need = kk + 1u;
          const long long c0 = STATS ? clock64() : 0;
          ok = spin([&]() { return ld_vol_s(&ctl.hcnt[w - 1]) >= need; });
          if constexpr (STATS) st_hc += clock64() - c0;
          if (!ok) break;
        }
        float hzm[VW], hxm[VW];                    // (Hz, Hx) of the column before this thread's first
        {
          const float4* const xr = xread + (kk & (kLeanXR - 1)) * C::XF4;
          unpack(lds16(xr), hzm, T()); unpack(lds16(xr + LPC), hxm, T());
        }
        __syncwarp();
        if (w > 0) {
          if (lane == 0) st_vol_s(&ctl.rcnt[w - 1], kk + 1u);
          // warp w-1 has finished the H half-step of this iteration: its copy of (Ex, Ez)[P+1] of
          // this warp's first column -- the only other reader -- has landed
          if (disc[0] && h == 0) {
            const unsigned vN = (unsigned)Pn * PVn;
            discard_l2_line(rEx + (vN + tv[0])); discard_l2_line(rEz + (vN + tv[0]));
          }
        }
        float ae[VW], be[VW], ike[VW];
        load_tab(0, ae); load_tab(1, be); load_tab(2, ike);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          float hx_bot = __shfl_up_sync(0xffffffffu, hx[k][VW - 1], 1, LPC);
          float hy_bot = __shfl_up_sync(0xffffffffu, hy[k][VW - 1], 1, LPC);
          if (bottom) { hx_bot = 0.f; hy_bot = 0.f; }
          if (own[k]) {
            float b0[VW], b1[VW], b2[VW], qsx[VW], qsy[VW];
            unpack(lds16(hc + (3 * CW + k) * LPC), b0, T()); unpack(lds16(hc + (4 * CW + k) * LPC), b1, T());
            unpack(lds16(hc + (5 * CW + k) * LPC), b2, T());
            const float4 aa = lds16(tail + K * h + k);
#pragma unroll
            for (int v = 0; v < VW; ++v) { qsx[v] = 0.f; qsy[v] = 0.f; }
            if (has_psi) {
              load_psi(pc + (2 * CW + k) * psi_row, qsx);
              load_psi(pc + (3 * CW + k) * psi_row, qsy);
            }
#pragma unroll
            for (int v = 0; v < VW; ++v) {
              const float hxz = (v > 0) ? hx[k][(v + VW - 1) % VW] : hx_bot;
              const float hyz = (v > 0) ? hy[k][(v + VW - 1) % VW] : hy_bot;
              const float hz_ym = (k > 0) ? hz[(k + K - 1) % K][v] : hzm[v];
              const float hx_ym = (k > 0) ? hx[(k + K - 1) % K][v] : hxm[v];
              e_cell(hx[k][v], hy[k][v], hz[k][v], hxz, hyz, hz_ym, hx_ym, hyp[k][v], hzp[k][v], ae[v],
                     be[v], ike[v], aa.x, aa.y, aa.z, b0[v], b1[v], b2[v], qsx[v], qsy[v], ex[k][v],
                     ey[k][v], ez[k][v]);
            }
            if (zsrc) {
              // z-plane source: same operation order as add_source() (channel 0, then 1)
              const float4 s = lds16(tail + CW + K * h + k);
#pragma unroll
              for (int v = 0; v < VW; ++v) {
                const float t0 = fmaf(w1, s.z, fmaf(w0, s.x, ex[k][v]));
                const float t1 = fmaf(w1, s.w, fmaf(w0, s.y, ey[k][v]));
                const bool hit = zhit && v == zidx;
                ex[k][v] = hit ? t0 : ex[k][v];
                ey[k][v] = hit ? t1 : ey[k][v];
              }
            } else if (g.src_axis == 0 ? (P == sp0 || P == sp1) : srcY[k]) {
              add_source<VW>(g, p.src, w0, w1, P, yk[k], q, ex[k], ey[k], ez[k]);
            }
            const unsigned o = vP + tv[k];
            __stcg(wHx + o, phx[k]); __stcg(wHy + o, phy[k]); __stcg(wHz + o, phz[k]);
            __stcg(wEx + o, pack(ex[k], T())); __stcg(wEy + o, pack(ey[k], T()));
            __stcg(wEz + o, pack(ez[k], T()));
            if (has_psi) {
              const unsigned pP = (unsigned)P * PPn + pv[k];
              store_psi(wPx + pP, psx[k]); store_psi(wPy + pP, psy[k]);
              store_psi(ePx + pP, qsx); store_psi(ePy + pP, qsy);
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
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int v = 0; v < VW; ++v) { hyp[k][v] = hy[k][v]; hzp[k][v] = hz[k][v]; }
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
      printf("lean16stats j %d t %d w %d iters %u cyc/iter %.0f  cp %.3f avail %.3f next %.3f rcnt %.3f hcnt %.3f\n",
             j, t, w, kk, tot / (kk ? kk : 1), st_cp / tot, st_avail / tot, st_next / tot,
             st_rc / tot, st_hc / tot);
    }
  }
  if (lane == 0) atomicAdd(&ctl.exited, 1u);
}

// Which instance serves a geometry: K = 2 columns per thread for fp32 (4 cells per vector), 1 for
// fp16 (8 cells per vector); 8 lanes per column when the column has at most 8 vectors.
template <typename T> constexpr int l16_k() { return sizeof(T) == 4 ? 2 : 1; }

// B200FDTD_LEAN_STATS=1: the per-warp wait accounting build (prints at the end of the launch).
template <typename T, int LPC>
inline const void* lean16_fn() {
  if (getenv("B200FDTD_LEAN_STATS") != nullptr)
    return (const void*)lean16_kernel<T, LPC, l16_k<T>(), true>;
  return (const void*)lean16_kernel<T, LPC, l16_k<T>(), false>;
}

template <typename T, int LPC>
inline bool lean16_configure_i(const Geom& g, int tile_y_req, int stages_req, int sms, int l2_bytes,
                               SystolicCfg* cfg, std::string* why) {
  using C = L16<T, LPC, l16_k<T>()>;
  if (g.N / C::VW * 3 >= (1ll << 32)) { *why = "domain too large for 32-bit vector indices"; return false; }
  const size_t warp_bytes = sizeof(float4) * C::warp_f4(g.npg);
  int warps = C::MaxWarps;
  while (warps >= 1 && warp_bytes * warps + 4096 > 227 * 1024) --warps;
  if (warps < 1) { *why = "staging ring does not fit in shared memory"; return false; }
  int max_tile = C::CW * warps - 1;              // CW*NW - 1 owned columns fill NW warps exactly
  if (tile_y_req > 0 && tile_y_req < max_tile) max_tile = tile_y_req;
  if (max_tile > g.Y) max_tile = g.Y;
  const int ntiles = (g.Y + max_tile - 1) / max_tile;
  const int widest = (g.Y + ntiles - 1) / ntiles;
  const int nw = (widest + C::CW) / C::CW;
  cfg->tile_y = widest;
  cfg->ntiles = ntiles;
  cfg->cols = 16;                                // marks the sub-warp variant for the launcher
  cfg->threads = 32 * (nw + 1);
  cfg->smem_bytes = (int)(warp_bytes * nw);
  cfg->max_lead = 10;
  cfg->pf_ahead = 0;                             // see lean_configure: 126.2 vs 122.2 Gcell/s (fp16 cfg2)
  cfg->svc_sleep_ns = 200;
  cfg->spin_ns_max = 400;
  // L2 discard of consumed lines: off by default here.  With these short columns (and half the
  // bytes per cell in fp16) HBM is nowhere near its limit, and the CCTL instructions cost 4-5 %
  // (256x256x128 fp16: 106.3 vs 102.4 Gcell/s; fp32 Z=64: 80.1 vs 76.3).  B200FDTD_LEAN_DISCARD=1
  // turns it on where a column is whole 128-byte lines (halves the HBM traffic, lower power).
  cfg->discard = 0;
  if (const char* e = getenv("B200FDTD_LEAN_DISCARD")) cfg->discard = (g.Zq % 8 == 0) && atoi(e) != 0;
  if (const char* e = getenv("B200FDTD_SPIN_NS")) cfg->spin_ns_max = atoi(e);
  if (const char* e = getenv("B200FDTD_MAX_LEAD")) cfg->max_lead = atoi(e);
  if (const char* e = getenv("B200FDTD_PF_AHEAD")) cfg->pf_ahead = atoi(e);
  if (const char* e = getenv("B200FDTD_SVC_SLEEP")) cfg->svc_sleep_ns = atoi(e);
  if (cfg->max_lead < 6) cfg->max_lead = 6;
  if (cfg->pf_ahead < 0) cfg->pf_ahead = 0;
  cfg->trap_on_timeout = 1;
  // back-off of a waiting warp: first sleep 200 ns, doubling up to 400 (round 2: 20 -> 160 before;
  // cfg2 96.2 -> 97.2 Gcell/s on a slower box, fp16 126.5 -> 129.0: half the polls, less power)
  cfg->spin_ns0 = 200;
  if (const char* e = getenv("B200FDTD_SPIN_NS0")) cfg->spin_ns0 = atoi(e) < 1 ? 1 : atoi(e);
  cfg->block_order = 2;                          // stage-major; fp16 cfg2 130.7 -> 131.1
  if (const char* e = getenv("B200FDTD_LEAN_MAP")) cfg->block_order = atoi(e) != 0 ? 2 : 1;
  int occ = 0;
  const void* fn = lean16_fn<T, LPC>();
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
  const int lag = 6;                             // planes a stage trails its predecessor by
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

template <typename T>
inline bool lean16_configure(const Geom& g, int tile_y_req, int stages_req, int sms, int l2_bytes,
                             SystolicCfg* cfg, std::string* why) {
  if (g.Zq > kL16ZR) { *why = "needs a z-column of at most 16 vectors"; return false; }
  return g.Zq <= 8 ? lean16_configure_i<T, 8>(g, tile_y_req, stages_req, sms, l2_bytes, cfg, why)
                   : lean16_configure_i<T, 16>(g, tile_y_req, stages_req, sms, l2_bytes, cfg, why);
}

template <typename T>
inline int lean16_launch(const Geom& g, const Ptrs<T>& p, const SystolicCfg& cfg, unsigned* sync,
                         cudaStream_t st) {
  const void* fn = g.Zq <= 8 ? lean16_fn<T, 8>() : lean16_fn<T, 16>();
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       cfg.smem_bytes);
  if (e != cudaSuccess) return (int)e;
  Geom gg = g;
  Ptrs<T> pp = p;
  SystolicCfg cc = cfg;
  void* args[] = {&gg, &pp, &cc, &sync};
  e = cudaLaunchCooperativeKernel(fn, dim3(cfg.stages * cfg.ntiles), dim3(cfg.threads), args,
                                  cfg.smem_bytes, st);
  if (e != cudaSuccess) return (int)e;
  systolic_check_kernel<<<1, 1, 0, st>>>(sync + (size_t)cfg.stages * cfg.ntiles * kSysFlagStride);
  return (int)cudaGetLastError();
}

}  // namespace b200
