// "systolic_lean", half-warp generation: the warp-autonomous kernel of kernels_lean.cuh (read
// that header first: same stage/tile decomposition, ping-pong buffers, progress protocol, service
// warp and L2 discard rule) for z-columns of AT MOST 16 16-byte vectors -- fp16 storage with
// Z <= 128 (pjz's default: use_reduced_precision=True, Z = 128 - sum(pml_widths) = 96,
// /root/reference/src/pjz/_field.py:36,56-58) and fp32 storage with Z <= 64.
//
//  * A warp still owns one PAIR of adjacent y-columns (tile-local 2w, 2w+1), but a thread owns
//    ONE column: lanes 0..15 hold z-vectors 0..15 of column A, lanes 16..31 those of column B.
//    With fp16 storage a vector is 8 cells, so a thread advances 8 cells per plane -- the same
//    work per thread as the fp32 kernel's 2 columns x 4 cells -- from half the shared-memory
//    reads and half the cp.async copies.
//  * z+-1 neighbours are shuffles of width 16; x-1 is carried in registers along the sweep.
//  * Each lane stages the operands of ITS column (E^n[P+1], H^{n-1/2}[P], B[P], psi[P]) into the
//    per-warp cp.async ring.  The y+1 neighbour of column A is column B's ring row (read after
//    cp.async.wait_group + __syncwarp); the column after the pair is copied once, Ex by the A
//    lanes and Ez by the B lanes (one copy instruction per lane).
//  * The y-1 neighbour (new Hz, Hx, already rounded to the storage type) goes through the 2-deep
//    boundary slot: the A half of a warp's slot is read by its own B lanes (program order +
//    __syncwarp), the B half by the A lanes of warp w+1 (produced / consumed counters).
//  * Lanes q >= Zq (columns shorter than 16 vectors) run along on the addresses of lane Zq-1 and
//    store nothing.
//  * CPML tables: registers for fp32 (24 values), shared memory for fp16 (48 values would not fit
//    next to the 8-cell working set under the 168-register cap of 12 warps).
//  * L2 discard needs whole 128-byte lines per column: only when Zq is a multiple of 8.
#pragma once

#include "kernels_lean.cuh"

namespace b200 {

constexpr int kL16MaxWarps = 11;       // compute warps per CTA (+1 service warp: 384 threads)
constexpr int kL16ZR = 16;             // lanes (16-byte vectors) per z-column / ring row
constexpr int kL16ERows = 8;           // rows per E slot: Ex[A,B,C], Ez[A,B,C], Ey[A,B]
constexpr int kL16HRows = 12;          // rows per H/B slot: Hx,Hy,Hz[A,B], Bx,By,Bz[A,B]

struct Lean16Ctl {
  unsigned avail, next, ok, front, exited;   // as LeanCtl
  unsigned wdone[kL16MaxWarps + 1];
  unsigned hcnt[kL16MaxWarps + 1];
  unsigned rcnt[kL16MaxWarps + 1];
};

template <typename T>
__global__ void __launch_bounds__(32 * (kL16MaxWarps + 1), 1)
lean16_kernel(const Geom g, const Ptrs<T> p, const SystolicCfg cfg, unsigned* sync) {
  constexpr int VW = VecTraits<T>::VW;
  constexpr int PV = VW / 4;                       // float4 per psi / table vector
  constexpr int ZR = kL16ZR;
  extern __shared__ float4 smem[];
  __shared__ Lean16Ctl ctl;
  __shared__ float4 stab[6][ZR * PV];              // CPML tables, [table][lane * PV + k]
  const int tid = threadIdx.x, lane = tid & 31;
  const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int NW = (int)(blockDim.x >> 5) - 1;       // compute warps
  const int S = cfg.stages, NT = cfg.ntiles;
  const int t = blockIdx.x % NT, j = blockIdx.x / NT;
  const int y0 = (int)((long long)t * g.Y / NT);
  const int Yt = (int)((long long)(t + 1) * g.Y / NT) - y0;
  const int X = g.X, Y = g.Y, Zq = g.Zq;
  const int psi_row = g.npg * PV;                  // float4 per psi row of a slot
  const int eslot_f4 = kL16ERows * ZR;
  const int hslot_f4 = kL16HRows * ZR + 8 * psi_row + 4;   // + 8 psi rows, 2 absorber, 2 z-source rows
  const int warp_f4 = 3 * eslot_f4 + 2 * hslot_f4 + kLeanXR * 4 * ZR;
  const int NWt = min(NW, (Yt + 2) / 2);           // warps with work on THIS tile (balanced tiles)

  unsigned* const status = sync + (size_t)S * NT * kSysFlagStride;
  unsigned* const my_prog = sync + ((size_t)j * NT + t) * kSysFlagStride;

  if (tid < (int)(sizeof(Lean16Ctl) / sizeof(unsigned))) reinterpret_cast<unsigned*>(&ctl)[tid] = 0u;
  for (int i = tid; i < 6 * ZR * PV; i += (int)blockDim.x) {
    const int k = i / (ZR * PV), r = i % (ZR * PV);
    stab[k][r] = r < Zq * PV ? __ldg(reinterpret_cast<const float4*>(p.tab + (size_t)k * g.Zp) + r)
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
  const int h = lane >> 4, q = lane & (ZR - 1);    // half (0: column A, 1: column B), z-vector
  const bool live = q < Zq;
  const int qc = live ? q : Zq - 1;                // addresses of a lane beyond the column's top
  const int c = 2 * w + h;                         // tile-local column; 0 is the y0-1 halo
  const bool own = c >= 1 && c <= Yt && live;      // E-updated and stored by this lane
  const bool doHB = 2 * w + 1 <= Yt;               // column B forms H (warp-uniform)
  const int y = wrapi(y0 - 1 + c, Y);
  const int yC = wrapi(y0 - 1 + (doHB ? 2 * w + 2 : 2 * w + 1), Y);
  const unsigned PVn = (unsigned)Y * Zq;           // vectors per x-plane
  const unsigned tv = (unsigned)y * Zq + qc, tvC = (unsigned)yC * Zq + qc;
  const int slot = psi_slot(g, qc);
  const bool has_psi = live && slot >= 0;
  const unsigned PPn = (unsigned)Y * psi_row;      // psi float4 per x-plane
  const unsigned pv = ((unsigned)y * g.npg + (has_psi ? slot : 0)) * PV;
  const bool top = q + 1 >= Zq, bottom = q == 0;
  const bool disc = cfg.discard && live && (q & 7) == 0 && c >= 2 && c <= Yt - 1;

  float4* const wbase = smem + (size_t)w * warp_f4;
  float4* const hbase = wbase + 3 * eslot_f4;
  float4* const xbase = hbase + 2 * hslot_f4;                     // boundary-H slots of this warp
  float4* const xmine = xbase + h * 2 * ZR + q;
  // (Hz, Hx) of column c-1: the A half of this warp's slot for the B lanes, the B half of warp
  // w-1's slot for the A lanes (column 0 has no predecessor and is never E-updated)
  const float4* const xread = h ? xbase + q : (w > 0 ? xbase - warp_f4 + 2 * ZR + q : xbase + q);

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
        const float4 r = stab[k][q * PV + i];
        v[4 * i] = r.x; v[4 * i + 1] = r.y; v[4 * i + 2] = r.z; v[4 * i + 3] = r.w;
      }
    }
  };
  auto load_psi = [&](const float4* ps, float (&v)[VW]) {
#pragma unroll
    for (int i = 0; i < PV; ++i) {
      const float4 r = ps[i];
      v[4 * i] = r.x; v[4 * i + 1] = r.y; v[4 * i + 2] = r.z; v[4 * i + 3] = r.w;
    }
  };
  auto store_psi = [&](float4* dst, const float (&v)[VW]) {
#pragma unroll
    for (int i = 0; i < PV; ++i)
      __stcg(dst + i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
  };

  // plane source: cheap pre-tests so that add_source() stays off the common path
  const int sp0 = g.src_pos, sp1 = wrapi(g.src_pos - 1, g.src_axis == 0 ? X : Y);
  const bool srcY = g.src_axis == 1 && (y == sp0 || y == sp1);
  const float dt = g.dt;
  const float4* const A4 = reinterpret_cast<const float4*>(p.A4);
  const float4* const S4 = reinterpret_cast<const float4*>(p.S4);
  const bool zsrc = g.src_axis == 2;
  const bool zhit = zsrc && q == g.src_pos / VW;
  const int zidx = g.src_pos % VW;

  bool ok = true;
  unsigned kk = 0;                                 // cumulative iteration count (never reset)
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
    const float4* const rEc = h ? rEz : rEx;       // this lane's share of the column after the pair
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
      return spin([&]() { return ld_vol_s(&ctl.avail) >= need && ld_vol_s(&ctl.next) >= need_next; });
    };

    // Async copies of one iteration (plane PL): E[PL+1] -> E slot se; H, B, psi, absorber row of
    // PL -> H/B slot sh.  Straight-line: halo / E-only columns are copied all the same.
    auto issue = [&](int PL, int PLn, float4* se, float4* sh, float4* se_first) {
      const unsigned vN = (unsigned)PLn * PVn, vP = (unsigned)PL * PVn;
      float4* const d = se + h * ZR + q;
      float4* const hh = sh + h * ZR + q;
      cp_async16(d + 0 * ZR, rEx + (vN + tv));
      cp_async16(d + 3 * ZR, rEz + (vN + tv));
      cp_async16(d + 6 * ZR, rEy + (vN + tv));
      cp_async16(se + (h ? 5 : 2) * ZR + q, rEc + (vN + tvC));
      cp_async16(hh + 0 * ZR, rHx + (vP + tv));
      cp_async16(hh + 2 * ZR, rHy + (vP + tv));
      cp_async16(hh + 4 * ZR, rHz + (vP + tv));
      cp_async16(hh + 6 * ZR, Bx + (vP + tv));
      cp_async16(hh + 8 * ZR, By + (vP + tv));
      cp_async16(hh + 10 * ZR, Bz + (vP + tv));
      float4* const tail = sh + kL16HRows * ZR + 8 * psi_row;
      if (q == 0) cp_async16(tail + h, A4 + ((unsigned)PL * (unsigned)Y + y));
      if (zsrc && q == 1) cp_async16(tail + 2 + h, S4 + ((unsigned)PL * (unsigned)Y + y));
      if (has_psi) {
        float4* const ps = sh + kL16HRows * ZR + h * psi_row + slot * PV;
        const unsigned pp = (unsigned)PL * PPn + pv;
#pragma unroll
        for (int k = 0; k < PV; ++k) {
          cp_async16(ps + 0 * psi_row + k, rPx + (pp + k));
          cp_async16(ps + 2 * psi_row + k, rPy + (pp + k));
          cp_async16(ps + 4 * psi_row + k, ePx + (pp + k));
          cp_async16(ps + 6 * psi_row + k, ePy + (pp + k));
        }
      }
      if (se_first) {                                // very first plane of the sweep: E[PL] too
        float4* const f = se_first + h * ZR + q;
        cp_async16(f + 0 * ZR, rEx + (vP + tv));
        cp_async16(f + 3 * ZR, rEz + (vP + tv));
        cp_async16(f + 6 * ZR, rEy + (vP + tv));
        cp_async16(se_first + (h ? 5 : 2) * ZR + q, rEc + (vP + tvC));
      }
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

    float hyp[VW], hzp[VW];                        // H^{n+1/2}[P-1] of the thread's own cells
#pragma unroll
    for (int v = 0; v < VW; ++v) { hyp[v] = 0.f; hzp[v] = 0.f; }

    for (int i = 0; i <= X && ok; ++i) {
      const bool real = i >= 1;
      const int Pn = P + 1 == X ? 0 : P + 1;
      const unsigned vP = (unsigned)P * PVn;
      cp_async_wait<0>();                          // this lane's copies of iteration i have landed
      if (i < X) {
        ok = wait_deps(i + 1);
        if (!ok) break;
        issue(Pn, Pn + 1 == X ? 0 : Pn + 1, snext, hnext, nullptr);
      }
      cp_async_commit();
      __syncwarp();                                // ... and so have those of the other lanes
      if (w == 0 && lane == 0) st_vol_s(&ctl.front, iters_done + (unsigned)i);
      // Dead lines (see kernels_lean.cuh): Ey and H of both columns and (Ex, Ez) of column B have
      // been read by their only reader; (Ex, Ez) of column A wait for warp w-1 (E half-step).
      if (disc && i >= 1) {
        const unsigned vN = (unsigned)Pn * PVn;
        discard_l2_line(rEy + (vN + tv));
        discard_l2_line(rHx + (vP + tv)); discard_l2_line(rHy + (vP + tv));
        discard_l2_line(rHz + (vP + tv));
        if (h) { discard_l2_line(rEx + (vN + tv)); discard_l2_line(rEz + (vN + tv)); }
      }

      const unsigned pP = (unsigned)P * PPn + pv;

      // ---------------------------------- H half-step ---------------------------------------------
      const float4* const ep = sprev + h * ZR + q;
      const float4* const ec = scur + h * ZR + q;
      const float4* const hc = hcur + h * ZR + q;
      const float4* const pc = hcur + kL16HRows * ZR + h * psi_row + (has_psi ? slot : 0) * PV;
      const float4* const tail = hcur + kL16HRows * ZR + 8 * psi_row;
      float ex[VW], ey[VW], ez[VW], hx[VW], hy[VW], hz[VW], psx[VW], psy[VW];
      {
        float exy[VW], ezy[VW], eyx[VW], ezx[VW], ah[VW], bh[VW], ikh[VW];
        unpack(lds16(ep + 0 * ZR), ex, T()); unpack(lds16(ep + 3 * ZR), ez, T());
        unpack(lds16(ep + 6 * ZR), ey, T());
        unpack(lds16(ep + 1 * ZR), exy, T()); unpack(lds16(ep + 4 * ZR), ezy, T());
        unpack(lds16(ec + 6 * ZR), eyx, T()); unpack(lds16(ec + 3 * ZR), ezx, T());
        unpack(lds16(hc + 0 * ZR), hx, T()); unpack(lds16(hc + 2 * ZR), hy, T());
        unpack(lds16(hc + 4 * ZR), hz, T());
#pragma unroll
        for (int v = 0; v < VW; ++v) { psx[v] = 0.f; psy[v] = 0.f; }
        if (has_psi) { load_psi(pc, psx); load_psi(pc + 2 * psi_row, psy); }
        load_tab(3, ah); load_tab(4, bh); load_tab(5, ikh);
        float ex_top = __shfl_down_sync(0xffffffffu, ex[0], 1, ZR);
        float ey_top = __shfl_down_sync(0xffffffffu, ey[0], 1, ZR);
        if (top) { ex_top = 0.f; ey_top = 0.f; }
#pragma unroll
        for (int v = 0; v < VW; ++v) {
          const float exz = (v + 1 < VW) ? ex[(v + 1) % VW] : ex_top;
          const float eyz = (v + 1 < VW) ? ey[(v + 1) % VW] : ey_top;
          h_cell(ex[v], ey[v], ez[v], exz, eyz, ezy[v], exy[v], eyx[v], ezx[v], ah[v], bh[v],
                 ikh[v], dt, psx[v], psy[v], hx[v], hy[v], hz[v]);
          hx[v] = round_store<T>(hx[v]); hy[v] = round_store<T>(hy[v]); hz[v] = round_store<T>(hz[v]);
        }
      }
      // boundary H for the next column: the B half waits until warp w+1 has consumed the slot
      {
        float4* const xs = xmine + (kk & (kLeanXR - 1)) * 4 * ZR;
        if (kk >= (unsigned)kLeanXR && w + 1 < NWt) {
          const unsigned need = kk + 1u - (unsigned)kLeanXR;
          ok = spin([&]() { return ld_vol_s(&ctl.rcnt[w]) >= need; });
          if (!ok) break;
        }
        xs[0] = pack(hz, T());
        xs[ZR] = pack(hx, T());
        __syncwarp();
        if (lane == 0) st_vol_s(&ctl.hcnt[w], kk + 1u);
      }

      // ---------------------------------- E half-step ---------------------------------------------
      if (real) {
        if (w > 0) {
          const unsigned need = kk + 1u;
          ok = spin([&]() { return ld_vol_s(&ctl.hcnt[w - 1]) >= need; });
          if (!ok) break;
        }
        float hzm[VW], hxm[VW];                    // (Hz, Hx) of column c-1
        {
          const float4* const xr = xread + (kk & (kLeanXR - 1)) * 4 * ZR;
          unpack(lds16(xr), hzm, T()); unpack(lds16(xr + ZR), hxm, T());
        }
        __syncwarp();
        if (w > 0) {
          if (lane == 0) st_vol_s(&ctl.rcnt[w - 1], kk + 1u);
          // warp w-1 has finished the H half-step of this iteration: its copy of column A's
          // (Ex, Ez)[P+1] -- the only other reader -- has landed
          if (disc && !h) {
            const unsigned vN = (unsigned)Pn * PVn;
            discard_l2_line(rEx + (vN + tv)); discard_l2_line(rEz + (vN + tv));
          }
        }
        float hx_bot = __shfl_up_sync(0xffffffffu, hx[VW - 1], 1, ZR);
        float hy_bot = __shfl_up_sync(0xffffffffu, hy[VW - 1], 1, ZR);
        if (bottom) { hx_bot = 0.f; hy_bot = 0.f; }
        if (own) {
          float b0[VW], b1[VW], b2[VW], qsx[VW], qsy[VW], ae[VW], be[VW], ike[VW];
          unpack(lds16(hc + 6 * ZR), b0, T()); unpack(lds16(hc + 8 * ZR), b1, T());
          unpack(lds16(hc + 10 * ZR), b2, T());
          const float4 aa = lds16(tail + h);
#pragma unroll
          for (int v = 0; v < VW; ++v) { qsx[v] = 0.f; qsy[v] = 0.f; }
          if (has_psi) { load_psi(pc + 4 * psi_row, qsx); load_psi(pc + 6 * psi_row, qsy); }
          load_tab(0, ae); load_tab(1, be); load_tab(2, ike);
#pragma unroll
          for (int v = 0; v < VW; ++v) {
            const float hxz = (v > 0) ? hx[(v + VW - 1) % VW] : hx_bot;
            const float hyz = (v > 0) ? hy[(v + VW - 1) % VW] : hy_bot;
            e_cell(hx[v], hy[v], hz[v], hxz, hyz, hzm[v], hxm[v], hyp[v], hzp[v], ae[v], be[v],
                   ike[v], aa.x, aa.y, aa.z, b0[v], b1[v], b2[v], qsx[v], qsy[v], ex[v], ey[v], ez[v]);
          }
          if (zsrc) {
            // z-plane source: same operation order as add_source() (channel 0, then 1)
            const float4 s = lds16(tail + 2 + h);
#pragma unroll
            for (int v = 0; v < VW; ++v) {
              const float t0 = fmaf(w1, s.z, fmaf(w0, s.x, ex[v]));
              const float t1 = fmaf(w1, s.w, fmaf(w0, s.y, ey[v]));
              const bool hit = zhit && v == zidx;
              ex[v] = hit ? t0 : ex[v];
              ey[v] = hit ? t1 : ey[v];
            }
          } else if (g.src_axis == 0 ? (P == sp0 || P == sp1) : srcY) {
            add_source<VW>(g, p.src, w0, w1, P, y, q, ex, ey, ez);
          }
          const unsigned o = vP + tv;
          __stcg(wHx + o, pack(hx, T())); __stcg(wHy + o, pack(hy, T())); __stcg(wHz + o, pack(hz, T()));
          __stcg(wEx + o, pack(ex, T())); __stcg(wEy + o, pack(ey, T())); __stcg(wEz + o, pack(ez, T()));
          if (has_psi) {
            store_psi(wPx + pP, psx); store_psi(wPy + pP, psy);
            store_psi(ePx + pP, qsx); store_psi(ePy + pP, qsy);
          }
          if (oi >= 0) {
#pragma unroll
            for (int v = 0; v < VW; ++v) {
              ex[v] = round_store<T>(ex[v]); ey[v] = round_store<T>(ey[v]); ez[v] = round_store<T>(ez[v]);
            }
            write_snapshot<VW>(g, p.out, oi, P, y, q, ex, ey, ez, p.proj);
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
  if (lane == 0) atomicAdd(&ctl.exited, 1u);
}

// Compute warps for a tile of `tile_y` owned columns: columns 0 .. tile_y form H, two per warp.
inline int lean16_warps(int tile_y) { return (tile_y + 2) / 2; }

inline size_t lean16_smem_bytes(const Geom& g, int vw, int tile_y) {
  const size_t psi_row = (size_t)g.npg * (vw / 4);
  const size_t eslot_f4 = (size_t)kL16ERows * kL16ZR;
  const size_t hslot_f4 = (size_t)kL16HRows * kL16ZR + 8 * psi_row + 4;
  const size_t warp_f4 = 3 * eslot_f4 + 2 * hslot_f4 + (size_t)kLeanXR * 4 * kL16ZR;
  return sizeof(float4) * warp_f4 * lean16_warps(tile_y);
}

template <typename T>
inline bool lean16_configure(const Geom& g, int tile_y_req, int stages_req, int sms, int l2_bytes,
                             SystolicCfg* cfg, std::string* why) {
  constexpr int VW = VecTraits<T>::VW;
  if (g.Zq > kL16ZR) { *why = "needs a z-column of at most 16 vectors"; return false; }
  if (g.N / VW * 3 >= (1ll << 32)) { *why = "domain too large for 32-bit vector indices"; return false; }
  int max_tile = 2 * kL16MaxWarps - 1;           // 2*NW - 1 owned columns fill NW warps exactly
  while (max_tile >= 1 && lean16_smem_bytes(g, VW, max_tile) + 4096 > 227 * 1024) --max_tile;
  if (max_tile < 1) { *why = "staging ring does not fit in shared memory"; return false; }
  if (tile_y_req > 0 && tile_y_req < max_tile) max_tile = tile_y_req;
  if (max_tile > g.Y) max_tile = g.Y;
  const int ntiles = (g.Y + max_tile - 1) / max_tile;
  const int widest = (g.Y + ntiles - 1) / ntiles;
  cfg->tile_y = widest;
  cfg->ntiles = ntiles;
  cfg->cols = 16;                                // marks the half-warp variant for the launcher
  cfg->threads = 32 * (lean16_warps(widest) + 1);
  cfg->smem_bytes = (int)lean16_smem_bytes(g, VW, widest);
  cfg->max_lead = 10;
  cfg->pf_ahead = 6;
  cfg->svc_sleep_ns = 200;
  cfg->spin_ns_max = 160;
  cfg->discard = (g.Zq % 8 == 0) ? 1 : 0;        // a column must be whole 128-byte lines
  if (const char* e = getenv("B200FDTD_LEAN_DISCARD")) cfg->discard = cfg->discard && atoi(e);
  if (const char* e = getenv("B200FDTD_SPIN_NS")) cfg->spin_ns_max = atoi(e);
  if (const char* e = getenv("B200FDTD_MAX_LEAD")) cfg->max_lead = atoi(e);
  if (const char* e = getenv("B200FDTD_PF_AHEAD")) cfg->pf_ahead = atoi(e);
  if (const char* e = getenv("B200FDTD_SVC_SLEEP")) cfg->svc_sleep_ns = atoi(e);
  if (cfg->max_lead < 6) cfg->max_lead = 6;
  if (cfg->pf_ahead < 0) cfg->pf_ahead = 0;
  cfg->trap_on_timeout = 1;
  cfg->need_zfix = 0;
  cfg->unroll = 1;
  int occ = 0;
  const void* fn = (const void*)lean16_kernel<T>;
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
inline int lean16_launch(const Geom& g, const Ptrs<T>& p, const SystolicCfg& cfg, unsigned* sync,
                         cudaStream_t st) {
  const void* fn = (const void*)lean16_kernel<T>;
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
