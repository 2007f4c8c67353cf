"""Event-level CPU emulation of the systolic kernel's SCHEDULE (pjz_b200/csrc/kernels_systolic.cuh).

It executes the same decomposition -- (stage, y-tile) actors, fused H+E update of one x-plane
per iteration, ping-pong buffer sets, start plane n mod X, prologue, progress counters with
the `min(k+3, X)` rule and the `max_lead` throttle -- but picks the next actor to advance at
RANDOM among those whose counters allow it.  If the dependency rule were too weak, some
interleaving would read a plane that is not yet written / already overwritten and the result
would differ from the oracle.  Arithmetic is float64 NumPy in the oracle's operation order.

``discard=True`` additionally models kernels_lean.cuh's ``discard.global.L2``: an actor keeps the
E plane it loaded as "next" for the following iteration (the shared-memory ring) and, right after
loading E^n[P+1] and H^{n-1/2}[P] in any iteration but the sweep's first, POISONS (NaN) the global
copy of those planes on the tile's exclusive columns (tile-local 2 .. Yt-1).  A line that was not
really dead -- read again by this tile, by a neighbour's halo, or by a later stage before being
rewritten -- would put NaNs into the result.
"""

import numpy as np

from oracle import fdtd_numpy as spec


class Emulator:

  def __init__(self, kw, ntiles, stages, max_lead=6, need_rule=3, seed=0, discard=False,
               discard_first=False):
    self.kw = kw
    self.discard, self.discard_first = discard, discard_first
    eps = np.asarray(kw["epsilon"], np.float32)
    self.st = spec.State(eps, kw["dt"], kw["absorption_mask"], kw["pml_kappa"], kw["pml_sigma"],
                         kw["pml_alpha"], kw["pml_widths"], kw["offset"], np.float64)
    self.X, self.Y, self.Z = self.st.shape
    self.sf = np.asarray(kw["source_field"], np.float64)
    self.wf = np.asarray(kw["source_waveform"], np.float32).astype(np.float64)
    self.axis = spec.source_axis(self.sf)
    self.tt = self.wf.shape[0]
    self.NT, self.S = ntiles, min(stages, self.tt)
    self.max_lead, self.need_rule = max_lead, need_rule
    self.rng = np.random.default_rng(seed)
    shp = (3, self.X, self.Y, self.Z)
    self.E = [np.zeros(shp), np.zeros(shp)]
    self.H = [np.zeros(shp), np.zeros(shp)]
    self.psiH = [np.zeros((2,) + shp[1:]), np.zeros((2,) + shp[1:])]
    self.psiE = np.zeros((2,) + shp[1:])
    self.prog = np.zeros((self.S, self.NT), np.int64)
    self.outs = list(range(*kw["output_steps"]))
    _, xx, yy, zz = eps.shape
    self.out = np.zeros((len(self.outs), 3, xx, yy, zz), np.float32)
    self.sub = (xx, yy, zz)
    # per-actor cursor: (n, k, hprev) ; k = -1 is the prologue
    self.actor = {}
    for j in range(self.S):
      for t in range(self.NT):
        self.actor[(j, t)] = dict(n=j, k=-1, hprev=None, ecache=None)

  def tile(self, t):
    return t * self.Y // self.NT, (t + 1) * self.Y // self.NT

  def _h_new(self, rb, P, cols, a=None, k=0):
    """H^{n+1/2}[P] on the given columns from read set rb (returns (3,len,Z), psiH new).

    With ``discard`` the E plane P comes from the actor's cache (it was loaded as P+1 one
    iteration earlier), E[P+1] and H[P] are loaded now and their exclusive columns poisoned."""
    X, Y = self.X, self.Y
    t = self.st.t
    E, H = self.E[rb], self.H[rb]
    Pn = (P + 1) % X
    cy = np.asarray(cols) % Y
    load = np.asarray(list(cols) + [cols[-1] + 1]) % Y        # columns y0-1 .. y1 (Yt + 2 of them)
    if self.discard and a is not None:
      ecur = a["ecache"] if k >= 0 else np.stack([E[c][P][load] for c in range(3)])
      enext = np.stack([E[c][Pn][load] for c in range(3)])
      hold = np.stack([H[c][P][cy] for c in range(3)])
      a["ecache"] = enext.copy()
      if k >= 0 or self.discard_first:                        # never the sweep's first loads
        dead = np.asarray(cols[2:-1]) % Y if len(cols) > 3 else np.asarray([], np.int64)
        for c in range(3):                                    # tile-local columns 2 .. Yt-1
          E[c][Pn][dead] = np.nan
          H[c][P][dead] = np.nan
    else:
      ecur = np.stack([E[c][P][load] for c in range(3)])
      enext = np.stack([E[c][Pn][load] for c in range(3)])
      hold = np.stack([H[c][P][cy] for c in range(3)])
    ex, ey, ez = ecur[0][:-1], ecur[1][:-1], ecur[2][:-1]
    dzEy = spec._dz_fwd(ey)
    dzEx = spec._dz_fwd(ex)
    px = t["b_h"] * self.psiH[rb][0][P][cy] + t["a_h"] * dzEy
    py = t["b_h"] * self.psiH[rb][1][P][cy] + t["a_h"] * dzEx
    cx = (ecur[2][1:] - ez) - (dzEy * t["ik_h"] + px)
    cyv = (dzEx * t["ik_h"] + py) - (enext[2][:-1] - ez)
    cz = (enext[1][:-1] - ey) - (ecur[0][1:] - ex)
    dt = self.st.dt_t
    h = np.stack([hold[0] - dt * cx, hold[1] - dt * cyv, hold[2] - dt * cz])
    return h, np.stack([px, py]), (ex, ey, ez)

  def ready(self, j, t):
    a = self.actor[(j, t)]
    n, k = a["n"], a["k"]
    if n >= self.tt:
      return False
    X, S, NT = self.X, self.S, self.NT
    m = n // S
    if n > 0:
      jp = (j - 1) % S
      base_prev = (m if j > 0 else m - 1) * X
      kk = max(k, 0)
      need = base_prev + min(kk + self.need_rule, X)
      for tt_ in (t - 1, t, t + 1):
        if self.prog[jp][tt_ % NT] < need:
          return False
    if n + 1 < self.tt and j + 1 < S and k > self.max_lead:
      if self.prog[j + 1][t] < m * X + k - self.max_lead:
        return False
    return True

  def advance(self, j, t):
    a = self.actor[(j, t)]
    n, k = a["n"], a["k"]
    X, Y = self.X, self.Y
    rb, wb = n & 1, (n + 1) & 1
    y0, y1 = self.tile(t)
    P = (n % X + k) % X
    cols_h = list(range(y0 - 1, y1))            # H formed on y0-1 .. y1-1
    h, psi, ecur = self._h_new(rb, P, cols_h, a, k)
    if k >= 0:
      own = np.arange(y0, y1)
      hx, hy, hz = h[0][1:], h[1][1:], h[2][1:]  # own columns
      hz_ym, hx_ym = h[2][:-1], h[0][:-1]        # y-1 neighbours (incl. recomputed halo)
      hy_xm, hz_xm = a["hprev"]
      tb = self.st.t
      dzHy, dzHx = spec._dz_bwd(hy), spec._dz_bwd(hx)
      qx = tb["b_e"] * self.psiE[0][P][own] + tb["a_e"] * dzHy
      qy = tb["b_e"] * self.psiE[1][P][own] + tb["a_e"] * dzHx
      cx = (hz - hz_ym) - (dzHy * tb["ik_e"] + qx)
      cy = (dzHx * tb["ik_e"] + qy) - (hz - hz_xm)
      cz = (hy - hy_xm) - (hx - hx_ym)
      A, B = self.st.A, self.st.B
      e = [A[c][P][own] * ecur[c][1:] + B[c][P][own] * cc for c, cc in enumerate((cx, cy, cz))]
      # source
      w = self.wf[n]
      p = int(self.kw["source_position"])
      if self.axis == 0:
        for ch in range(2):
          if P == (p - ch) % X:
            e[1] = e[1] + w[ch] * self.sf[0, 0][own]
            e[2] = e[2] + w[ch] * self.sf[1, 0][own]
      elif self.axis == 1:
        for ch in range(2):
          yp = (p - ch) % Y
          if y0 <= yp < y1:
            e[0][yp - y0] += w[ch] * self.sf[0, P, 0]
            e[2][yp - y0] += w[ch] * self.sf[1, P, 0]
      else:
        for ch in range(2):
          e[0][:, p] += w[ch] * self.sf[ch, 0, P, own, 0]
          e[1][:, p] += w[ch] * self.sf[ch, 1, P, own, 0]
      for c in range(3):
        self.E[wb][c][P][own] = e[c]
        self.H[wb][c][P][own] = h[c][1:]
      self.psiH[wb][0][P][own] = psi[0][1:]
      self.psiH[wb][1][P][own] = psi[1][1:]
      self.psiE[0][P][own] = qx
      self.psiE[1][P][own] = qy
      if n in self.outs:
        oi = self.outs.index(n)
        ox, oy, oz = (int(o) for o in self.kw["offset"])
        xx, yy, zz = self.sub
        if ox <= P < ox + xx:
          for yy_ in own:
            if oy <= yy_ < oy + yy:
              for c in range(3):
                self.out[oi, c, P - ox, yy_ - oy] = e[c][yy_ - y0][oz:oz + zz]
      self.prog[j][t] = (n // self.S) * X + k + 1
    a["hprev"] = (h[1][1:].copy(), h[2][1:].copy())
    a["k"] = k + 1
    if a["k"] == X:
      a["n"], a["k"], a["hprev"], a["ecache"] = n + self.S, -1, None, None

  def run(self):
    keys = list(self.actor.keys())
    while True:
      live = [key for key in keys if self.actor[key]["n"] < self.tt]
      if not live:
        break
      ready = [key for key in live if self.ready(*key)]
      if not ready:
        raise RuntimeError("deadlock: no actor can advance")
      self.advance(*ready[self.rng.integers(len(ready))])
    return self.out
