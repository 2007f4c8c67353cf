"""Event-level CPU emulation of the y-slab sessions' schedule (pjz_b200/csrc/kernels_lean.cuh,
``SLAB = true``): W ranks, each with its own ghost columns; per rank the (stage, tile) actors of
tests/systolic_emulator.py plus COURIER actors -- one per (side, stage) -- that copy each newly
published plane of a slab's edge column into the neighbour's ghost storage and then forward the
edge tile's counter into the neighbour's mirror slot.  Edge tiles take the counters of the tile
beyond the slab from their mirror slots, and (the write-after-read rule the courier adds) may
overwrite a plane of their edge column only once the courier has carried off what the step two
back wrote there.  The next actor is picked at RANDOM among the ready ones, couriers included, so
a courier may lag arbitrarily: if any rule were too weak some interleaving would read a ghost that
has not arrived, or copy an edge column that was already overwritten, and the result would differ
from the oracle.  Arithmetic is the base emulator's (float64, the oracle's operation order).
"""

import numpy as np

from tests.systolic_emulator import Emulator


class SlabEmulator(Emulator):

  def __init__(self, kw, world, tiles_per_rank, stages, max_lead=6, need_rule=3, seed=0,
               courier_war_rule=True, counter_before_data=False, discard=False,
               discard_edge_columns=False):
    super().__init__(kw, world * tiles_per_rank, stages, max_lead=max_lead, need_rule=need_rule,
                     seed=seed, discard=discard)
    assert self.Y % world == 0 and (self.Y // world) >= tiles_per_rank
    self.W, self.NTr = world, tiles_per_rank
    self.war = courier_war_rule
    self.discard_edges = discard_edge_columns               # negative control: edge columns too
    self.counter_first = counter_before_data                # negative control: forward, copy later
    self.pending = []                                       # copies owed in that mode
    self.war_would_block = 0                                # times the WAR rule was the binding one
    X, Z = self.X, self.Z
    nan = lambda *s: np.full(s, np.nan)
    # ghost storage per rank and buffer set: low ghost = the low neighbour's last column
    # (E, H, psiH), high ghost = the high neighbour's first column (Ex, Ez; Ey is never sent)
    self.glo = [[dict(E=np.zeros((3, X, Z)), H=np.zeros((3, X, Z)), psiH=np.zeros((2, X, Z)))
                 for _ in range(2)] for _ in range(world)]
    self.ghi = [[dict(E=np.stack([np.zeros((X, Z)), nan(X, Z), np.zeros((X, Z))]))
                 for _ in range(2)] for _ in range(world)]
    self.mirror = np.zeros((world, 2, self.S), np.int64)    # [rank][0: low nb's last tile | 1: high nb's first][stage]
    self.cour = np.zeros((world, 2, self.S), np.int64)      # [rank][side][stage]: counts carried off
    for (j, t), a in self.actor.items():
      a["t"] = t
    self.couriers = [(r, side, j) for r in range(world) for side in range(2) for j in range(self.S)]

  # ---- geometry ------------------------------------------------------------------------------------
  def rank_of(self, t):
    return t // self.NTr

  def tile(self, t):
    r, tl = divmod(t, self.NTr)
    yo = self.Y // self.W
    return r * yo + tl * yo // self.NTr, r * yo + (tl + 1) * yo // self.NTr

  # ---- loads with ghost substitution ---------------------------------------------------------------
  def _h_new(self, rb, P, cols, a=None, k=0):
    t = a["t"]
    r, tl = divmod(t, self.NTr)
    first, last = tl == 0, tl == self.NTr - 1
    X, Y = self.X, self.Y
    tb = self.st.t
    E, H = self.E[rb], self.H[rb]
    Pn = (P + 1) % X
    cy = np.asarray(cols) % Y
    load = np.asarray(list(cols) + [cols[-1] + 1]) % Y
    cached = self.discard and a is not None and k >= 0      # E[P] was loaded as "next" last iteration
    ecur = a["ecache"] if cached else np.stack([E[c][P][load] for c in range(3)])
    enext = np.stack([E[c][Pn][load] for c in range(3)])
    hold = np.stack([H[c][P][cy] for c in range(3)])
    psx, psy = self.psiH[rb][0][P][cy].copy(), self.psiH[rb][1][P][cy].copy()
    if first:                                               # column y0-1 is this rank's LOW ghost
      g = self.glo[r][rb]
      if not cached:
        ecur[:, 0] = g["E"][:, P]
      enext[:, 0], hold[:, 0] = g["E"][:, Pn], g["H"][:, P]
      psx[0], psy[0] = g["psiH"][0][P], g["psiH"][1][P]
    if last:                                                # column y1 is this rank's HIGH ghost
      g = self.ghi[r][rb]
      if not cached:
        ecur[:, -1] = g["E"][:, P]
      enext[:, -1] = g["E"][:, Pn]
    if self.discard and a is not None:
      a["ecache"] = enext.copy()
      if k >= 0:                                            # kernels_lean.cuh: discard.global.L2 of the
        dead = np.asarray(cols[2:-1]) % Y if len(cols) > 3 else np.asarray([], np.int64)
        if self.discard_edges:
          dead = np.asarray(cols[1:]) % Y
        for c in range(3):                                  # tile's exclusive columns, never the edge
          E[c][Pn][dead] = np.nan                           # columns a neighbour or a courier reads
          H[c][P][dead] = np.nan
    ex, ey, ez = ecur[0][:-1], ecur[1][:-1], ecur[2][:-1]
    from oracle import fdtd_numpy as spec
    dzEy, dzEx = spec._dz_fwd(ey), spec._dz_fwd(ex)
    px = tb["b_h"] * psx + tb["a_h"] * dzEy
    py = tb["b_h"] * psy + tb["a_h"] * dzEx
    cx = (ecur[2][1:] - ez) - (dzEy * tb["ik_h"] + px)
    cyv = (dzEx * tb["ik_h"] + py) - (enext[2][:-1] - ez)
    cz = (enext[1][:-1] - ey) - (ecur[0][1:] - ex)
    dt = self.st.dt_t
    h = np.stack([hold[0] - dt * cx, hold[1] - dt * cyv, hold[2] - dt * cz])
    return h, np.stack([px, py]), (ex, ey, ez)

  # ---- dependencies --------------------------------------------------------------------------------
  def ready(self, j, t):
    a = self.actor[(j, t)]
    n, k = a["n"], a["k"]
    if n >= self.tt:
      return False
    X, S = self.X, self.S
    r, tl = divmod(t, self.NTr)
    m = n // S
    kk = max(k, 0)
    if n > 0:
      jp = (j - 1) % S
      base_prev = (m if j > 0 else m - 1) * X
      need = base_prev + min(kk + self.need_rule, X)
      for d in (-1, 0, 1):
        if 0 <= tl + d < self.NTr:
          have = self.prog[jp][t + d]
        else:                                               # beyond the slab: the neighbour's edge tile
          have = self.mirror[r][0 if d < 0 else 1][jp]
        if have < need:
          return False
    if n + 1 < self.tt and j + 1 < S and k > self.max_lead:
      if self.prog[j + 1][t] < m * X + k - self.max_lead:
        return False
    if n - 2 >= 0 and k >= 0:                               # the courier's write-after-read rule
      jq, mq = (n - 2) % S, (n - 2) // S
      need_c = mq * X + min(k + 3, X)
      late = ((tl == 0 and self.cour[r][0][jq] < need_c) or
              (tl == self.NTr - 1 and self.cour[r][1][jq] < need_c))
      if late:
        self.war_would_block += 1
        if self.war:
          return False
    return True

  # ---- couriers ------------------------------------------------------------------------------------
  def courier_ready(self, r, side, j):
    t_edge = r * self.NTr + (0 if side == 0 else self.NTr - 1)
    return self.prog[j][t_edge] > self.cour[r][side][j]

  def courier_advance(self, r, side, j):
    cnt = self.cour[r][side][j] + 1
    if self.counter_first:
      self.pending.append((r, side, j, cnt))
    else:
      self._copy(r, side, j, cnt)
    self.cour[r][side][j] = cnt
    nb = (r - 1) % self.W if side == 0 else (r + 1) % self.W
    self.mirror[nb][1 if side == 0 else 0][j] = cnt         # I am their high / low neighbour

  def _copy(self, r, side, j, cnt):
    X, S, Y = self.X, self.S, self.Y
    m, k = divmod(cnt - 1, X)
    n = j + m * S
    P = (n % X + k) % X
    wb = (n + 1) & 1
    yo = Y // self.W
    if side == 0:                                           # first owned column -> low neighbour's HIGH ghost
      y, dst = r * yo, self.ghi[(r - 1) % self.W][wb]
      dst["E"][0][P] = self.E[wb][0][P][y]
      dst["E"][2][P] = self.E[wb][2][P][y]
    else:                                                   # last owned column -> high neighbour's LOW ghost
      y, dst = r * yo + yo - 1, self.glo[(r + 1) % self.W][wb]
      for c in range(3):
        dst["E"][c][P] = self.E[wb][c][P][y]
        dst["H"][c][P] = self.H[wb][c][P][y]
      for c in range(2):
        dst["psiH"][c][P] = self.psiH[wb][c][P][y]

  def run(self):
    keys = list(self.actor.keys())
    while True:
      live = [key for key in keys if self.actor[key]["n"] < self.tt]
      ready = [("tile",) + key for key in live if self.ready(*key)]
      ready += [("courier",) + c for c in self.couriers if self.courier_ready(*c)]
      ready += [("copy", i) for i in range(len(self.pending))]
      if not live and not ready:
        break
      if not ready:
        raise RuntimeError("deadlock: no actor can advance")
      kind, *key = ready[self.rng.integers(len(ready))]
      if kind == "tile":
        self.advance(*key)
      elif kind == "copy":
        self._copy(*self.pending.pop(key[0]))
      else:
        self.courier_advance(*key)
    return self.out
