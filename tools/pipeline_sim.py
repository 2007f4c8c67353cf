"""Time-stepped model of the systolic protocol (stages x tiles of CTAs, k+3 rule on the three
predecessor tiles, max_lead throttle on the successor, counters seen `lat` plane-times late):
what fraction of its time does a CTA wait, and what does the pipeline lose against its slowest
member?  Used to reason about the wait accounting of DESIGN.md 4.0 without a GPU.

  python tools/pipeline_sim.py [--stages 8 --tiles 18 --X 256 --max-lead 10 --lat 1.0 --jitter 0.1 ...]
"""
import argparse
import numpy as np


def simulate(S=8, NT=18, X=256, sweeps=5, max_lead=10, lat=1.0, jitter=0.1, sm_spread=0.04,
             stage0_extra=0.1, dt=0.02, seed=0, rule=2):
  rng = np.random.default_rng(seed)
  speed = 1.0 + sm_spread * rng.standard_normal((S, NT))          # persistent per-CTA factor
  speed[0] += stage0_extra                                          # stage 0 reads from HBM
  it = np.zeros((S, NT), int)            # iteration being executed (0..X), per sweep
  m = np.zeros((S, NT), int)             # sweep index
  remaining = np.full((S, NT), -1.0)     # time left of the current iteration; < 0 = waiting to start
  count = np.zeros((S, NT), int)         # published count (finished sweep indices, cumulative)
  hist = []                              # (time, count) snapshots for the delayed view
  waited = np.zeros((S, NT))
  busy = np.zeros((S, NT))
  done = np.zeros((S, NT), bool)
  t = 0.0
  nlag = max(1, int(round(lat / dt)))
  ring = [count.copy() for _ in range(nlag + 1)]
  jj = np.arange(S)[:, None]
  while not done.all():
    seen = ring[0]                        # counts as they were `lat` ago
    # dependency check for CTAs that want to start an iteration
    want = (remaining < 0) & ~done
    prev = np.roll(seen, 1, axis=0)       # predecessor stage (stage 0 <- stage S-1, previous round)
    avail = np.minimum(np.minimum(np.roll(prev, 1, axis=1), prev), np.roll(prev, -1, axis=1))
    base_prev = np.where(jj > 0, m, m - 1) * X
    has_prev = (m > 0) | (jj > 0)
    need = np.where(has_prev, base_prev + np.minimum(it + rule, X), 0)
    nxt = np.roll(seen, -1, axis=0)
    lead = np.minimum(it, X) - 1 - max_lead
    has_next = (jj + 1 < S)
    need_next = np.where(has_next & (lead > 0), m * X + lead, 0)
    ok = want & (avail >= need) & (nxt >= need_next)
    dur = speed * (1.0 + jitter * rng.standard_normal((S, NT)))
    remaining = np.where(ok, np.maximum(dur, 0.2), remaining)
    waited += np.where(want & ~ok, dt, 0.0)
    running = remaining >= 0
    busy += np.where(running, dt, 0.0)
    remaining = np.where(running, remaining - dt, remaining)
    fin = running & (remaining <= 0)
    # an iteration finished: publish (i >= 1), advance
    count = np.where(fin & (it >= 1), m * X + it, count)
    it = np.where(fin, it + 1, it)
    wrap = fin & (it > X)
    m = np.where(wrap, m + 1, m)
    it = np.where(wrap, 0, it)
    done |= wrap & (m >= sweeps)
    remaining = np.where(fin, -1.0, remaining)
    ring.pop(0)
    ring.append(count.copy())
    t += dt
  total_iters = sweeps * (X + 1)
  return {"time_per_plane": t / total_iters, "slowest_cta_busy_per_plane": float((busy / total_iters).max()),
          "mean_busy_per_plane": float((busy / total_iters).mean()),
          "wait_frac_stage0": float((waited[0] / t).mean()), "wait_frac_others": float((waited[1:] / t).mean())}


if __name__ == "__main__":
  ap = argparse.ArgumentParser()
  ap.add_argument("--stages", type=int, default=8)
  ap.add_argument("--tiles", type=int, default=18)
  ap.add_argument("--X", type=int, default=256)
  ap.add_argument("--sweeps", type=int, default=4)
  ap.add_argument("--max-lead", type=int, nargs="+", default=[10])
  ap.add_argument("--lat", type=float, nargs="+", default=[0.5, 1.0, 2.0])
  ap.add_argument("--jitter", type=float, default=0.1)
  ap.add_argument("--sm-spread", type=float, default=0.04)
  ap.add_argument("--stage0-extra", type=float, default=0.1)
  a = ap.parse_args()
  for L in a.max_lead:
    for lat in a.lat:
      r = simulate(a.stages, a.tiles, a.X, a.sweeps, L, lat, a.jitter, a.sm_spread, a.stage0_extra)
      print(f"max_lead {L:3d} lat {lat:4.1f}: time/plane {r['time_per_plane']:.3f} (slowest CTA busy "
            f"{r['slowest_cta_busy_per_plane']:.3f}, mean {r['mean_busy_per_plane']:.3f}); waits: stage 0 "
            f"{r['wait_frac_stage0']:.2f}, others {r['wait_frac_others']:.2f}")
