"""CPU simulation of the selection kernel's shared threshold at the DAVIS shape (16 200 keys, 22 virtual splits in lock
step, 64-key tiles, trackers fed with the maxima of the 8-column groups, thresholds one or two tiles behind): appends per
query and candidates left at the hand-off for the ways of combining the published scores (csrc/select_tc.cu):

  min    minimum over the virtual splits of their 2nd best (round 1 / start of round 2)
  pair   splits combined two at a time: the 3rd largest of a pair's 4 published values
  quad   four at a time (refresh_grouped, kept): five groups vouch for 6 keys each, the last pair for 3
  exact  the 33rd largest of all 44 published values (what probe counting approaches; too many instructions, dropped)

Runs on the CPU (oracle scores); no GPU needed.  Output kept in profiles/r2_threshold_sim.txt."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import readout_oracle as orc
from tests import synth
from tests.test_thresholds_cpu import grouped_bound

g = torch.Generator().manual_seed(1236)
n, h, w, Q, S, V = 16200, 30, 54, 200, 11, 22
k, s, e = synth.keys(g, n)
qk, qe = synth.query(g, h, w)
sim = orc.anisotropic_l2(k, s, qk.flatten(2), qe.flatten(2))[0].numpy()[:, :Q]      # N x Q
tiles = (n + 63) // 64
bounds = [tiles * i // S for i in range(S + 1)]
steps = (max(bounds[i + 1] - bounds[i] for i in range(S)) + 1) // 2


def run(mode, lag):
    appends = np.zeros(Q)
    b1 = np.full((V, Q), -np.inf)
    b2 = np.full((V, Q), -np.inf)
    hist = [np.full(Q, -np.inf)]
    kept = []
    for t in range(steps):
        pending = []
        for sp in range(S):
            for half in range(2):
                ti = bounds[sp] + 2 * t + half
                if ti >= bounds[sp + 1]:
                    continue
                sc = sim[ti * 64:(ti + 1) * 64]
                pending.append(sc)
                v = sp * 2 + half
                gm = sc.reshape(-1, 8, Q).max(1) if sc.shape[0] % 8 == 0 else sc
                for m in gm:
                    lo = np.minimum(b1[v], m)
                    b1[v] = np.maximum(b1[v], m)
                    b2[v] = np.maximum(b2[v], lo)
        if mode == 'min':
            cur = b2.min(0)
        elif mode == 'pair':
            cur = np.minimum(np.minimum(b1[0::2], b1[1::2]), np.maximum(b2[0::2], b2[1::2])).min(0)
        elif mode == 'quad':
            cur = np.array([grouped_bound(b1[:, q], b2[:, q]) for q in range(Q)])
        else:
            cur = np.maximum(np.sort(np.concatenate([b1, b2], 0), 0)[::-1][32], b2.min(0))
        hist.append(np.maximum(hist[-1], cur))
        thr = hist[max(1, len(hist) - 1 - lag)]      # the first tile waits for every split's first publication
        for sc in pending:
            appends += (sc >= thr[None, :]).sum(0)
            kept.append(np.where(sc >= thr[None, :], sc, -np.inf))
    left = (np.concatenate(kept, 0) >= hist[-1][None, :]).sum(0)
    return appends.mean(), left.mean()


for mode in ('min', 'pair', 'quad', 'exact'):
    for lag in (1, 2):
        a, c = run(mode, lag)
        print(f'{mode:5s} thresholds {lag} tile(s) behind: {a:6.1f} appends per query, {c:5.1f} candidates at the hand-off')
