"""Sizes of the candidate lists the tcgen05 selection hands to the merge (DAVIS shape), read straight out of the workspace
(layout: csrc/common.cuh carve_workspace).  Answers: how many candidates per query does the merge see, and how many of
them are above the final shared threshold / how close is that threshold to the true 33rd best score?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vos_e_sam_b200 as vos
from vos_e_sam_b200 import ops, _native as N
from tests import synth

g = torch.Generator().manual_seed(1)
n, h, w = 16200, 30, 54
hw = h * w
k, s, _ = synth.keys(g, n)
store = vos.KeyValueMemoryStore(False)
store.add(k.cuda(), [], s.cuda(), None, None)
qk, qe = synth.query(g, h, w)
q2, e2 = qk.cuda().flatten(2)[0], qe.cuda().flatten(2)[0]
seg = [store.key_segment(0, n)]
score, index = ops.select_topk(q2, e2, seg, 32, path=N.PATH_TCGEN05)
torch.cuda.synchronize()
ws = ops.workspace_for(q2.device, 64, hw)
r256 = lambda x: (x + 255) // 256 * 256
n_qt = (hw + 127) // 128
hw_pad = n_qt * 128
splits_cap = max(148 // n_qt, -(-4736 // hw), 4)
splits_cap = min(splits_cap, 32)
cap = splits_cap * 2
off = 256 + r256(n_qt * 69632)
pub_off = off
off += 2 * r256(cap * hw_pad * 8)
cand_off = off
off += r256(splits_cap * hw_pad * 120 * 8)
cnt = ws[off:off + cap * hw_pad * 4].view(torch.int32).view(cap, hw_pad)
splits = 148 // n_qt
c = cnt[:splits, :hw].float()
print(f'splits {splits}: candidates per (split, query): mean {c.mean():.1f} max {c.max():.0f}; per query: mean {c.sum(0).mean():.1f} '
      f'min {c.sum(0).min():.0f} max {c.sum(0).max():.0f}; lists longer than 32: {(c > 32).float().mean() * 100:.2f} %')
cand = ws[cand_off:cand_off + splits_cap * hw_pad * 120 * 8].view(torch.float32).view(splits_cap, hw_pad, 120, 2)[:splits, :hw, :, 0]
mask = torch.arange(120, device=cand.device).view(1, 1, 120) < cnt[:splits, :hw].unsqueeze(-1)
k32 = score[:, 31]                                  # true 32nd best per query
above = ((cand >= k32.view(1, hw, 1)) & mask).sum((0, 2)).float()
pub = ws[pub_off:pub_off + cap * hw_pad * 8].view(torch.float32).view(cap, hw_pad, 2)[:2 * splits, :hw, 0]
tau = pub.min(0).values
rank_tau = ((cand >= tau.view(1, hw, 1)) & mask).sum((0, 2)).float()
print(f'candidates at or above the true 32nd best: mean {above.mean():.1f}; at or above the final shared threshold: mean {rank_tau.mean():.1f} '
      f'(threshold below the 32nd best by {float((k32 - tau).mean()):.3f} on average, score spread of the top 32: {float((score[:, 0] - k32).mean()):.3f})')
