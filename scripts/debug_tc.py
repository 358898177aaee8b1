"""Scratch diagnostics for the tcgen05 select path (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vos_e_sam_b200 as vos
from vos_e_sam_b200 import ops, _native as N
from tests import synth

torch.manual_seed(0)
g = torch.Generator().manual_seed(1235)
h, w = 30, 54
store = vos.KeyValueMemoryStore(count_usage=True, value_dtype=torch.float32)
ks, ss = [], []
for f in range(8):
    k, s, e = synth.keys(g, h * w)
    ks.append(k); ss.append(s)
    store.add(k.cuda(), torch.zeros(1, 8, h * w, device='cuda'), s.cuda(), e.cuda(), [1])
n = store.size
# 1. image built incrementally == image packed in one go?
img2 = torch.zeros_like(store._image)
ops.pack_keys(store._k.buf[0], store._s.buf.view(-1), 0, n, img2, store._k.capacity)
torch.cuda.synchronize()
nb = ops.key_image_bytes(64, n)
full_tiles = (n // 64) * 34816
print('image bytes equal (whole tiles):', bool((store._image[:full_tiles] == img2[:full_tiles]).all()), 'n', n, 'cap', store._k.capacity)
qk, qe = synth.query(g, h, w)
q2, e2 = qk.cuda().flatten(2)[0], qe.cuda().flatten(2)[0]
seg = [store.key_segment(0, n)]
for trial in range(3):
    s_tc, i_tc = ops.select_topk(q2, e2, seg, 30, path=N.PATH_TCGEN05)
    s_si, i_si = ops.select_topk(q2, e2, seg, 30, path=N.PATH_SIMT)
    torch.cuda.synchronize()
    same = (torch.sort(i_tc, 1).values == torch.sort(i_si, 1).values).all(1)
    print(f'trial {trial}: queries with identical sets {int(same.sum())}/{same.numel()}; max score diff {float((s_tc - s_si).abs().max()):.3e}')
    bad = (~same).nonzero().flatten()[:5].tolist()
    for q in bad:
        a, b = set(i_tc[q].tolist()), set(i_si[q].tolist())
        print('  q', q, 'tc-only', sorted(a - b)[:8], 'simt-only', sorted(b - a)[:8], 'tc scores', [round(x, 3) for x in s_tc[q, :4].tolist()], 'simt', [round(x, 3) for x in s_si[q, :4].tolist()])
# poison the capacity tail with NaN to see whether garbage beyond `n` leaks
store._image[nb:].fill_(0xFF)
s_tc, i_tc = ops.select_topk(q2, e2, seg, 30, path=N.PATH_TCGEN05)
same = (torch.sort(i_tc, 1).values == torch.sort(i_si, 1).values).all(1)
print('after poisoning beyond the last tile:', int(same.sum()), '/', same.numel())
