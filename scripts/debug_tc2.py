"""Reproduce test_match_memory_davis_shape_vs_oracle[tcgen05] with a poisoned caching allocator."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vos_e_sam_b200 as vos
from vos_e_sam_b200 import ops, _native as N
from oracle import readout_oracle as orc
from tests import synth
from tests.test_gpu_parity import build_manager

poison = [torch.full((256 * 2 ** 20,), 0xFF, dtype=torch.uint8, device='cuda') for _ in range(8)]
del poison
for path in ('simt', 'tcgen05', 'tcgen05'):
    g = torch.Generator().manual_seed(1234 + 1)
    m, ref = build_manager(vos, g, (30, 54), 8, 1, 512, value_dtype='fp32', path=path)
    qk, qe = synth.query(g, 30, 54)
    q2, e2 = qk.cuda().flatten(2)[0], qe.cuda().flatten(2)[0]
    seg = [m.work_mem.key_segment(0, m.work_mem.size)]
    s_si, i_si = ops.select_topk(q2, e2, seg, 30, path=N.PATH_SIMT)
    s_tc, i_tc = ops.select_topk(q2, e2, seg, 30, path=N.PATH_TCGEN05)
    same = (torch.sort(i_tc, 1).values == torch.sort(i_si, 1).values).all(1)
    print(path, 'select sets equal', int(same.sum()), '/', same.numel(), 'nan in tc scores', int(torch.isnan(s_tc).sum()))
    bad = (~same).nonzero().flatten()[:4].tolist()
    for q in bad:
        print('  q', q, 'tc', i_tc[q, :6].tolist(), [round(x, 3) for x in s_tc[q, :6].tolist()], 'simt', i_si[q, :6].tolist(), [round(x, 3) for x in s_si[q, :6].tolist()])
    got = m.match_memory(qk.cuda(), qe.cuda())
    want = ref.match_memory(qk, qe)
    print(path, 'match rel err', orc.rel_err(got.cpu(), want))

print('---- detail')
sim = ops.similarity_dense(m.work_mem._k.buf[0][:, :m.work_mem.size], m.work_mem._s.buf.view(-1)[:m.work_mem.size], q2, e2)  # N x HW
for q in (~same).nonzero().flatten().tolist():
    a, b = set(i_tc[q].tolist()), set(i_si[q].tolist())
    col = sim[:, q]
    srt = torch.sort(col, descending=True)
    rank = {int(i): r for r, i in enumerate(srt.indices[:64].tolist())}
    print('q', q, 'tile', q // 128, 'row', q % 128)
    print('   tc-only ', [(i, round(float(col[i]), 4), rank.get(i), i // 64) for i in sorted(a - b)])
    print('   simt-only', [(i, round(float(col[i]), 4), rank.get(i), i // 64) for i in sorted(b - a)])
    print('   30th/31st/32nd/33rd exact', [round(float(x), 4) for x in srt.values[28:34].tolist()])
