"""CPU, world_size 2 over gloo: the host-side protocol of the N-sharded long-term readout
(shard bounds, per-owner candidate lists, all-to-all, merge, query-sliced readout, output all-gather) with an oracle-backed
compute backend injected in place of the CUDA kernels.  Result must equal the unsharded oracle readout."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import readout_oracle as orc
from tests import synth


class OracleBackend:
    """CPU stand-in for vos_e_sam_b200.sharded.CudaBackend (test infrastructure): the same calls, on the tensors the
    engine passes next to the raw addresses (collective exchange mode only -- there is no peer memory on the CPU)."""
    K = 32     # sharded.EXCH_K

    def load_keys(self, key, shrinkage):
        self.key, self.shrinkage = key, shrinkage

    def load_values(self, value):
        self.value = value.reshape(-1, value.shape[-1])
        return self.value.shape[0]

    def new_buffer(self, nbytes):
        return torch.zeros(nbytes, dtype=torch.uint8)

    def select_push(self, qk, qe, top_k, index_base, per, world, rank, dst_ptrs, flag_ptrs, seq, ticket_ptr, send=None,
                    segments=None, seg0_len=-1, index_base1=0):
        assert flag_ptrs is None and send is not None
        hw = qk.shape[1]
        score = torch.full((world * per, self.K), float('-inf'))
        index = torch.full((world * per, self.K), -1, dtype=torch.int32)
        if segments is None:
            key, shrinkage = self.key, self.shrinkage
        else:       # a sharded MemoryManager's call: this rank's candidate ranges of the [long-term | working] banks
            key = torch.cat([sg.key[:, sg.begin:sg.end] for sg in segments], -1).unsqueeze(0)
            shrinkage = torch.cat([sg.shrinkage[sg.begin:sg.end] for sg in segments], -1).view(1, 1, -1)
        if key.shape[-1] > 0:
            sim = orc.anisotropic_l2(key, shrinkage, qk.unsqueeze(0), qe.unsqueeze(0) if qe is not None else None)
            kk = min(top_k, sim.shape[1])
            v, i = torch.topk(sim[0], kk, dim=0)                       # kk x HW
            i = i.t()
            glob = i + index_base if seg0_len < 0 else torch.where(i < seg0_len, i + index_base, i - seg0_len + index_base1)
            score[:hw, :kk], index[:hw, :kk] = v.t(), glob.to(torch.int32)
        packed = torch.stack((score.view(torch.int32), index), dim=-1)           # [world * per][K][2] = 8-byte entries
        send.view(torch.int32).view(-1)[:packed.numel()].copy_(packed.flatten())

    def exchange_readout(self, lists_ptr, n_lists, list_stride, first_entry, flags_ptr, seq, status_ptr, n_q, top_k, rows,
                         n_total, out, lists=None, values=None):
        assert flags_ptr is None and lists is not None and first_entry == 0
        value = self.value if values is None else torch.cat([vs.shadow[:vs.count].t() for vs in values], -1)
        e = lists.view(torch.int32).view(-1)[:n_lists * list_stride * 2].view(n_lists, list_stride // self.K, self.K, 2)
        s = e[..., 0].contiguous().view(torch.float32)[:, :n_q].permute(1, 0, 2).reshape(n_q, -1)
        i = e[..., 1][:, :n_q].permute(1, 0, 2).reshape(n_q, -1).to(torch.int64)
        v, pos = torch.topk(s, top_k, dim=1)
        idx = torch.gather(i, 1, pos)
        w = torch.softmax(v, dim=1)
        picked = value[:, idx.clamp(min=0)]                            # rows x n_q x k
        out.copy_((picked * w.unsqueeze(0)).sum(-1))
        return out


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, result_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from importlib import import_module
    sharded = import_module('vos_e_sam_b200.sharded')
    g = torch.Generator().manual_seed(42)
    k, s, _ = synth.keys(g, n)
    v = torch.randn(2, 16, n, generator=g)
    qk, qe = synth.query(g, 5, 9)
    eng = sharded.ShardedLongTermReadout(dict(top_k=30), rank, world, 'cpu', backend=OracleBackend())
    eng.load_long_term(k, s, v)
    out = eng.match(qk, qe)
    mine = eng.match(qk, qe, gather=False)                     # this rank's query slice only
    q_lo, q_hi = eng.query_range(45)
    assert mine.shape == (32, q_hi - q_lo) and torch.equal(mine, out[:, q_lo:q_hi])
    # unsharded oracle
    sim = orc.anisotropic_l2(k, s, qk.flatten(2), qe.flatten(2))
    aff = orc.topk_affinity(sim, 30)
    want = torch.matmul(v.reshape(32, n), aff[0])
    err = orc.rel_err(out, want)
    torch.save(dict(err=err, shape=tuple(out.shape), bounds=(eng.lo, eng.hi)), f'{result_path}.{rank}')
    dist.destroy_process_group()


def _manager_worker(rank, world, port, result_path):
    """One object group of a sharded MemoryManager call: [long-term suffix | working suffix] candidate ranges cut into
    per-rank shards (plan_shard_ranges), two index bases, value segments on the global candidate axis."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from importlib import import_module
    sharded = import_module('vos_e_sam_b200.sharded')
    ops = import_module('vos_e_sam_b200.ops')
    g = torch.Generator().manual_seed(7)
    n_long, n_work, lb, wb = 300, 700, 129, 64          # the group sees long[129:300) and work[64:700)
    lk, ls, _ = synth.keys(g, n_long)
    wk, ws, _ = synth.keys(g, n_work)
    lv, wv = torch.randn(24, n_long - lb, generator=g), torch.randn(24, n_work - wb, generator=g)
    qk, qe = synth.query(g, 5, 9)
    ranges, sizes = [(lb, n_long), (wb, n_work)], [n_long, n_work]
    mine, base0, seg0_len, base1, share_ok = sharded.plan_shard_ranges(ranges, sizes, world, rank)
    segs = [ops.KeySegment(key=k[0], shrinkage=s.view(-1), image=None, begin=lo, end=hi)
            for (k, s), (lo, hi) in zip(((lk, ls), (wk, ws)), mine)]
    vals = [ops.ValueSegment(shadow=lv.t().contiguous(), first=0, count=n_long - lb),
            ops.ValueSegment(shadow=wv.t().contiguous(), first=n_long - lb, count=n_work - wb)]
    prob = sharded.ShardProblem(segs, base0, seg0_len, base1, vals, 24, (n_long - lb) + (n_work - wb), share_ok)
    eng = sharded.ShardedLongTermReadout(dict(top_k=30), rank, world, 'cpu', backend=OracleBackend())
    out = eng.match(qk, qe, problem=prob)
    mk = torch.cat([lk[:, :, lb:], wk[:, :, wb:]], -1)
    ms = torch.cat([ls[:, :, lb:], ws[:, :, wb:]], -1)
    aff = orc.topk_affinity(orc.anisotropic_l2(mk, ms, qk.flatten(2), qe.flatten(2)), 30)
    want = torch.matmul(torch.cat([lv, wv], -1), aff[0])
    torch.save(dict(err=orc.rel_err(out, want), shape=tuple(out.shape), mine=mine), f'{result_path}.{rank}')
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_manager_problem_gloo(tmp_path, world):
    port = _free_port()
    path = str(tmp_path / 'res')
    mp.spawn(_manager_worker, args=(world, port, path), nprocs=world, join=True)
    res = [torch.load(f'{path}.{r}') for r in range(world)]
    for r in res:
        assert r['shape'] == (24, 45) and r['err'] < 1e-5
    # the ranks' shards tile both candidate ranges
    for b, (begin, end) in enumerate([(129, 300), (64, 700)]):
        covered = sorted(sum((list(range(*r['mine'][b])) for r in res), []))
        assert covered == list(range(begin, end))


def test_plan_shard_ranges_maps_every_candidate_once():
    from importlib import import_module
    sh = import_module('vos_e_sam_b200.sharded')
    for (L, Lb, W, Wb, G) in [(1000, 0, 3240, 0, 2), (1000, 300, 3240, 1620, 4), (100, 0, 1620, 0, 8), (0, 0, 700, 100, 3)]:
        ranges = [(Lb, L), (Wb, W)]
        seen, any_empty = [], False
        for r in range(G):
            mine, b0, s0, b1, ok = sh.plan_shard_ranges(ranges, [L, W], G, r)
            n0, n1 = mine[0][1] - mine[0][0], mine[1][1] - mine[1][0]
            any_empty = any_empty or n0 + n1 == 0
            assert s0 == n0 and all(lo % 64 == 0 or lo in (Lb, Wb) or lo == hi for lo, hi in mine)
            seen += [i + b0 if i < s0 else i - s0 + b1 for i in range(n0 + n1)]
        assert sorted(seen) == list(range((L - Lb) + (W - Wb)))
        assert ok == (not any_empty)
    mine, b0, s0, b1, ok = sh.plan_shard_ranges([(100, 700)], [700], 3, 1)      # working memory only
    assert s0 == -1 and b0 == mine[0][0] - 100


@pytest.mark.parametrize('n', [1000, 70])   # 70: the second rank's shard is nearly empty (64-key alignment)
def test_sharded_protocol_world2_gloo(tmp_path, n):
    world, port = 2, _free_port()
    path = str(tmp_path / 'res')
    mp.spawn(_worker, args=(world, port, n, path), nprocs=world, join=True)
    res = [torch.load(f'{path}.{r}') for r in range(world)]
    for r in res:
        assert r['shape'] == (32, 45)
        assert r['err'] < 1e-5
    assert res[0]['bounds'][0] == 0 and res[0]['bounds'][1] == res[1]['bounds'][0] and res[1]['bounds'][1] == n


def test_partition_helpers():
    from importlib import import_module
    sh = import_module('vos_e_sam_b200.sharded')
    assert sh.shard_bounds(100_000, 8, 0) == (0, 12544) and sh.shard_bounds(100_000, 8, 7) == (87808, 100_000)
    covered = sorted(sum((list(range(*sh.shard_bounds(1000, 3, r))) for r in range(3)), []))
    assert covered == list(range(1000))
    assert sh.shard_bounds(64, 4, 2) == (64, 64)                              # empty shard
    assert sorted(sum((sh.partition_sequences(64, 8, r) for r in range(8)), [])) == list(range(64))
    assert sh.query_slices(8160, 8) == 1024 and sh.query_slices(1620, 8) == 208 and sh.query_slices(45, 2) == 32
