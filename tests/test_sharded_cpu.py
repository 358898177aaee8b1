"""CPU, world_size 2 over gloo: the host-side protocol of the N-sharded long-term readout
(shard bounds, candidate all-gather, merge, query-sliced readout, output all-gather) with an oracle-backed
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
    """CPU stand-in for vos_e_sam_b200.sharded.CudaBackend (test infrastructure)."""

    def load_keys(self, key, shrinkage):
        self.key, self.shrinkage = key, shrinkage

    def load_values(self, value):
        self.value = value.reshape(-1, value.shape[-1])
        return self.value.shape[0]

    def select(self, qk, qe, top_k, index_base):
        sim = orc.anisotropic_l2(self.key, self.shrinkage, qk.unsqueeze(0), qe.unsqueeze(0) if qe is not None else None)
        kk = min(top_k, sim.shape[1])
        v, i = torch.topk(sim[0], kk, dim=0)                       # kk x HW
        score = torch.full((sim.shape[2], top_k), float('-inf'))
        index = torch.full((sim.shape[2], top_k), -1, dtype=torch.int64)
        score[:, :kk], index[:, :kk] = v.t(), i.t() + index_base
        return score, index

    def merge(self, scores, indices):
        g, hw, k = scores.shape
        s = scores.permute(1, 0, 2).reshape(hw, g * k)
        i = indices.permute(1, 0, 2).reshape(hw, g * k)
        v, pos = torch.topk(s, k, dim=1)
        return v, torch.gather(i, 1, pos)

    def readout(self, score, index, rows, n_total, out):
        w = torch.softmax(score, dim=1)
        picked = self.value[:, index.clamp(min=0)]                   # rows x hw x k
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
    # unsharded oracle
    sim = orc.anisotropic_l2(k, s, qk.flatten(2), qe.flatten(2))
    aff = orc.topk_affinity(sim, 30)
    want = torch.matmul(v.reshape(32, n), aff[0])
    err = orc.rel_err(out, want)
    torch.save(dict(err=err, shape=tuple(out.shape), bounds=(eng.lo, eng.hi)), f'{result_path}.{rank}')
    dist.destroy_process_group()


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
