"""CPU: the oracle restatement must reproduce the reference's own outputs (tests/golden)."""
import numpy as np
import pytest
import torch

from oracle import readout_oracle as orc
from tests.replay import load, t, lifecycle_config, replay_lifecycle

TOL = dict(rtol=1e-5, atol=1e-5)


def test_similarity_variants():
    z = load('util_cases.npz')
    mk, ms, qk, qe = (t(z[k]) for k in ('mk', 'ms', 'qk', 'qe'))
    torch.testing.assert_close(orc.anisotropic_l2(mk, ms, qk, qe), t(z['sim_aniso']), **TOL)
    torch.testing.assert_close(orc.anisotropic_l2(mk, ms, qk, None), t(z['sim_iso']), **TOL)
    torch.testing.assert_close(orc.anisotropic_l2(mk, None, qk, qe), t(z['sim_noshrink']), **TOL)
    torch.testing.assert_close(orc.anisotropic_l2(mk, None, qk, None), t(z['sim_plain']), **TOL)


def test_affinity_topk_and_dense():
    z = load('util_cases.npz')
    sim, k = t(z['sim_aniso']), int(z['top_k'])
    aff, usage = orc.affinity(sim, top_k=k, want_usage=True)
    torch.testing.assert_close(aff, t(z['aff_topk']), **TOL)
    torch.testing.assert_close(aff, t(z['aff_topk_inplace']), **TOL)
    torch.testing.assert_close(usage, t(z['usage_topk']), **TOL)
    assert int((aff != 0).sum(dim=1).min()) == k and int((aff != 0).sum(dim=1).max()) == k
    daff, dusage = orc.affinity(sim, top_k=None, want_usage=True)
    torch.testing.assert_close(daff, t(z['aff_dense']), **TOL)
    torch.testing.assert_close(dusage, t(z['usage_dense']), **TOL)
    mk, ms, qk, qe = (t(z[k]) for k in ('mk', 'ms', 'qk', 'qe'))
    torch.testing.assert_close(orc.dense_affinity(orc.anisotropic_l2(mk, ms, qk, qe)), t(z['aff_get_affinity']), **TOL)


def test_readout_5d():
    z = load('util_cases.npz')
    torch.testing.assert_close(orc.readout_5d(t(z['aff_for_readout']), t(z['mv'])), t(z['readout']), **TOL)


def bank_from_golden(z, prefix, count_usage):
    b = orc.Bank(count_usage=count_usage)
    b.key, b.shrinkage = t(z[prefix + 'key']), t(z[prefix + 'shrinkage'])
    if prefix + 'selection' in z:
        b.selection = t(z[prefix + 'selection'])
    b.values = [t(z[f'{prefix}value{g}']) for g in range(int(z[prefix + 'num_groups']))]
    if count_usage:
        b.use_count, b.life_count = t(z[prefix + 'use_count']), t(z[prefix + 'life_count'])
    return b


@pytest.mark.parametrize('name', ['w1', 'w2', 'l1', 'l3', 'nousage', 'nolongusage', 'k5'])
def test_match_cases(name):
    z = load('match_cases.npz')
    p = name + '/'
    top_k, en_long, en_long_usage, cv = (int(v) for v in z[p + 'cfg'])
    work = bank_from_golden(z, p + 'work_', count_usage=bool(en_long))
    long = bank_from_golden(z, p + 'long_', count_usage=bool(en_long_usage)) if (p + 'long_key') in z else None
    if long is None and en_long:
        long = orc.Bank(count_usage=bool(en_long_usage))
    r = orc.match(work, long, t(z[p + 'qk']), t(z[p + 'qe']), top_k, bool(en_long), bool(en_long_usage), cv)
    torch.testing.assert_close(r.readout, t(z[p + 'readout']), rtol=1e-4, atol=1e-5)
    if (p + 'work_use_after') in z:
        work.record_usage(r.work_usage)
        torch.testing.assert_close(work.use_count, t(z[p + 'work_use_after']), **TOL)
        torch.testing.assert_close(work.life_count, t(z[p + 'work_life_after']), **TOL)
    else:
        assert r.work_usage is None or not work.count_usage
    if (p + 'long_use_after') in z:
        long.record_usage(r.long_usage)
        torch.testing.assert_close(long.use_count, t(z[p + 'long_use_after']), **TOL)
        torch.testing.assert_close(long.life_count, t(z[p + 'long_life_after']), **TOL)
    # sparse-only path agrees with the dense one and maps back onto the key axis
    rs = orc.match(work, long, t(z[p + 'qk']), t(z[p + 'qe']), top_k, bool(en_long), bool(en_long_usage), cv,
                   sparse_only=True)
    for a, b in zip(r.indices, rs.indices):
        assert bool(orc.index_sets_equal(a, b).all())
    n_total = work.size + (long.size if long is not None else 0)
    for span, idx in zip(r.group_spans, r.indices):
        kidx = orc.group_key_index(span, idx)
        assert int(kidx.min()) >= 0 and int(kidx.max()) < n_total


@pytest.mark.parametrize('name', ['evict', 'groups'])
def test_lifecycle_replay(name):
    z = load(f'lifecycle_{name}.npz')
    mgr = orc.Readout(lifecycle_config(z))
    results = replay_lifecycle(z, mgr)
    assert len(results) > 20
    for i, got, want in results:
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5, msg=lambda m: f'event {i}: {m}')
    torch.testing.assert_close(mgr.work_mem.key, t(z['final/work_key']), **TOL)
    torch.testing.assert_close(mgr.work_mem.use_count, t(z['final/work_use_count']), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(mgr.long_mem.key, t(z['final/long_key']), **TOL)
    torch.testing.assert_close(mgr.long_mem.shrinkage, t(z['final/long_shrinkage']), rtol=1e-4, atol=1e-5)
    for g in range(int(z['final/long_num_groups'])):
        torch.testing.assert_close(mgr.long_mem.values[g], t(z[f'final/long_value{g}']), rtol=1e-4, atol=1e-5)


def test_parity_helpers():
    g = torch.Generator().manual_seed(3)
    sim = torch.randn(1, 100, 7, generator=g, dtype=torch.float64)
    gap = orc.topk_gap(sim, 5)
    assert gap.shape == (1, 7) and bool((gap >= 0).all())
    idx = torch.topk(sim, 5, dim=1).indices[0]
    perm = idx[torch.randperm(5, generator=g)]
    assert bool(orc.index_sets_equal(idx, perm).all())
    assert orc.rel_err(sim, sim) == 0.0


@pytest.mark.parametrize('name', ['small', 'xmem'])
def test_key_projection_oracle_vs_reference_module(name):
    """oracle.key_projection == the reference's KeyProjection module (tracker/model/modules.py:194-211) on the seeded
    parameters / input that oracle/gen_golden.py used."""
    from tests import synth
    z = load('keyproj_cases.npz')
    in_dim, h, w = (int(v) for v in z[f'{name}/shape'])
    g = torch.Generator().manual_seed(500 + in_dim)
    prm = synth.keyproj_params(g, in_dim)
    x = torch.randn(1, in_dim, h, w, generator=g)
    key, shrinkage, selection = orc.key_projection(x, **prm)
    torch.testing.assert_close(key, t(z[f'{name}/key']), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(shrinkage, t(z[f'{name}/shrinkage']), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(selection, t(z[f'{name}/selection']), rtol=1e-4, atol=1e-5)
    key_only = orc.key_projection(x, **prm, need_s=False, need_e=False)
    assert key_only[1] is None and key_only[2] is None
