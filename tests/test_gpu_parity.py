"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden vectors
generated from the unmodified reference.

Parity rule (BASELINE.json north_star): top-k index SETS are identical wherever the fp64 gap between the
k-th and (k+1)-th similarity exceeds 1e-3; readout / usage agree within 1e-2 relative error (norm-wise,
``oracle.rel_err``).  With fp32 value storage the readout is held to 1e-4.
"""
import math
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import readout_oracle as orc
from tests import synth
from tests.replay import load, t, lifecycle_config, replay_lifecycle

pytestmark = pytest.mark.gpu

GAP = 1e-3
TOL_BF16 = 1e-2
TOL_F32 = 1e-4


@pytest.fixture(scope='module')
def vos():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    import vos_e_sam_b200 as v
    from vos_e_sam_b200 import ops, _native
    v.ops, v.N = ops, _native
    return v


def dev(x):
    return x.cuda() if x is not None else None


def oracle_sim64(mk, ms, qk, qe):
    f = lambda x: x.double() if x is not None else None
    return orc.anisotropic_l2(f(mk), f(ms), f(qk).flatten(2), f(qe).flatten(2) if qe is not None else None)


def check_selection(score, index, sim64, k, what):
    """score/index: HW x k from the device; sim64: 1 x N x HW fp64 oracle scores."""
    n, hw = sim64.shape[1:]
    kk = min(k, n)
    top = torch.topk(sim64, k=kk, dim=1)
    gap = orc.topk_gap(sim64, kk)[0]
    idx = index.cpu().t()[:kk]                                   # k x HW
    same = orc.index_sets_equal(idx, top.indices[0])
    decided = gap > GAP
    bad = decided & ~same
    assert not bool(bad.any()), f'{what}: {int(bad.sum())} of {int(decided.sum())} decided queries differ'
    # scores of the chosen candidates are the oracle scores of those candidates
    got = score.cpu().t()[:kk].double()
    want = torch.gather(sim64[0], 0, idx.clamp(min=0))
    assert float((got - want).abs().max()) < 2e-3, f'{what}: max score error {float((got - want).abs().max())}'
    if kk < k:
        assert bool((index.cpu()[:, kk:] == -1).all()) and bool(torch.isinf(score.cpu()[:, kk:]).all())
    return float(decided.float().mean())


# ------------------------------------------------------------------------------------------------
def test_umma_tile_matches_fp32(vos):
    """One 128 x 64 tcgen05 tile through the packed images (query operand staged in TMEM by tcgen05.cp, key operand
    from shared memory) == the fp32 similarity (descriptor / layout check)."""
    g = torch.Generator().manual_seed(7)
    mk, ms, _ = synth.keys(g, 64)
    qk, qe = synth.query(g, 8, 16)
    img_k = torch.zeros(vos.ops.key_image_bytes(64, 64), dtype=torch.uint8, device='cuda')
    vos.ops.pack_keys(dev(mk)[0], dev(ms).view(-1), 0, 64, img_k, 64)
    qbytes = vos.N.lib.vosmem_query_image_bytes(64, 128)
    img_q = torch.zeros(qbytes, dtype=torch.uint8, device='cuda')
    q2, e2 = dev(qk).flatten(2)[0].contiguous(), dev(qe).flatten(2)[0].contiguous()
    vos.N.check(vos.N.lib.vosmem_debug_pack_query(q2.data_ptr(), e2.data_ptr(), 64, 128, img_q.data_ptr(), 0), 'pack_query')
    out = torch.zeros((128, 64), dtype=torch.float32, device='cuda')
    vos.N.check(vos.N.lib.vosmem_debug_umma_tile(img_q.data_ptr(), img_k.data_ptr(), out.data_ptr(), 0), 'umma_tile')
    torch.cuda.synchronize()
    want = oracle_sim64(mk, ms, qk, qe)[0].t()                     # HW x N
    err = float((out.cpu().double() - want).abs().max())
    assert err < 1e-3, f'tcgen05 tile differs from the fp32 similarity by {err}'


@pytest.mark.parametrize('path', ['simt', 'tcgen05'])
@pytest.mark.parametrize('n,h,w,k', [(2000, 10, 30, 30), (700, 9, 15, 30), (64, 4, 8, 30), (20, 3, 5, 30),
                                      (5000, 12, 11, 5), (12960, 30, 54, 30)])
def test_select_topk_vs_oracle(vos, path, n, h, w, k):
    g = torch.Generator().manual_seed(1234 + n)
    mk, ms, _ = synth.keys(g, n)
    qk, qe = synth.query(g, h, w)
    store = vos.KeyValueMemoryStore(count_usage=False)
    store.add(dev(mk), [torch.zeros(1, 8, n, device='cuda')], dev(ms), None, None)
    seg = store.key_segment(0, n)
    p = {'simt': vos.N.PATH_SIMT, 'tcgen05': vos.N.PATH_TCGEN05}[path]
    score, index = vos.ops.select_topk(dev(qk).flatten(2)[0], dev(qe).flatten(2)[0], [seg], k, path=p)
    torch.cuda.synchronize()
    frac = check_selection(score, index, oracle_sim64(mk, ms, qk, qe), k, f'{path} N={n} HW={h * w}')
    assert frac > 0.5


@pytest.mark.parametrize('hw,splits', [(16000, 1), (8160, 2), (6000, 3), (4700, 4), (3700, 5), (2500, 7), (1620, 11),
                                       (500, 23),
                                       # R == 2, grouped thresholds (select_tc.cu refresh_grouped): 18 ... 32 virtual splits,
                                       # i.e. every way the groups of four (+ a pair) share out the 33 keys
                                       (2000, 9), (1700, 10), (1500, 12), (1400, 13), (1200, 14), (1100, 16)])
def test_select_tc_every_rank_variant_matches_simt(vos, hw, splits):
    """The tcgen05 kernel is instantiated per published rank R = ceil(33 / (2 * splits)); the query count picks the
    number of key splits (148 SMs / query tiles) and with it the variant.  Every variant must return the candidates
    of the exact fp32 SIMT kernel (sets equal wherever the SIMT k / k+1 gap is decided), with scores within 2e-3."""
    n = 6000
    g = torch.Generator().manual_seed(4321 + hw)
    mk, ms, _ = synth.keys(g, n)
    qk = torch.randn(1, 64, hw, generator=g)
    qe = torch.sigmoid(torch.randn(1, 64, hw, generator=g))
    store = vos.KeyValueMemoryStore(count_usage=False)
    store.add(dev(mk), [torch.zeros(1, 8, n, device='cuda')], dev(ms), None, None)
    seg = [store.key_segment(0, n)]
    q2, e2 = qk.cuda()[0], qe.cuda()[0]
    s_tc, i_tc = vos.ops.select_topk(q2, e2, seg, 30, path=vos.N.PATH_TCGEN05)
    s_si, i_si = vos.ops.select_topk(q2, e2, seg, 31, path=vos.N.PATH_SIMT)      # 31st score -> the k / k+1 gap
    torch.testing.assert_close(s_tc, s_si[:, :30], rtol=0, atol=2e-3)
    decided = (s_si[:, 29] - s_si[:, 30]) > 2 * GAP
    same = (torch.sort(i_tc, 1).values == torch.sort(i_si[:, :30], 1).values).all(1)
    assert float(decided.float().mean()) > 0.5
    assert not bool((decided & ~same).any()), f'{int((decided & ~same).sum())} decided queries differ (splits={splits})'


@pytest.mark.parametrize('path', ['simt', 'tcgen05'])
def test_select_redundant_video(vos, path):
    """Near-duplicate memory frames (static background): many near-ties around the k-th score."""
    g = torch.Generator().manual_seed(99)
    hw = 20 * 27
    mk, ms = synth.redundant_keys(g, 6, hw, 0.05)
    qk, qe = synth.query(g, 20, 27)
    store = vos.KeyValueMemoryStore(count_usage=False)
    store.add(dev(mk), [torch.zeros(1, 8, 6 * hw, device='cuda')], dev(ms), None, None)
    p = {'simt': vos.N.PATH_SIMT, 'tcgen05': vos.N.PATH_TCGEN05}[path]
    score, index = vos.ops.select_topk(dev(qk).flatten(2)[0], dev(qe).flatten(2)[0], [store.key_segment(0, 6 * hw)], 30,
                                       path=p)
    check_selection(score, index, oracle_sim64(mk, ms, qk, qe), 30, f'{path} redundant')


@pytest.mark.parametrize('n,h,w', [(40000, 40, 60), (65536, 68, 120)])
def test_select_ascending_stream_overflows_every_list(vos, n, h, w):
    """Worst case for the streaming selection: the shrinkage falls along the key axis, so every query's similarity RISES
    along the stream and nearly every tile brings new best candidates -- the candidate lists overflow over and over
    (cooperative cuts to the best 32) and, with group candidates, the per-list score logs wrap (flip_log: hundreds of
    appends per list when a CTA streams 512 tiles).  Results must still be the oracle's (checked on 256 sampled queries)."""
    g = torch.Generator().manual_seed(77 + n)
    mk, _, _ = synth.keys(g, n)
    ms = torch.linspace(60.0, 1.0, n).view(1, 1, n)
    qk, qe = synth.query(g, h, w)
    store = vos.KeyValueMemoryStore(count_usage=False)
    store.add(dev(mk), [torch.zeros(1, 8, n, device='cuda')], dev(ms), None, None)
    score, index = vos.ops.select_topk(dev(qk).flatten(2)[0], dev(qe).flatten(2)[0], [store.key_segment(0, n)], 30,
                                       path=vos.N.PATH_TCGEN05)
    torch.cuda.synchronize()
    cols = torch.randperm(h * w, generator=g)[:256]
    sim64 = oracle_sim64(mk, ms, qk.flatten(2)[:, :, cols].unsqueeze(-1), qe.flatten(2)[:, :, cols].unsqueeze(-1))
    index, score = index.cpu()[cols], score.cpu()[cols]
    # scores reach ~ -60 * 64 / 8: the error of the bf16 hi / lo split is relative, so the gap that decides a query and
    # the score tolerance scale with the magnitude of its k-th score
    top = torch.topk(sim64, k=31, dim=1)
    scale = top.values[0, 29].abs().clamp(min=1.0)
    decided = (top.values[0, 29] - top.values[0, 30]) > GAP * scale
    same = orc.index_sets_equal(index.t(), top.indices[0, :30])
    assert float(decided.float().mean()) > 0.5
    assert not bool((decided & ~same).any()), f'{int((decided & ~same).sum())} of {int(decided.sum())} decided queries differ'
    got = score.t().double()
    want = torch.gather(sim64[0], 0, index.t().clamp(min=0))
    assert float(((got - want).abs() / scale).max()) < 2e-3


def test_group_candidate_mode_in_a_child_process(vos):
    """VOSMEM_TC_CANDIDATES=groups (8-key group candidates + expansion in the merge, select_tc.cu / merge.cuh) is read
    once per process: run the selection, tie, overflow, graph-replay and full-size match tests again in a child."""
    if os.environ.get('VOSMEM_TEST_CHILD'):
        pytest.skip('already the child')
    env = dict(os.environ, VOSMEM_TC_CANDIDATES='groups', VOSMEM_TEST_CHILD='1')
    pick = ('select_topk_vs_oracle or ascending or every_rank or two_segments or redundant or exact_ties or '
            'davis_shape or davis5_full or replayed or batch_equals or sharded_engine')
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.abspath(__file__), '-x', '-q', '-m', 'gpu', '-k', pick,
                        '-p', 'no:cacheprovider'], env=env, capture_output=True, text=True, timeout=900,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert ' passed' in r.stdout


@pytest.mark.parametrize('path', ['simt', 'tcgen05'])
def test_select_two_segments_with_suffix_and_no_selection(vos, path):
    """[long suffix | work suffix] candidate axis with unaligned range starts, isotropic query (qe=None)."""
    g = torch.Generator().manual_seed(5)
    lk, ls, _ = synth.keys(g, 300)
    wk, ws, _ = synth.keys(g, 1000)
    qk, _ = synth.query(g, 7, 19)
    long, work = vos.KeyValueMemoryStore(False), vos.KeyValueMemoryStore(False)
    long.add(dev(lk), [torch.zeros(1, 8, 300, device='cuda')], dev(ls), None, None)
    work.add(dev(wk), [torch.zeros(1, 8, 1000, device='cuda')], dev(ws), None, None)
    segs = [long.key_segment(300 - 171, 300), work.key_segment(1000 - 333, 1000)]
    p = {'simt': vos.N.PATH_SIMT, 'tcgen05': vos.N.PATH_TCGEN05}[path]
    score, index = vos.ops.select_topk(dev(qk).flatten(2)[0], None, segs, 30, path=p)
    mk = torch.cat([lk[:, :, -171:], wk[:, :, -333:]], -1)
    ms = torch.cat([ls[:, :, -171:], ws[:, :, -333:]], -1)
    check_selection(score, index, oracle_sim64(mk, ms, qk, None), 30, f'{path} two segments')


def test_merge_topk_of_shards_equals_global(vos):
    """Local top-k per N-shard + merge == top-k over the whole bank (the cross-GPU exchange, on one GPU)."""
    g = torch.Generator().manual_seed(11)
    n, shards, k = 4000, 4, 30
    mk, ms, _ = synth.keys(g, n)
    qk, qe = synth.query(g, 10, 13)
    store = vos.KeyValueMemoryStore(False)
    store.add(dev(mk), [torch.zeros(1, 8, n, device='cuda')], dev(ms), None, None)
    q2, e2 = dev(qk).flatten(2)[0], dev(qe).flatten(2)[0]
    per = n // shards
    parts = [vos.ops.select_topk(q2, e2, [store.key_segment(r * per, (r + 1) * per)], k, index_base=r * per)
             for r in range(shards)]
    score, index = vos.ops.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    check_selection(score, index, oracle_sim64(mk, ms, qk, qe), k, 'sharded merge')


# ------------------------------------------------------------------------------------------------
def test_memory_util_twins_vs_golden(vos):
    z = load('util_cases.npz')
    mk, ms, qk, qe = (t(z[k], 'cuda') for k in ('mk', 'ms', 'qk', 'qe'))
    tol = dict(rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(vos.get_similarity(mk, ms, qk, qe).cpu(), t(z['sim_aniso']), **tol)
    torch.testing.assert_close(vos.get_similarity(mk, ms, qk, None).cpu(), t(z['sim_iso']), **tol)
    torch.testing.assert_close(vos.get_similarity(mk, None, qk, qe).cpu(), t(z['sim_noshrink']), **tol)
    torch.testing.assert_close(vos.get_similarity(mk, None, qk, None).cpu(), t(z['sim_plain']), **tol)
    sim, k = t(z['sim_aniso'], 'cuda'), int(z['top_k'])
    aff, usage = vos.do_softmax(sim.clone(), top_k=k, inplace=False, return_usage=True)
    torch.testing.assert_close(aff.cpu(), t(z['aff_topk']), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(usage.cpu(), t(z['usage_topk']), rtol=1e-4, atol=1e-6)
    s2 = sim.clone()
    aff_in = vos.do_softmax(s2, top_k=k, inplace=True)
    assert aff_in.data_ptr() == s2.data_ptr()
    torch.testing.assert_close(aff_in.cpu(), t(z['aff_topk']), rtol=1e-4, atol=1e-6)
    daff, dusage = vos.do_softmax(sim.clone(), top_k=None, return_usage=True)
    torch.testing.assert_close(daff.cpu(), t(z['aff_dense']), rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(dusage.cpu(), t(z['usage_dense']), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(vos.get_affinity(mk, ms, qk, qe).cpu(), t(z['aff_get_affinity']), rtol=1e-3, atol=1e-7)
    torch.testing.assert_close(vos.readout(t(z['aff_for_readout'], 'cuda'), t(z['mv'], 'cuda')).cpu(), t(z['readout']),
                               rtol=1e-4, atol=1e-5)
    with pytest.raises(RuntimeError):
        vos.do_softmax(sim[:, :10].contiguous(), top_k=30)          # torch.topk raises in the reference too
    with pytest.raises(RuntimeError):
        vos.get_similarity(mk.cpu(), ms.cpu(), qk.cpu(), qe.cpu())  # no CPU path


def manager_from_golden(vos, z, p, value_dtype, path='auto'):
    top_k, en_long, en_long_usage, cv = (int(v) for v in z[p + 'cfg'])
    cfg = dict(hidden_dim=8, top_k=top_k, enable_long_term=bool(en_long), enable_long_term_count_usage=bool(en_long_usage),
               max_mid_term_frames=10, min_mid_term_frames=5, num_prototypes=128, max_long_term_elements=10000,
               vosmem_value_dtype=value_dtype, vosmem_path=path)
    m = vos.MemoryManager(cfg)
    m.CV = cv

    def fill(store, prefix):
        key, shr = t(z[prefix + 'key'], 'cuda'), t(z[prefix + 'shrinkage'], 'cuda')
        sel = t(z[prefix + 'selection'], 'cuda') if (prefix + 'selection') in z else None
        vals = [t(z[f'{prefix}value{g}'], 'cuda') for g in range(int(z[prefix + 'num_groups']))]
        # banks with groups of different extents cannot be built by one add(); add the key prefix first
        n = key.shape[-1]
        lens = [v.shape[-1] for v in vals]
        cuts = sorted(set([0] + [n - l for l in lens] + [n]))
        for a, b in zip(cuts[:-1], cuts[1:]):
            part = [v[:, :, a - (n - l):b - (n - l)] if a >= n - l else None for v, l in zip(vals, lens)]
            store.add(key[:, :, a:b], part, shr[:, :, a:b], sel[:, :, a:b] if sel is not None else None, None)
        if store.count_usage:
            store._use.view().copy_(t(z[prefix + 'use_count'], 'cuda'))
            store._life.view().copy_(t(z[prefix + 'life_count'], 'cuda'))

    fill(m.work_mem, p + 'work_')
    if (p + 'long_key') in z:
        fill(m.long_mem, p + 'long_')
    return m


@pytest.mark.parametrize('value_dtype,tol', [('fp32', TOL_F32), ('bf16', TOL_BF16)])
@pytest.mark.parametrize('path', ['simt', 'tcgen05'])
@pytest.mark.parametrize('name', ['w1', 'w2', 'l1', 'l3', 'nousage', 'nolongusage', 'k5'])
def test_match_memory_vs_reference_golden(vos, name, path, value_dtype, tol):
    """MemoryManager.match_memory on hand-built banks == what the unmodified reference returned."""
    z = load('match_cases.npz')
    p = name + '/'
    m = manager_from_golden(vos, z, p, value_dtype, path)
    got = m.match_memory(t(z[p + 'qk'], 'cuda'), t(z[p + 'qe'], 'cuda'))
    torch.cuda.synchronize()
    want = t(z[p + 'readout'])
    assert got.shape == want.shape
    assert orc.rel_err(got.cpu(), want) < tol
    if (p + 'work_use_after') in z:
        assert orc.rel_err(m.work_mem.use_count.cpu(), t(z[p + 'work_use_after'])) < 1e-3
        torch.testing.assert_close(m.work_mem.life_count.cpu(), t(z[p + 'work_life_after']), rtol=1e-5, atol=1e-5)
    if (p + 'long_use_after') in z:
        assert orc.rel_err(m.long_mem.use_count.cpu(), t(z[p + 'long_use_after'])) < 1e-3
        torch.testing.assert_close(m.long_mem.life_count.cpu(), t(z[p + 'long_life_after']), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('value_dtype,tol', [('fp32', 2e-3), ('bf16', TOL_BF16)])
@pytest.mark.parametrize('name', ['evict', 'groups'])
def test_lifecycle_replay_vs_reference(vos, name, value_dtype, tol):
    """Every MemoryManager call the reference InferenceCore made over a short video (growth, consolidation
    into prototypes, least-used eviction, a second object group appearing) replayed on the CUDA path."""
    z = load(f'lifecycle_{name}.npz')
    cfg = lifecycle_config(z)
    cfg['vosmem_value_dtype'] = value_dtype
    m = vos.MemoryManager(cfg)
    results = replay_lifecycle(z, m, device='cuda')
    assert len(results) > 20
    for i, got, want in results:
        assert got.shape == want.shape
        assert orc.rel_err(got.cpu(), want) < tol, f'event {i}'
    torch.testing.assert_close(m.work_mem.key.cpu(), t(z['final/work_key']), rtol=1e-5, atol=1e-5)
    assert orc.rel_err(m.work_mem.use_count.cpu(), t(z['final/work_use_count'])) < 5e-3
    assert m.long_mem.key.shape == t(z['final/long_key']).shape


# ------------------------------------------------------------------------------------------------
def build_manager(vos, gen, hw_shape, n_frames, n_obj, cv, n_long=0, value_dtype='bf16', path='auto', top_k=30):
    h, w = hw_shape
    cfg = dict(hidden_dim=64, top_k=top_k, enable_long_term=True, enable_long_term_count_usage=True,
               max_mid_term_frames=10 ** 6, min_mid_term_frames=5, num_prototypes=128, max_long_term_elements=10 ** 7,
               vosmem_value_dtype=value_dtype, vosmem_path=path)
    m = vos.MemoryManager(cfg)
    ref = orc.Readout({k: v for k, v in cfg.items() if not k.startswith('vosmem')})
    for _ in range(n_frames):
        k, s, e = synth.keys(gen, h * w)
        v = torch.randn(1, n_obj, cv, h, w, generator=gen)
        args = (k.view(1, -1, h, w), s.view(1, 1, h, w), v, list(range(1, n_obj + 1)))
        ref.add_memory(*args, selection=e.view(1, -1, h, w))
        m.add_memory(*(a.cuda() if isinstance(a, torch.Tensor) else a for a in args), selection=e.view(1, -1, h, w).cuda())
    if n_long:
        k, s, _ = synth.keys(gen, n_long)
        v = torch.randn(n_obj, cv, n_long, generator=gen)
        ref.long_mem.append(k, [v], s, None, None)
        m.long_mem.add(k.cuda(), [v.cuda()], s.cuda(), None, None)
    return m, ref


def readout_from_indices(sim64, values2d, idx):
    """fp64 readout of the candidates the device chose: softmax of the oracle scores at idx (HW x k) times values."""
    sc = torch.gather(sim64[0].t(), 1, idx)                           # HW x k
    w = torch.softmax(sc, dim=1)
    picked = values2d.double()[:, idx]                                 # rows x HW x k
    return (picked * w.unsqueeze(0)).sum(-1)


@pytest.mark.parametrize('path', ['simt', 'tcgen05'])
def test_match_memory_davis_shape_vs_oracle(vos, path):
    """cfg-1 (BASELINE.json configs[0]): 8 frames x 1620 tokens, 1 object, CV=512, top-30 -- against the oracle.

    Queries whose k-th / (k+1)-th fp64 similarity gap is below 1e-3 are "undecided" under the parity rule (either
    candidate is a correct top-k member, and they carry a few percent of softmax weight each); for those the
    readout is checked against an fp64 recomputation from the candidates the device picked."""
    g = torch.Generator().manual_seed(1234 + 1)
    m, ref = build_manager(vos, g, (30, 54), 8, 1, 512, value_dtype='fp32', path=path)
    qk, qe = synth.query(g, 30, 54)
    got = m.match_memory(qk.cuda(), qe.cuda()).cpu().view(512, 1620)
    want = ref.match_memory(qk, qe).view(512, 1620)
    sim64 = oracle_sim64(ref.work_mem.key, ref.work_mem.shrinkage, qk, qe)
    decided = orc.topk_gap(sim64, 30)[0] > GAP
    assert float(decided.float().mean()) > 0.8
    assert orc.rel_err(got[:, decided], want[:, decided]) < TOL_F32
    p = {'simt': vos.N.PATH_SIMT, 'tcgen05': vos.N.PATH_TCGEN05}[path]
    _, idx = vos.ops.select_topk(qk.cuda().flatten(2)[0], qe.cuda().flatten(2)[0],
                                 [m.work_mem.key_segment(0, m.work_mem.size)], 30, path=p)
    und = (~decided).nonzero().flatten()
    redo = readout_from_indices(sim64[:, :, und], ref.work_mem.values[0][0], idx.cpu()[und])
    assert orc.rel_err(got[:, und], redo) < TOL_F32
    assert orc.rel_err(m.work_mem.use_count.cpu(), ref.work_mem.use_count) < 1e-2
    assert abs(float(m.work_mem.use_count.sum()) - 1620) < 0.1       # each call adds exactly HW of usage mass


def test_match_memory_long_term_multi_object_vs_oracle(vos):
    """Long-term prototypes + working memory, 3 objects, bf16 value shadow."""
    g = torch.Generator().manual_seed(77)
    m, ref = build_manager(vos, g, (15, 20), 6, 3, 128, n_long=1000, value_dtype='bf16')
    qk, qe = synth.query(g, 15, 20)
    for _ in range(2):
        got = m.match_memory(qk.cuda(), qe.cuda())
        want = ref.match_memory(qk, qe)
        assert orc.rel_err(got.cpu(), want) < TOL_BF16
    assert orc.rel_err(m.long_mem.use_count.cpu(), ref.long_mem.use_count) < 1e-3
    assert orc.rel_err(m.work_mem.use_count.cpu(), ref.work_mem.use_count) < 1e-3
    torch.testing.assert_close(m.long_mem.life_count.cpu(), ref.long_mem.life_count)


def test_long_video_shape_vs_oracle(vos):
    """cfg-3 (BASELINE.json configs[2]) at full size: 10 000 long-term elements + 9 working frames x 1620 tokens, one
    object, long-term usage counted -- readout and both usage counters against the oracle."""
    g = torch.Generator().manual_seed(1234 + 3)
    m, ref = build_manager(vos, g, (30, 54), 9, 1, 512, n_long=10_000, value_dtype='bf16')
    qk, qe = synth.query(g, 30, 54)
    got = m.match_memory(qk.cuda(), qe.cuda())
    want = ref.match_memory(qk, qe)
    assert got.shape == (1, 512, 30, 54)
    # parity rule: queries whose k / k+1 fp64 gap is below 1e-3 may legitimately pick either candidate
    mk = torch.cat([ref.long_mem.key, ref.work_mem.key], -1)
    ms = torch.cat([ref.long_mem.shrinkage, ref.work_mem.shrinkage], -1)
    decided = orc.topk_gap(oracle_sim64(mk, ms, qk, qe), 30)[0] > GAP
    assert float(decided.float().mean()) > 0.8
    assert orc.rel_err(got.cpu().view(512, -1)[:, decided], want.view(512, -1)[:, decided]) < TOL_BF16
    # an undecided query moves at most one survivor's weight (a few percent of HW's unit mass) between two keys
    assert orc.rel_err(m.long_mem.use_count.cpu(), ref.long_mem.use_count) < 5e-2
    assert orc.rel_err(m.work_mem.use_count.cpu(), ref.work_mem.use_count) < 5e-2
    assert abs(float(m.long_mem.use_count.sum() + m.work_mem.use_count.sum()) - 1620) < 0.2
    torch.testing.assert_close(m.long_mem.life_count.cpu(), ref.long_mem.life_count)


def test_lvos_shape_tc_matches_simt(vos):
    """cfg-4 shape (BASELINE.json configs[3]) on one GPU: 100 000 keys x 8160 queries.  The N x HW matrix would be
    3.3 GB, so the check is size-independent: the tcgen05 selection against the exact fp32 SIMT kernel (scores within
    2e-3, index sets equal wherever the SIMT k / k+1 gap is decided), and every softmax row of the readout sums to 1."""
    g = torch.Generator().manual_seed(1234 + 4)
    n, h, w = 100_000, 68, 120
    mk, ms, _ = synth.keys(g, n)
    qk, qe = synth.query(g, h, w)
    store = vos.KeyValueMemoryStore(count_usage=False)
    store.add(dev(mk), [torch.zeros(1, 8, n, device='cuda')], dev(ms), None, None)
    seg = [store.key_segment(0, n)]
    q2, e2 = dev(qk).flatten(2)[0], dev(qe).flatten(2)[0]
    s_tc, i_tc = vos.ops.select_topk(q2, e2, seg, 30, path=vos.N.PATH_TCGEN05)
    s_si, i_si = vos.ops.select_topk(q2, e2, seg, 31, path=vos.N.PATH_SIMT)
    torch.testing.assert_close(s_tc, s_si[:, :30], rtol=0, atol=2e-3)
    decided = (s_si[:, 29] - s_si[:, 30]) > 2 * GAP
    same = (torch.sort(i_tc, 1).values == torch.sort(i_si[:, :30], 1).values).all(1)
    assert float(decided.float().mean()) > 0.5
    assert not bool((decided & ~same).any()), f'{int((decided & ~same).sum())} decided queries differ'
    shadow = torch.randn(n, 64, device='cuda').to(torch.bfloat16)
    vals = [vos.ops.ValueSegment(shadow=shadow, first=0, count=n)]
    out, wgt = vos.ops.softmax_readout(s_tc, i_tc, vals, 64, want_weight=True)
    torch.testing.assert_close(wgt.sum(1), torch.ones(h * w, device='cuda'), rtol=1e-5, atol=1e-5)
    assert bool(torch.isfinite(out).all())


def test_full_size_properties(vos):
    """cfg-2 shape (5 objects, 10 frames x 1620 tokens): properties that hold at any size --
    SIMT and tcgen05 paths agree, the softmax weights of every query sum to 1, usage mass == HW,
    the readout is a convex combination of values (bounded by their extrema)."""
    g = torch.Generator().manual_seed(1234 + 2)
    m, _ = build_manager(vos, g, (30, 54), 10, 5, 512, value_dtype='bf16')
    qk, qe = synth.query(g, 30, 54)
    q2, e2 = qk.cuda().flatten(2)[0], qe.cuda().flatten(2)[0]
    seg = [m.work_mem.key_segment(0, m.work_mem.size)]
    s_tc, i_tc = vos.ops.select_topk(q2, e2, seg, 30, path=vos.N.PATH_TCGEN05)
    s_si, i_si = vos.ops.select_topk(q2, e2, seg, 30, path=vos.N.PATH_SIMT)
    torch.testing.assert_close(s_tc, s_si, rtol=0, atol=2e-3)
    gap = (s_si[:, -2] - s_si[:, -1])                               # crude decidedness proxy: 29th vs 30th
    same = (torch.sort(i_tc, 1).values == torch.sort(i_si, 1).values).all(1)
    assert int((~same).sum()) <= int((gap < 2e-3).sum())
    vals = [m.work_mem.value_segment(0, 0, with_usage=False)]
    out, wgt = vos.ops.softmax_readout(s_tc, i_tc, vals, m.work_mem.group_rows(0), want_weight=True)
    torch.testing.assert_close(wgt.sum(1), torch.ones(1620, device='cuda'), rtol=1e-5, atol=1e-5)
    before = float(m.work_mem.use_count.sum())
    r = m.match_memory(qk.cuda(), qe.cuda())
    assert r.shape == (5, 512, 30, 54)
    assert abs(float(m.work_mem.use_count.sum()) - before - 1620) < 0.2
    v = m.work_mem.value[0]
    assert float(r.max()) <= float(v.max()) + 1e-2 and float(r.min()) >= float(v.min()) - 1e-2
    torch.testing.assert_close(r.view(2560, 1620), out, rtol=1e-5, atol=1e-5)


def test_davis5_full_size_vs_oracle(vos):
    """cfg-2, the BENCHMARKED configuration (BASELINE.json configs[1]: 5 objects, 10 frames x 1620 tokens, bf16 value
    shadow) at full size against the oracle: readout of the decided queries, readout of the undecided ones recomputed
    in fp64 from the candidates the device picked, and the usage counters."""
    g = torch.Generator().manual_seed(1234 + 2)
    m, ref = build_manager(vos, g, (30, 54), 10, 5, 512, value_dtype='bf16')
    qk, qe = synth.query(g, 30, 54)
    got = m.match_memory(qk.cuda(), qe.cuda()).cpu().view(2560, 1620)
    want = ref.match_memory(qk, qe).view(2560, 1620)
    sim64 = oracle_sim64(ref.work_mem.key, ref.work_mem.shrinkage, qk, qe)
    decided = orc.topk_gap(sim64, 30)[0] > GAP
    assert float(decided.float().mean()) > 0.8
    assert orc.rel_err(got[:, decided], want[:, decided]) < TOL_BF16
    _, idx = vos.ops.select_topk(qk.cuda().flatten(2)[0], qe.cuda().flatten(2)[0],
                                 [m.work_mem.key_segment(0, m.work_mem.size)], 30)
    und = (~decided).nonzero().flatten()
    redo = readout_from_indices(sim64[:, :, und], ref.work_mem.values[0].reshape(2560, -1), idx.cpu()[und])
    assert orc.rel_err(got[:, und], redo) < TOL_BF16
    assert orc.rel_err(m.work_mem.use_count.cpu(), ref.work_mem.use_count) < 1e-2
    assert abs(float(m.work_mem.use_count.sum()) - 1620) < 0.2


def test_lvos_shape_readout_sampled_columns_vs_fp64(vos):
    """cfg-4 shape (100 000 keys x 8160 queries) through vosmem_match: 512 sampled query columns of the readout
    against an fp64 evaluation of memory_util.py:7-65 + the value product on those columns (the full N x HW
    matrix would be 3.3 GB; a column of it is independent of the others)."""
    g = torch.Generator().manual_seed(1234 + 4)
    n, h, w, rows = 100_000, 68, 120, 64
    mk, ms, _ = synth.keys(g, n)
    qk, qe = synth.query(g, h, w)
    val = torch.randn(1, rows, n, generator=g)
    store = vos.KeyValueMemoryStore(count_usage=False, value_dtype=torch.float32)
    store.add(dev(mk), [dev(val)], dev(ms), None, None)
    out = vos.ops.match(dev(qk).flatten(2)[0], dev(qe).flatten(2)[0], [store.key_segment(0, n)],
                        [store.value_segment(0, 0, with_usage=False)], rows, 30).cpu()
    cols = torch.randperm(h * w, generator=g)[:512]
    sim64 = oracle_sim64(mk, ms, qk.flatten(2)[:, :, cols], qe.flatten(2)[:, :, cols])        # 1 x N x 512
    decided = orc.topk_gap(sim64, 30)[0] > GAP
    assert float(decided.float().mean()) > 0.8
    want = torch.matmul(val[0].double(), orc.topk_affinity(sim64, 30)[0])                     # rows x 512
    got = out[:, cols]
    assert orc.rel_err(got[:, decided], want[:, decided].float()) < TOL_F32


@pytest.mark.parametrize('path', ['simt', 'tcgen05'])
def test_exact_ties_duplicated_frames(vos, path):
    """Exactly duplicated memory frames (identical keys, shrinkage and values): every score occurs six times.  The
    reference's torch.topk always returns k candidates; here too -- k distinct valid indices per query, the top-k SCORE
    multiset of the oracle, and (duplicates carry identical values, so the result is unique) the oracle's readout."""
    g = torch.Generator().manual_seed(17)
    h, w, frames, rows = 12, 25, 6, 32
    hw = h * w
    k1, s1, _ = synth.keys(g, hw)
    v1 = torch.randn(1, rows, hw, generator=g)
    mk, ms, val = k1.repeat(1, 1, frames), s1.repeat(1, 1, frames), v1.repeat(1, 1, frames)
    n = frames * hw
    qk, qe = synth.query(g, h, w)
    store = vos.KeyValueMemoryStore(count_usage=False, value_dtype=torch.float32)
    store.add(dev(mk), [dev(val)], dev(ms), None, None)
    p = {'simt': vos.N.PATH_SIMT, 'tcgen05': vos.N.PATH_TCGEN05}[path]
    q2, e2 = dev(qk).flatten(2)[0], dev(qe).flatten(2)[0]
    score, index = vos.ops.select_topk(q2, e2, [store.key_segment(0, n)], 30, path=p)
    index, score = index.cpu(), score.cpu()
    assert bool((index >= 0).all()) and bool((index < n).all())
    assert all(len(set(r.tolist())) == 30 for r in index), 'fewer than top_k distinct survivors'
    sim64 = oracle_sim64(mk, ms, qk, qe)
    want_scores = torch.topk(sim64[0], 30, dim=0).values.t()                                  # HW x 30, descending
    assert float((score.double() - want_scores).abs().max()) < 2e-3
    out = vos.ops.match(q2, e2, [store.key_segment(0, n)], [store.value_segment(0, 0, with_usage=False)], rows, 30,
                        path=p).cpu()
    # the k-th / (k+1)-th scores tie exactly for most queries, so "decided" is about distinct VALUES: compare against
    # the fp64 readout, which is unique because tied keys have identical values
    want = torch.matmul(val[0].double(), orc.topk_affinity(sim64, 30)[0])
    gap = orc.topk_gap(sim64, 30)[0]
    ok = (gap > GAP) | (gap == 0)
    assert orc.rel_err(out[:, ok], want[:, ok].float()) < 1e-3


@pytest.mark.parametrize('path', ['simt', 'tcgen05'])
def test_exact_ties_all_equal_bank(vos, path):
    """An all-equal key bank (uniform / black frames): every key scores the same for a query.  k distinct valid
    survivors, uniform softmax weights, and the readout equals the mean of the chosen value rows."""
    g = torch.Generator().manual_seed(19)
    n, h, w, rows = 2000, 9, 16, 16
    k1, s1, _ = synth.keys(g, 1)
    mk, ms = k1.repeat(1, 1, n), s1.repeat(1, 1, n)
    val = torch.randn(1, rows, n, generator=g)
    qk, qe = synth.query(g, h, w)
    store = vos.KeyValueMemoryStore(count_usage=False, value_dtype=torch.float32)
    store.add(dev(mk), [dev(val)], dev(ms), None, None)
    p = {'simt': vos.N.PATH_SIMT, 'tcgen05': vos.N.PATH_TCGEN05}[path]
    q2, e2 = dev(qk).flatten(2)[0], dev(qe).flatten(2)[0]
    score, index = vos.ops.select_topk(q2, e2, [store.key_segment(0, n)], 30, path=p)
    vals = [store.value_segment(0, 0, with_usage=False)]
    out, wgt = vos.ops.softmax_readout(score, index, vals, rows, want_weight=True)
    index = index.cpu()
    assert bool((index >= 0).all()) and bool((index < n).all())
    assert all(len(set(r.tolist())) == 30 for r in index), 'fewer than top_k distinct survivors'
    torch.testing.assert_close(wgt.cpu(), torch.full((h * w, 30), 1 / 30.), rtol=1e-3, atol=1e-4)
    want = val[0][:, index].mean(-1)                                                          # rows x HW
    assert orc.rel_err(out.cpu(), want) < 1e-3


def test_one_captured_graph_replayed_with_new_queries(vos):
    """the path must be graph-capturable.  ONE captured graph of vosmem_match, replayed with new
    query contents in the same buffers (and once after the memory grew in place), must equal the oracle every time:
    the launch epoch that validates published thresholds is a device-side word, not a host value baked into the graph."""
    g = torch.Generator().manual_seed(1234 + 9)
    m, ref = build_manager(vos, g, (30, 54), 6, 2, 64, value_dtype='fp32')
    qk_d = torch.zeros(1, 64, 30, 54, device='cuda')
    qe_d = torch.zeros(1, 64, 30, 54, device='cuda')
    queries = [synth.query(g, 30, 54) for _ in range(4)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        qk_d.copy_(queries[0][0]); qe_d.copy_(queries[0][1])
        (p,), out = m._plan_match(qk_d, qe_d)
        run = lambda: vos.ops.match(p.qk, p.qe, p.segments, p.values, p.rows, 30, out=p.out)
        run()                                         # allocates + initialises the workspace outside the capture
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            run()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    sim_keys = (ref.work_mem.key, ref.work_mem.shrinkage)
    for qk, qe in queries[1:] + queries[:1]:
        qk_d.copy_(qk); qe_d.copy_(qe)
        graph.replay()
        torch.cuda.synchronize()
        want = ref.match_memory(qk, qe).view(128, 1620)
        decided = orc.topk_gap(oracle_sim64(*sim_keys, qk, qe), 30)[0] > GAP
        got = out.cpu().view(128, 1620)
        assert orc.rel_err(got[:, decided], want[:, decided]) < TOL_F32


def test_match_memory_batch_equals_individual_calls(vos):
    """vosmem_match_batch (independent sequences in one selection + one readout launch, BASELINE configs[4]): three
    managers of different memory sizes / object counts -- one of them with long-term memory and two object groups'
    worth of rows -- against their own match_memory, readout and usage side effects alike."""
    g = torch.Generator().manual_seed(77)
    specs = [dict(n_frames=6, n_obj=2, n_long=0), dict(n_frames=9, n_obj=1, n_long=700), dict(n_frames=4, n_obj=3, n_long=0)]
    pairs = []
    for sp in specs:
        state = g.get_state()
        a, _ = build_manager(vos, g, (12, 20), sp['n_frames'], sp['n_obj'], 64, n_long=sp['n_long'], value_dtype='fp32')
        g.set_state(state)
        b, _ = build_manager(vos, g, (12, 20), sp['n_frames'], sp['n_obj'], 64, n_long=sp['n_long'], value_dtype='fp32')
        pairs.append((a, b))
    queries = [synth.query(g, 12, 20) for _ in specs]
    want = [a.match_memory(qk.cuda(), qe.cuda()) for (a, _), (qk, qe) in zip(pairs, queries)]
    got = vos.match_memory_batch([b for _, b in pairs], [qk.cuda() for qk, _ in queries], [qe.cuda() for _, qe in queries])
    torch.cuda.synchronize()
    for (a, b), w_, g_ in zip(pairs, want, got):
        assert g_.shape == w_.shape
        assert orc.rel_err(g_.cpu(), w_.cpu()) < 1e-3          # split counts differ -> candidate order / fp32 sums differ
        assert orc.rel_err(b.work_mem.use_count.cpu(), a.work_mem.use_count.cpu()) < 1e-3
        torch.testing.assert_close(b.work_mem.life_count, a.work_mem.life_count)


def test_readout_into_decoder_input_buffer(vos):
    """the readout written straight into channels [0, CV) of the decoder's concatenated
    num_objects x (CV + CH) x h x w input (model/modules.py:232-233) == torch.cat([match_memory, hidden], 2), for a
    single-group manager and for one with two object groups (batched launch path)."""
    g = torch.Generator().manual_seed(31)
    a, _ = build_manager(vos, g, (10, 16), 5, 3, 64, value_dtype='fp32')
    z = load('lifecycle_groups.npz')
    cfg = lifecycle_config(z)
    cfg['vosmem_value_dtype'] = 'fp32'
    b = vos.MemoryManager(cfg)
    replay_lifecycle(z, b, device='cuda')            # ends with two object groups
    assert b.work_mem.num_groups >= 2
    for m in (a, b):
        n_obj = sum(m.work_mem.group_rows(gi) for gi in range(m.work_mem.num_groups)) // m.CV
        h, w = (10, 16) if m is a else (m.H, m.W)
        m.hidden = torch.randn(1, n_obj, 64, h, w, device='cuda')
        qk, qe = synth.query(g, h, w)
        qk, qe = (qk[:, :m.CK].cuda(), qe[:, :m.CK].cuda())
        plain = m.match_memory(qk, qe)               # (usage counters do not feed back into the readout)
        fused = m.readout_with_hidden(qk, qe)
        torch.cuda.synchronize()
        assert fused.shape == (1, n_obj, m.CV + 64, h, w)
        torch.testing.assert_close(fused[0, :, :m.CV], plain, rtol=0, atol=0)
        torch.testing.assert_close(fused[:, :, m.CV:], m.hidden, rtol=0, atol=0)
    with pytest.raises(RuntimeError, match='contiguous'):
        a.match_memory_into(qk, qe, torch.empty(2, 64, 10, 16, device='cuda'))


def test_sharded_engine_single_rank_vs_oracle(vos):
    """ShardedLongTermReadout with world == 1 (CUDA backend, bf16 value shadow) == unsharded oracle readout."""
    from vos_e_sam_b200.sharded import ShardedLongTermReadout
    g = torch.Generator().manual_seed(21)
    n = 3000
    k, s, _ = synth.keys(g, n)
    v = torch.randn(2, 64, n, generator=g)
    qk, qe = synth.query(g, 9, 14)
    eng = ShardedLongTermReadout(dict(top_k=30), 0, 1, torch.device('cuda'))
    eng.load_long_term(k, s, v)
    out = eng.match(qk.cuda(), qe.cuda())
    sim = orc.anisotropic_l2(k, s, qk.flatten(2), qe.flatten(2))
    want = torch.matmul(v.reshape(128, n), orc.topk_affinity(sim, 30)[0])
    assert out.shape == want.shape and orc.rel_err(out.cpu(), want) < TOL_BF16


def test_peer_exchange_equals_nccl_exchange_on_two_gpus(vos):
    """Sharded long-term readout on 2 ranks: candidate lists pushed into the owner's peer-mapped memory over NVLink (flags,
    no collective) == exchange by one NCCL all-to-all == the unsharded readout, gathered and as query slices, over
    several frames (both slots of the double buffers reused).  Needs two GPUs; scripts/peer_check.py is the program."""
    import os, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29577', os.path.join(root, 'scripts', 'peer_check.py')],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count('max |nccl - peer|') == 2 and r.stdout.count('vs unsharded') == 6


def test_sharded_memory_manager_replays_the_lifecycle_on_two_gpus(vos):
    """MemoryManager(config['vosmem_shard'] = 'n') on 2 ranks == the single-GPU manager == the reference's recorded
    readouts, over the recorded lifecycles (consolidation, eviction, several object groups; memory_manager.py:57-150,
    211-286).  Needs two GPUs; scripts/sharded_manager_check.py is the program."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29578',
                        os.path.join(root, 'scripts', 'sharded_manager_check.py')], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count('vs single-GPU manager') == 8


def test_sharded_memory_manager_on_one_rank(vos):
    """The same code path with a world of one (runs on the driver's single-GPU box): a process group of size 1, the
    lifecycle replay through ShardedMatch (candidate ranges, two index bases, usage deltas + sync) == the recorded
    reference readouts."""
    import torch.distributed as dist
    created = False
    if not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29579')
        dist.init_process_group('gloo', rank=0, world_size=1)
        created = True
    try:
        for name in ('lifecycle_evict.npz', 'lifecycle_groups.npz'):
            z = load(name)
            cfg = dict(lifecycle_config(z), vosmem_value_dtype='fp32', vosmem_shard='n')
            m = vos.MemoryManager(cfg)
            single = vos.MemoryManager(dict(lifecycle_config(z), vosmem_value_dtype='fp32'))
            got = replay_lifecycle(z, m, device='cuda')
            replay_lifecycle(z, single, device='cuda')
            m._sharded.sync_usage()
            for i, g, want in got:
                assert orc.rel_err(g.cpu(), want) < 2e-3, f'{name} event {i}'
            assert orc.rel_err(m.work_mem.use_count.cpu(), single.work_mem.use_count.cpu()) < 1e-4
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.parametrize('rows,n,hw', [(2560, 8100, 128), (200, 1000, 77), (640, 4000, 200), (64, 256, 1), (3, 500, 128)])
def test_dense_readout_consolidation_shapes_vs_fp64(vos, rows, n, hw):
    """MemoryManager._readout as consolidation calls it (memory_manager.py:53-55,280-285): value rows taken as a slice
    of a longer bank (row pitch > n), the tcgen05 split-K product for large shapes, the few-rows kernel for the
    shrinkage row.  Against an fp64 product; the bf16 hi / lo split keeps ~16 bits per factor."""
    g = torch.Generator().manual_seed(rows + n)
    bank = torch.randn(rows, n + 300, generator=g)
    aff = torch.softmax(torch.randn(n, hw, generator=g) * 3, dim=0)
    v = bank.cuda()[:, 100:100 + n]
    got = vos.ops.readout_dense(v, aff.cuda())
    torch.cuda.synchronize()
    want = bank[:, 100:100 + n].double() @ aff.double()
    assert got.shape == (rows, hw)
    assert float((got.cpu().double() - want).abs().max()) < 5e-5 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize('n,hw', [(8100, 128), (300, 45), (129, 1), (5000, 70)])
def test_dense_softmax_parallel_over_rows(vos, n, hw):
    """do_softmax(top_k=None) (memory_util.py:55-60) through the row-parallel statistics + apply kernels, with usage."""
    g = torch.Generator().manual_seed(n)
    sim = torch.randn(1, n, hw, generator=g) * 4
    aff, usage = vos.do_softmax(sim.cuda(), top_k=None, return_usage=True)
    want = torch.softmax(sim.double(), dim=1)
    torch.testing.assert_close(aff.cpu().double(), want, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(usage.cpu().double(), want.sum(dim=2), rtol=1e-5, atol=1e-6)


def test_dense_softmax_twin_takes_large_top_k(vos):
    """memory_util.do_softmax accepts any top_k (memory_util.py:46); the dense twin serves k up to 512 (the fused
    per-frame path is limited to 32, checked at MemoryManager construction)."""
    g = torch.Generator().manual_seed(23)
    sim = torch.randn(1, 700, 50, generator=g) * 3
    for k in (33, 64, 100):
        aff, usage = vos.do_softmax(sim.cuda(), top_k=k, return_usage=True)
        want, want_usage = orc.topk_affinity(sim, k, want_usage=True)
        torch.testing.assert_close(aff.cpu(), want, rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(usage.cpu(), want_usage, rtol=1e-4, atol=1e-6)
    s2 = sim.cuda().clone()
    assert vos.do_softmax(s2, top_k=64, inplace=True).data_ptr() == s2.data_ptr()
    torch.testing.assert_close(s2.cpu(), orc.topk_affinity(sim, 64), rtol=1e-4, atol=1e-6)
    with pytest.raises(ValueError, match='top_k'):
        vos.MemoryManager(dict(hidden_dim=64, top_k=513, enable_long_term=False, enable_long_term_count_usage=False))


@pytest.mark.parametrize('top_k,n_long', [(40, 0), (64, 500), (100, 500)])
def test_match_memory_wide_top_k(vos, top_k, n_long):
    """top_k above the fused path's 32 (the reference takes any k: memory_util.py:46): MemoryManager runs the dense
    sequence of memory_manager.py:57-150 on the dense kernels.  Two object groups, usage recorded on both banks."""
    g = torch.Generator().manual_seed(top_k)
    h, w, cv = 6, 9, 32
    m, ref = build_manager(vos, g, (h, w), 3, 1, cv, n_long=n_long, value_dtype='fp32', top_k=top_k)
    for _ in range(2):       # a second object appears: group 1 sees only the later frames
        k, s, e = synth.keys(g, h * w)
        v = torch.randn(1, 2, cv, h, w, generator=g)
        args = (k.view(1, -1, h, w), s.view(1, 1, h, w), v, [1, 2])
        ref.add_memory(*args, selection=e.view(1, -1, h, w))
        m.add_memory(*(a.cuda() if isinstance(a, torch.Tensor) else a for a in args), selection=e.view(1, -1, h, w).cuda())
    assert m.work_mem.num_groups == 2
    for _ in range(2):
        qk, qe = synth.query(g, h, w)
        want = ref.match_memory(qk, qe)
        got = m.match_memory(qk.cuda(), qe.cuda())
        assert got.shape == want.shape == (2, cv, h, w)
        assert orc.rel_err(got.cpu(), want) < 1e-4
    torch.testing.assert_close(m.work_mem.use_count.cpu().flatten(), ref.work_mem.use_count.flatten(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(m.work_mem.life_count.cpu().flatten(), ref.work_mem.life_count.flatten())
    if n_long:
        torch.testing.assert_close(m.long_mem.use_count.cpu().flatten(), ref.long_mem.use_count.flatten(), rtol=1e-4, atol=1e-5)
    out = torch.zeros(2, cv + 8, h, w, device='cuda')
    qk, qe = synth.query(g, h, w)
    view = m.match_memory_into(qk.cuda(), qe.cuda(), out)
    assert view.data_ptr() == out.data_ptr() and orc.rel_err(view.cpu(), ref.match_memory(qk, qe)) < 1e-4


def test_bounded_banks_never_move(vos):
    """With the XMem bounds known (max_mid_term_frames, max_long_term_elements) both banks are allocated once: the
    device pointers the kernels (and any captured graph) hold stay valid over growth, consolidation and eviction."""
    z = load('lifecycle_evict.npz')
    cfg = lifecycle_config(z)
    m = vos.MemoryManager(cfg)
    ptrs = []

    def on_match(i, got):
        w, l = m.work_mem, m.long_mem
        ptrs.append((w._k.buf.data_ptr(), w._image.data_ptr() if w._image is not None else 0, w._groups[0].shadow.data_ptr(),
                     l._k.buf.data_ptr() if l.engaged() else None,
                     l._groups[0].shadow.data_ptr() if l.engaged() else None))
    results = replay_lifecycle(z, m, device='cuda', on_match=on_match)
    assert m.long_mem.engaged() and len(results) > 20
    assert len({p[:3] for p in ptrs}) == 1, 'working-memory buffers moved'
    assert len({p[3:] for p in ptrs if p[3] is not None}) == 1, 'long-term buffers moved'
    for i, got, want in results:
        assert orc.rel_err(got.cpu(), want) < TOL_BF16, f'event {i}'


def test_errors_are_loud(vos):
    g = torch.Generator().manual_seed(3)
    mk, ms, _ = synth.keys(g, 100)
    store = vos.KeyValueMemoryStore(False)
    store.add(dev(mk), [torch.zeros(1, 8, 100, device='cuda')], dev(ms), None, None)
    q = torch.randn(64, 10, device='cuda')
    with pytest.raises(RuntimeError, match='top_k'):
        vos.ops.select_topk(q, None, [store.key_segment(0, 100)], 33)
    with pytest.raises(RuntimeError, match='CUDA tensor'):
        vos.ops.select_topk(q.cpu(), None, [store.key_segment(0, 100)], 30)
