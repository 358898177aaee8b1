"""GPU: the tcgen05 KeyProjection (csrc/keyproj.cu, vos_e_sam_b200.KeyProjection) against the reference module's recorded
outputs (tests/golden/keyproj_cases.npz) and, at DAVIS-480p / 1080p feature-map sizes, against an fp64 evaluation of
tracker/model/modules.py:206-211 by the oracle.  Tolerance: the kernel computes with bf16 (hi, lo) pairs, ~16 mantissa
bits per factor -- outputs of magnitude ~1 are held to 2e-4 absolute (the north_star allows 1e-2 relative)."""
import pytest
import torch

from oracle import readout_oracle as orc
from tests import synth
from tests.replay import load, t

pytestmark = pytest.mark.gpu
ATOL = 2e-4


@pytest.fixture(scope='module')
def vos():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    import vos_e_sam_b200 as v
    return v


def module_with(vos, prm, in_dim):
    m = vos.KeyProjection(in_dim, 64).cuda()
    with torch.no_grad():
        m.key_proj.weight.copy_(prm['key_w']); m.key_proj.bias.copy_(prm['key_b'])
        m.d_proj.weight.copy_(prm['d_w']); m.d_proj.bias.copy_(prm['d_b'])
        m.e_proj.weight.copy_(prm['e_w']); m.e_proj.bias.copy_(prm['e_b'])
    return m


@pytest.mark.parametrize('name', ['small', 'xmem'])
def test_key_projection_vs_reference_golden(vos, name):
    z = load('keyproj_cases.npz')
    in_dim, h, w = (int(v) for v in z[f'{name}/shape'])
    g = torch.Generator().manual_seed(500 + in_dim)
    prm = synth.keyproj_params(g, in_dim)
    x = torch.randn(1, in_dim, h, w, generator=g)
    m = module_with(vos, prm, in_dim)
    key, shrinkage, selection = m(x.cuda(), True, True)
    torch.cuda.synchronize()
    assert key.shape == (1, 64, h, w) and shrinkage.shape == (1, 1, h, w) and selection.shape == (1, 64, h, w)
    torch.testing.assert_close(key.cpu(), t(z[f'{name}/key']), rtol=0, atol=ATOL)
    torch.testing.assert_close(shrinkage.cpu(), t(z[f'{name}/shrinkage']), rtol=2e-4, atol=ATOL)
    torch.testing.assert_close(selection.cpu(), t(z[f'{name}/selection']), rtol=0, atol=ATOL)
    key_only = m(x.cuda(), False, False)
    assert key_only[1] is None and key_only[2] is None and torch.equal(key_only[0], key)


@pytest.mark.parametrize('h,w', [(30, 54), (68, 120), (1, 1), (3, 240)])
def test_key_projection_full_size_vs_fp64_oracle(vos, h, w):
    """DAVIS-480p (30 x 54) and 1080p (68 x 120) feature maps, a single position, the widest supported row."""
    g = torch.Generator().manual_seed(900 + h)
    prm = synth.keyproj_params(g, 1024)
    x = torch.randn(2 if h == 30 else 1, 1024, h, w, generator=g)
    m = module_with(vos, prm, 1024)
    key, shrinkage, selection = m(x.cuda(), True, True)
    torch.cuda.synchronize()
    want = orc.key_projection(x.double(), **{k: v.double() for k, v in prm.items()})
    for got, ref, what in zip((key, shrinkage, selection), want, ('key', 'shrinkage', 'selection')):
        err = float((got.cpu().double() - ref).abs().max() / ref.abs().max().clamp(min=1.0))
        assert err < ATOL, f'{what} {h}x{w}: max error {err}'


def test_key_projection_feeds_the_readout(vos):
    """encode_key -> match_memory as the tracker chains them (network.py:55, inference_core.py:77-81): the projected
    key / selection of a query frame and the key / shrinkage of memory frames go through MemoryManager; the readout
    must equal the oracle's readout of the fp64-projected operands."""
    g = torch.Generator().manual_seed(41)
    h, w, n_obj, cv = 12, 20, 2, 64
    prm = synth.keyproj_params(g, 1024)
    m = module_with(vos, prm, 1024)
    mgr = vos.MemoryManager(dict(hidden_dim=8, top_k=30, enable_long_term=False, enable_long_term_count_usage=False,
                                 vosmem_value_dtype='fp32'))
    feats = [torch.randn(1, 1024, h, w, generator=g) for _ in range(4)]
    values = [torch.randn(1, n_obj, cv, h, w, generator=g) for _ in range(3)]
    for f, v in zip(feats[:3], values):
        key, shrinkage, selection = m(f.cuda(), True, True)
        mgr.add_memory(key, shrinkage, v.cuda(), [1, 2], selection=selection)
    qk, _, qe = m(feats[3].cuda(), False, True)
    got = mgr.match_memory(qk, qe)
    torch.cuda.synchronize()
    p64 = {k: v.double() for k, v in prm.items()}
    mem = [orc.key_projection(f.double(), **p64) for f in feats[:3]]
    mk = torch.cat([k.flatten(2) for k, _, _ in mem], -1)
    ms = torch.cat([s.flatten(2) for _, s, _ in mem], -1)
    q64, _, e64 = orc.key_projection(feats[3].double(), **p64, need_s=False)
    aff = orc.topk_affinity(orc.anisotropic_l2(mk, ms, q64.flatten(2), e64.flatten(2)), 30)
    mv = torch.cat([v[0].flatten(2) for v in values], -1).double().reshape(n_obj * cv, -1)
    want = torch.matmul(mv, aff[0]).view(n_obj, cv, h, w)
    assert orc.rel_err(got.cpu().double(), want) < 1e-2


def test_key_projection_errors_are_loud(vos):
    with pytest.raises(RuntimeError, match='CUDA'):
        vos.KeyProjection(64, 64)(torch.zeros(1, 64, 4, 4), True, True)
    with pytest.raises(RuntimeError, match='not supported'):
        vos.KeyProjection(64, 32).cuda()(torch.zeros(1, 64, 4, 4, device='cuda'), True, True)
    with pytest.raises(RuntimeError, match='expected'):
        vos.KeyProjection(64, 64).cuda()(torch.zeros(1, 32, 4, 4, device='cuda'), True, True)
