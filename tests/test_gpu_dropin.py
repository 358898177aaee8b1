"""GPU: the UNMODIFIED reference ``InferenceCore.step`` (tracker/inference/inference_core.py:43-150, staged under
oracle/_ref by oracle/build_ref.py) driven over a short video with a stub network, once with the reference's own
``MemoryManager`` on CUDA tensors and once after ``dropin.install()`` -- the masks must be the same (SURVEY.md
section 7 step 2 / section 8b "drop-in check")."""
import pytest
import torch

from oracle import build_ref

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not build_ref.available(), reason='oracle/_ref not staged (python oracle/build_ref.py)')]

CK, CV = 64, 64


class StubNetwork:
    """Stands in for XMem (model/network.py:40,72,107): seeded random features on the GPU, and a decoder whose logits
    are a fixed random projection of the memory readout -- so the masks depend on the readout of every object."""

    def __init__(self, seed, device='cuda'):
        self.gen, self.device = torch.Generator().manual_seed(seed), device      # host generator: same stream on any device
        self.proj = self.rnd(CV) / CV ** 0.5

    def rnd(self, *shape):
        return torch.randn(*shape, generator=self.gen).to(self.device)

    def encode_key(self, image, need_ek=True, need_sk=True):
        h, w = image.shape[-2] // 16, image.shape[-1] // 16
        key = self.rnd(1, CK, h, w)
        shrinkage = 1 + self.rnd(1, 1, h, w) ** 2 if need_sk else None
        selection = torch.sigmoid(self.rnd(1, CK, h, w)) if need_ek else None
        return key, shrinkage, selection, None, None, None

    def segment(self, feats, memory_readout, hidden, h_out=True, strip_bg=False):
        # memory_readout: 1 x num_objects x CV x h x w  ->  logits 1 x (1 + num_objects) x H x W
        obj = torch.einsum('bochw,c->bohw', memory_readout, self.proj) * 4.0
        bg = torch.zeros_like(obj[:, :1])
        logits = torch.nn.functional.interpolate(torch.cat([bg, obj], 1), scale_factor=16, mode='bilinear')
        return hidden, logits, torch.softmax(logits, dim=1)

    def encode_value(self, image, f16, hidden, masks, is_deep_update=True):
        n = masks.shape[1]
        h, w = image.shape[-2] // 16, image.shape[-1] // 16
        return self.rnd(1, n, CV, h, w), hidden


def run_video(core_cls, cfg, frames, h, w, new_object_at, device='cuda'):
    g = torch.Generator().manual_seed(77)
    core = core_cls(StubNetwork(900, device), cfg)
    readouts, snaps = [], []
    inner = core.memory.match_memory

    def logged(k, e):                                 # log every readout and what it was computed from
        mem = core.memory
        banks = ([mem.long_mem] if mem.long_mem.engaged() else []) + [mem.work_mem]
        snaps.append((k.double().cpu(), e.double().cpu(), torch.cat([b.key.double().cpu() for b in banks], -1),
                      torch.cat([b.shrinkage.double().cpu() for b in banks], -1)))
        readouts.append(inner(k, e))
        return readouts[-1]
    core.memory.match_memory = logged
    labels = [1, 2]
    core.set_all_labels(labels)
    H, W = h * 16, w * 16
    probs = []
    for t in range(frames):
        image = torch.randn(3, H, W, generator=g).to(device)
        if t == 0:
            mask = (torch.rand(len(labels), H, W, generator=g) > 0.7).float().to(device)
            prob, _ = core.step(image, mask, labels)
        elif t == new_object_at:
            labels = labels + [3]
            core.set_all_labels(labels)
            mask = (torch.rand(len(labels), H, W, generator=g) > 0.8).float().to(device)
            prob, _ = core.step(image, mask, [3])
        else:
            prob, _ = core.step(image)
        probs.append(prob.float().cpu())
    return probs, core.memory, [r.float().cpu() for r in readouts], snaps


@pytest.mark.parametrize('new_object_at', [None, 9])
def test_reference_inference_core_with_dropin_manager(new_object_at):
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    RefManager, _, InferenceCore = build_ref.load()
    import tracker.inference.inference_core as ic
    from vos_e_sam_b200 import dropin
    import vos_e_sam_b200 as vos
    cfg = dict(mem_every=2, deep_update_every=-1, enable_long_term=True, enable_long_term_count_usage=True,
               hidden_dim=8, top_k=30, max_mid_term_frames=6, min_mid_term_frames=3, num_prototypes=64,
               max_long_term_elements=4000, vosmem_value_dtype='fp32')
    frames, h, w = 36, 12, 20
    assert ic.MemoryManager is RefManager
    want, ref_mem, want_rd, snaps = run_video(InferenceCore, cfg, frames, h, w, new_object_at)
    assert type(ref_mem) is RefManager
    dropin.install()
    try:
        got, our_mem, got_rd, _ = run_video(InferenceCore, cfg, frames, h, w, new_object_at)
        assert isinstance(our_mem, vos.MemoryManager)
    finally:
        dropin.uninstall()
    assert our_mem.work_mem.size == ref_mem.work_mem.size and our_mem.long_mem.size == ref_mem.long_mem.size
    assert our_mem.long_mem.size > 0, 'the video must be long enough to consolidate into long-term memory'
    from oracle import readout_oracle as orc
    assert len(got_rd) == len(want_rd) == frames - 1
    # Parity rule (BASELINE.json north_star): a query whose k-th / (k+1)-th fp64 similarity gap is below 1e-3 may pick
    # either candidate (it moves ~1/30 of that query's weight), so readouts are compared on the decided queries.  Only
    # group 0 sees every key; a later group's candidate range is a suffix, whose own gaps are not tracked here, hence
    # the looser per-frame bound on the number of differing queries for the two-group video.
    undecided_total = 0
    for t, (a, b) in enumerate(zip(got_rd, want_rd)):
        qk, qe, mk, ms = snaps[t]
        gap = orc.topk_gap(orc.anisotropic_l2(mk, ms, qk.flatten(2), qe.flatten(2)), 30)[0]
        decided = gap > 1e-3
        undecided_total += int((~decided).sum())
        assert a.shape == b.shape
        a2, b2 = a.flatten(2), b.flatten(2)
        if new_object_at is None:
            assert orc.rel_err(a2[:, :, decided], b2[:, :, decided]) < 2e-3, f'readout {t}'
        differ = ((a2 - b2).abs().amax((0, 1)) > 2e-3)
        assert int(differ.sum()) <= int((~decided).sum()) + (3 if new_object_at is not None else 0), f'readout {t}'
    assert undecided_total < 0.2 * h * w * len(got_rd)
    for t, (a, b) in enumerate(zip(got, want)):
        assert a.shape == b.shape
        same = (a.argmax(0) == b.argmax(0)).float().mean()
        assert float(same) > 0.995, f'frame {t}: {100 * float(same):.2f}% of the mask pixels agree'
