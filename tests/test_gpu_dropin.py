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

    def __init__(self, seed):
        self.gen = torch.Generator(device='cuda').manual_seed(seed)
        self.proj = torch.randn(CV, generator=self.gen, device='cuda') / CV ** 0.5

    def rnd(self, *shape):
        return torch.randn(*shape, generator=self.gen, device='cuda')

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


def run_video(core_cls, cfg, frames, h, w, new_object_at):
    torch.manual_seed(5)
    g = torch.Generator(device='cuda').manual_seed(77)
    core = core_cls(StubNetwork(900), cfg)
    labels = [1, 2]
    core.set_all_labels(labels)
    H, W = h * 16, w * 16
    probs = []
    for t in range(frames):
        image = torch.randn(3, H, W, generator=g, device='cuda')
        if t == 0:
            mask = (torch.rand(len(labels), H, W, generator=g, device='cuda') > 0.7).float()
            prob, _ = core.step(image, mask, labels)
        elif t == new_object_at:
            labels = labels + [3]
            core.set_all_labels(labels)
            mask = (torch.rand(len(labels), H, W, generator=g, device='cuda') > 0.8).float()
            prob, _ = core.step(image, mask, [3])
        else:
            prob, _ = core.step(image)
        probs.append(prob.float().cpu())
    return probs, core.memory


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
    want, ref_mem = run_video(InferenceCore, cfg, frames, h, w, new_object_at)
    assert type(ref_mem) is RefManager
    dropin.install()
    try:
        got, our_mem = run_video(InferenceCore, cfg, frames, h, w, new_object_at)
        assert isinstance(our_mem, vos.MemoryManager)
    finally:
        dropin.uninstall()
    assert our_mem.work_mem.size == ref_mem.work_mem.size and our_mem.long_mem.size == ref_mem.long_mem.size
    assert our_mem.long_mem.size > 0, 'the video must be long enough to consolidate into long-term memory'
    for t, (a, b) in enumerate(zip(got, want)):
        assert a.shape == b.shape
        same = (a.argmax(0) == b.argmax(0)).float().mean()
        assert float(same) > 0.999, f'frame {t}: {100 * float(same):.2f}% of the mask pixels agree'
        assert float((a - b).abs().max()) < 2e-2, f'frame {t}: probabilities differ by {float((a - b).abs().max())}'
