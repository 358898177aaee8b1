"""KeyProjection timing on the GPU box: the tcgen05 kernel sequence (pack + implicit GEMM + finalize) against the
reference module's three cuDNN convolutions (fp32 and TF32), CUDA-graph replays, L2 flushed between replays."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import vos_e_sam_b200 as vos
from vos_e_sam_b200 import _native as N
from tests import synth

dev = torch.device('cuda')
flush = torch.empty(512 * 2 ** 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=30):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        fn()
    ts = []
    for it in range(reps):
        flush.fill_(it & 0xff)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts[5:])


class RefKeyProjection(nn.Module):     # the reference's op sequence (tracker/model/modules.py:194-211)
    def __init__(self):
        super().__init__()
        self.key_proj, self.d_proj, self.e_proj = nn.Conv2d(1024, 64, 3, padding=1), nn.Conv2d(1024, 1, 3, padding=1), nn.Conv2d(1024, 64, 3, padding=1)

    def forward(self, x):
        return self.key_proj(x), self.d_proj(x) ** 2 + 1, torch.sigmoid(self.e_proj(x))


for h, w in ((30, 54), (68, 120)):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 1024, h, w, generator=g).to(dev)
    ours = vos.KeyProjection(1024, 64).to(dev)
    ref = RefKeyProjection().to(dev)
    flops = 2.0 * h * w * 1024 * 9 * 129
    with torch.no_grad():
        t_ours = timed(lambda: ours(x, True, True))
        torch.backends.cudnn.allow_tf32 = False
        t_fp32 = timed(lambda: ref(x))
        torch.backends.cudnn.allow_tf32 = True
        t_tf32 = timed(lambda: ref(x))
    print(f'{h}x{w}: tcgen05 KeyProjection {t_ours:.1f} us ({flops / t_ours * 1e-6:.1f} TFLOP/s algorithmic, x3 executed) | '
          f'reference module cuDNN fp32 {t_fp32:.1f} us, TF32 {t_tf32:.1f} us', flush=True)
