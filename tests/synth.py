"""Seeded synthetic inputs shared by the GPU parity tests, smoke() and bench.py."""
import torch

CK = 64


def keys(gen, n, ck=CK):
    """keys ~ N(0,1); shrinkage = 1 + N(0,1)^2 (modules.py:208); selection = sigmoid(N(0,1)) (modules.py:209)."""
    k = torch.randn(1, ck, n, generator=gen)
    s = 1 + torch.randn(1, 1, n, generator=gen) ** 2
    e = torch.sigmoid(torch.randn(1, ck, n, generator=gen))
    return k, s, e


def redundant_keys(gen, frames, hw, noise, ck=CK):
    """'redundant video': every memory frame is frame 0 plus small noise -> many near-ties."""
    base = torch.randn(1, ck, hw, generator=gen)
    k = torch.cat([base + noise * torch.randn(1, ck, hw, generator=gen) for _ in range(frames)], dim=-1)
    s = 1 + torch.randn(1, 1, frames * hw, generator=gen) ** 2
    return k, s


def query(gen, h, w, ck=CK):
    qk = torch.randn(1, ck, h, w, generator=gen)
    qe = torch.sigmoid(torch.randn(1, ck, h, w, generator=gen))
    return qk, qe


def keyproj_params(gen, in_dim, keydim=CK):
    """Seeded KeyProjection parameters (modules.py:198-202 shapes) at the scale of a trained projection: outputs ~ N(0,1)."""
    sc = (in_dim * 9) ** -0.5
    return dict(key_w=torch.randn(keydim, in_dim, 3, 3, generator=gen) * sc, key_b=torch.randn(keydim, generator=gen) * 0.1,
                d_w=torch.randn(1, in_dim, 3, 3, generator=gen) * sc, d_b=torch.randn(1, generator=gen) * 0.1,
                e_w=torch.randn(keydim, in_dim, 3, 3, generator=gen) * sc, e_b=torch.randn(keydim, generator=gen) * 0.1)
