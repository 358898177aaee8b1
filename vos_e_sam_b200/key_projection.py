"""Drop-in for the reference's ``KeyProjection`` (tracker/model/modules.py:194-211; built by ``XMem`` at
tracker/model/network.py:33 and called from ``encode_key``, network.py:55): the three 3x3 convolutions that turn the
1/16-scale feature map into the query key, the shrinkage and the selection the memory readout consumes.

Same constructor, same parameter names (``key_proj`` / ``d_proj`` / ``e_proj`` are ``nn.Conv2d`` holders, so an XMem
checkpoint loads into it unchanged) and the same ``forward(x, need_s, need_e) -> (key, shrinkage, selection)``.  The
arithmetic is ONE implicit-GEMM tcgen05 kernel of libvosmem.so over all 129 output channels (csrc/keyproj.cu): the
input is read and packed once instead of three times.  Inference only (no autograd), CUDA only, no fallback.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _native as N
from . import ops
from ._native import check


class KeyProjection(nn.Module):
    def __init__(self, in_dim: int, keydim: int):
        super().__init__()
        self.key_proj = nn.Conv2d(in_dim, keydim, kernel_size=3, padding=1)
        # shrinkage
        self.d_proj = nn.Conv2d(in_dim, 1, kernel_size=3, padding=1)
        # selection
        self.e_proj = nn.Conv2d(in_dim, keydim, kernel_size=3, padding=1)
        nn.init.orthogonal_(self.key_proj.weight.data)
        nn.init.zeros_(self.key_proj.bias.data)
        self.in_dim, self.keydim = in_dim, keydim
        self._packed: Optional[torch.Tensor] = None
        self._packed_tag = None
        self._workspaces: Dict[Tuple[int, int, int, int], torch.Tensor] = {}

    def _weights(self) -> torch.Tensor:
        """Packed tensor-core image of the three weight tensors; rebuilt when a weight changed (load_state_dict, .to())."""
        ws = (self.key_proj.weight, self.d_proj.weight, self.e_proj.weight)
        tag = tuple((w.data_ptr(), w._version, w.device) for w in ws)
        if self._packed is None or tag != self._packed_tag:
            nbytes = int(N.lib.vosmem_keyproj_weight_bytes(self.in_dim, self.keydim))
            if nbytes == 0:
                raise RuntimeError(f'vos_e_sam_b200.KeyProjection: in_dim={self.in_dim}, keydim={self.keydim} not supported '
                                   f'(keydim must be 64, in_dim a multiple of 32)')
            kw, dw, ew = (ops._need(w.detach(), 'weight').contiguous() for w in ws)
            packed = torch.empty(nbytes, dtype=torch.uint8, device=kw.device)
            check(N.lib.vosmem_keyproj_pack_weights(kw.data_ptr(), dw.data_ptr(), ew.data_ptr(), self.in_dim, self.keydim,
                                                    packed.data_ptr(), ops._stream()), 'vosmem_keyproj_pack_weights')
            self._packed, self._packed_tag = packed, tag
        return self._packed

    def _workspace(self, device, h: int, w: int) -> torch.Tensor:
        index = device.index if device.index is not None else torch.cuda.current_device()
        key = (index, torch.cuda.current_stream(device).cuda_stream, h, w)
        ws = self._workspaces.get(key)
        if ws is None:
            ws = torch.empty(int(N.lib.vosmem_keyproj_workspace_bytes(self.in_dim, self.keydim, h, w)), dtype=torch.uint8,
                             device=device)
            self._workspaces[key] = ws
        return ws

    @torch.no_grad()
    def forward(self, x, need_s, need_e):
        ops._need(x, 'x')
        if x.dim() != 4 or x.shape[1] != self.in_dim:
            raise RuntimeError(f'KeyProjection: expected B x {self.in_dim} x h x w, got {tuple(x.shape)}')
        b, _, h, w = x.shape
        x = x.contiguous()
        packed = self._weights()
        ws = self._workspace(x.device, h, w)
        key = torch.empty((b, self.keydim, h, w), dtype=torch.float32, device=x.device)
        shrinkage = torch.empty((b, 1, h, w), dtype=torch.float32, device=x.device) if need_s else None
        selection = torch.empty((b, self.keydim, h, w), dtype=torch.float32, device=x.device) if need_e else None
        kb, db, eb = (ops._need(c.bias.detach(), 'bias').contiguous() for c in (self.key_proj, self.d_proj, self.e_proj))
        for i in range(b):      # (inference runs at batch 1: inference_core.py:53)
            check(N.lib.vosmem_keyproj_forward(x[i].data_ptr(), self.in_dim, self.keydim, h, w, packed.data_ptr(),
                                               kb.data_ptr(), db.data_ptr(), eb.data_ptr(), key[i].data_ptr(),
                                               ops._p(shrinkage[i]) if need_s else None,
                                               ops._p(selection[i]) if need_e else None, ws.data_ptr(), ws.numel(),
                                               ops._stream()), 'vosmem_keyproj_forward')
        return key, shrinkage, selection
