"""Mirror of src/models/ContExt.py: contextual extraction (ContExt / H-ContExt),
forward(mix, ctx, se=None, cue='joint') -> est_source [B,T,1] (only est_mask[0] is decoded,
ContExt.py:113-119)."""
import random

import torch
import torch.nn as nn

from ..modules import (Decoder, Dual_Computation_Block_CSE, Encoder, SBTransformerBlock_CSE,  # noqa: F401
                       _make_masknet, _SepformerBase)
from ..modules import Dual_Path_Model_CSE_Ext as Dual_Path_Model_CSE


class Sepformer(_SepformerBase):
    def __init__(self, num_spks=2, add_ctx=False, add_se=False, ctx_dim=4096) -> None:
        super().__init__()
        self.encoder = Encoder(kernel_size=16, out_channels=256)
        self.masknet = _make_masknet(Dual_Path_Model_CSE, num_spks, llm_dim=ctx_dim if add_ctx else None)
        self.decoder = Decoder(in_channels=256, out_channels=1, kernel_size=16, stride=8, bias=False)
        self.se_embedding = None
        self.num_spks = num_spks
        self.add_ctx = add_ctx
        self.add_se = add_se
        self.ctx_dim = ctx_dim
        self._init_common()

    def add_ctx_pipeline(self):
        self.masknet.add_ctx()

    def add_se_pipeline(self):
        self.se_embedding = torch.nn.Linear(192, self.ctx_dim)

    def _n_masks(self):
        return 1 if self.add_ctx else self.num_spks

    def forward(self, mix: torch.Tensor, ctx: torch.Tensor, se=None, cue='joint') -> torch.Tensor:
        """ContExt.py:54-129."""
        if not self.add_ctx:
            est, _ = self._run(mix, None, self.num_spks, False)
            return est
        if self.add_se:
            se = self.se_embedding(se)                                # B x 1 x d (host-side Linear)
            if self.training:                                         # ContExt.py:98-104 (two draws)
                if random.random() < 0.3:
                    ctx = torch.cat([ctx, se], 1)
                elif 0.3 <= random.random() < 0.8:
                    ctx = torch.cat([ctx, torch.zeros_like(ctx)], 1)
                else:
                    ctx = torch.cat([torch.zeros_like(se), se], 1)
            else:
                if cue == 'joint':
                    ctx = torch.cat([ctx, se], 1)
                elif cue == 'history':
                    ctx = torch.cat([ctx, torch.zeros_like(ctx)], 1)
                elif cue == 'voice':
                    ctx = torch.cat([torch.zeros_like(se), se], 1)
        est, _ = self._run(mix, ctx, 1, False)                        # mask 0 only
        return est
