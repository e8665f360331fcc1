"""Mirror of src/models/ContSep.py: contextual separation, forward(mix, ctx) ->
(est_source [B,T,spk], context_pred [B, spk or 1])."""
import torch
import torch.nn as nn

from ..modules import (Decoder, Dual_Computation_Block_CSE, Dual_Path_Model_CSE, Encoder,  # noqa: F401
                       SBTransformerBlock_CSE, _make_masknet, _SepformerBase)


class Sepformer(_SepformerBase):
    def __init__(self, num_spks=2, add_mt=False, ctx_dim=4096, ce=True) -> None:
        super().__init__()
        self.encoder = Encoder(kernel_size=16, out_channels=256)
        self.masknet = _make_masknet(Dual_Path_Model_CSE, num_spks, llm_dim=ctx_dim if add_mt else None)
        self.decoder = Decoder(in_channels=256, out_channels=1, kernel_size=16, stride=8, bias=False)
        self.context_selector = None
        self.num_spks = num_spks
        self.add_ctx = add_mt
        self.ce = ce
        self._init_common()

    def add_mt_pipeline(self):
        """ContSep.py:46-51."""
        self.masknet.add_ctx()
        if self.num_spks == 2 and not self.ce:
            self.context_selector = nn.Linear(256, 1)
        else:
            self.context_selector = nn.Linear(256, self.num_spks)

    def _wants_pred(self):
        return self.add_ctx

    def forward(self, mix: torch.Tensor, ctx: torch.Tensor, se=None):
        """ContSep.py:53-100."""
        if not self.add_ctx:
            est, _ = self._run(mix, None, self.num_spks, False)
            return est
        est, pred_head = self._run(mix, ctx, self.num_spks, True)
        context_pred = self.context_selector(pred_head)          # Linear(256 -> spk | 1), host side
        return est, context_pred
