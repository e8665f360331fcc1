"""Mirror of src/models/sepformer.py: plain 2-block Sepformer (c = 0), forward(mix) -> est."""
import torch

from ..modules import (Decoder, Dual_Path_Model, Encoder, SBTransformerBlock_CSE,  # noqa: F401
                       _make_masknet, _SepformerBase)


class Sepformer(_SepformerBase):
    def __init__(self, num_spks=2) -> None:
        super().__init__()
        self.encoder = Encoder(kernel_size=16, out_channels=256)
        self.masknet = _make_masknet(Dual_Path_Model, num_spks)
        self.decoder = Decoder(in_channels=256, out_channels=1, kernel_size=16, stride=8, bias=False)
        self.num_spks = num_spks
        self._init_common()

    def forward(self, mix: torch.Tensor) -> torch.Tensor:
        """sepformer.py:42-81: mix [B,T] -> est_source [B,T,num_spks]."""
        est, _ = self._run(mix, None, self.num_spks, False)
        return est
