"""Import shim (test infrastructure): rotary_embedding_torch.RotaryEmbedding as constructed at
reference unet_model.py:439. Forward of the reference never rotates anything;
only the frozen `freqs` parameter must exist so that state_dict keys match."""
import torch
from torch import nn


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, theta=10000):
        super().__init__()
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
        self.freqs = nn.Parameter(freqs, requires_grad=False)

    def rotate_queries_or_keys(self, t, seq_dim=-2):
        raise NotImplementedError('temporal attention is never executed by the reference forward')
