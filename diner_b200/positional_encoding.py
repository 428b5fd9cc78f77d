"""PositionalEncoding with the reference's constructor, persistent buffers and output order
(src/models/positional_encoding.py:14-53).  Inside the render path the encoding is evaluated in
registers by the CUDA kernels (csrc/common.cuh pe_sin/pe_cos); this module exists for the
state_dict (`_freqs`, `_phases`) and for the encoder's padding code (image_encoder.py:63)."""
import numpy as np
import torch


class PositionalEncoding(torch.nn.Module):
    def __init__(self, num_freqs=6, d_in=3, freq_factor=np.pi, include_input=True):
        super().__init__()
        self.num_freqs, self.d_in, self.include_input = num_freqs, d_in, include_input
        self.freq_factor = float(freq_factor)
        self.freqs = freq_factor * 2.0 ** torch.arange(0, num_freqs)
        self.d_out = num_freqs * 2 * d_in + (d_in if include_input else 0)
        phases = torch.zeros(2 * num_freqs)
        phases[1::2] = np.pi * 0.5                      # cos(x) = sin(x + pi/2)
        self.register_buffer("_freqs", torch.repeat_interleave(self.freqs, 2).view(1, -1, 1))
        self.register_buffer("_phases", phases.view(1, -1, 1))

    def forward(self, x):
        lead = x.shape[:-1]
        flat = x.reshape(-1, x.shape[-1])
        arg = torch.addcmul(self._phases, flat.unsqueeze(1), self._freqs)      # (N, 2F, d)
        enc = torch.sin(arg).flatten(1)
        if self.include_input:
            enc = torch.cat((flat, enc), dim=-1)
        return enc.reshape(*lead, self.d_out)
