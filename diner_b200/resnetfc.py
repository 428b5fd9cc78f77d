"""ResnetFC parameter container with the reference's constructor and state_dict keys
(src/models/resnetfc.py:73-127: lin_in, lin_out, blocks.N.fc_0/fc_1, lin_z.N).

In the render path the network is evaluated by the fused CUDA kernels of libdiner_b200 (tcgen05 in
parity/fast mode, CUDA cores in fp32 mode), which read packed copies of these parameters;
`packed_state()` hands them over.  `forward` keeps the reference's standalone semantics
(resnetfc.py:129-159) for callers that use the module outside the renderer."""
import torch
from torch import nn


class ResnetBlockFC(nn.Module):
    def __init__(self, size_in, size_out=None, size_h=None, beta=0.0):
        super().__init__()
        size_out = size_in if size_out is None else size_out
        size_h = min(size_in, size_out) if size_h is None else size_h
        self.size_in, self.size_h, self.size_out = size_in, size_h, size_out
        self.fc_0 = nn.Linear(size_in, size_h)
        self.fc_1 = nn.Linear(size_h, size_out)
        nn.init.kaiming_normal_(self.fc_0.weight, a=0, mode="fan_in")
        nn.init.zeros_(self.fc_0.bias)
        nn.init.zeros_(self.fc_1.weight)
        nn.init.zeros_(self.fc_1.bias)
        self.activation = nn.Softplus(beta=beta) if beta > 0 else nn.ReLU()
        self.shortcut = None
        if size_in != size_out:
            self.shortcut = nn.Linear(size_in, size_out, bias=False)
            nn.init.kaiming_normal_(self.shortcut.weight, a=0, mode="fan_in")

    def forward(self, x):
        dx = self.fc_1(self.activation(self.fc_0(self.activation(x))))
        return (x if self.shortcut is None else self.shortcut(x)) + dx


class ResnetFC(nn.Module):
    def __init__(self, d_in, d_out=4, n_blocks=5, d_latent=0, d_hidden=128, beta=0.0, combine_layer=1000,
                 combine_type="average"):
        super().__init__()
        self.n_blocks, self.d_latent, self.d_in, self.d_out, self.d_hidden = n_blocks, d_latent, d_in, d_out, d_hidden
        self.combine_layer, self.combine_type, self.beta = combine_layer, combine_type, beta
        if d_in > 0:
            self.lin_in = nn.Linear(d_in, d_hidden)
        self.lin_out = nn.Linear(d_hidden, d_out)
        self.blocks = nn.ModuleList([ResnetBlockFC(d_hidden, beta=beta) for _ in range(n_blocks)])
        lins = [self.lin_out] + ([self.lin_in] if d_in > 0 else [])
        if d_latent != 0:
            self.lin_z = nn.ModuleList([nn.Linear(d_latent, d_hidden) for _ in range(min(combine_layer, n_blocks))])
            lins += list(self.lin_z)
        for lin in lins:
            nn.init.kaiming_normal_(lin.weight, a=0, mode="fan_in")
            nn.init.zeros_(lin.bias)
        self.activation = nn.Softplus(beta=beta) if beta > 0 else nn.ReLU()

    def packed_state(self):
        """(dict of contiguous fp32 parameter tensors, version stamp) for libdiner_b200."""
        sd = {k: v.detach().float().contiguous() for k, v in self.named_parameters()}
        stamp = tuple((k, v._version, v.data_ptr()) for k, v in self.named_parameters())
        return sd, stamp

    def forward(self, zx, combine_dim):
        assert zx.size(-1) == self.d_latent + self.d_in
        z, x = zx[..., :self.d_latent], zx[..., self.d_latent:]
        x = self.lin_in(x) if self.d_in > 0 else torch.zeros(self.d_hidden, device=zx.device)
        for b, blk in enumerate(self.blocks):
            if b == self.combine_layer:
                if self.combine_type != "average":          # resnetfc.py:9-14: the reference's combine() raises here as well
                    raise NotImplementedError
                x = torch.mean(x, dim=combine_dim)
            if self.d_latent > 0 and b < self.combine_layer:
                x = x + self.lin_z[b](z)
            x = blk(x)
        return self.lin_out(self.activation(x))
