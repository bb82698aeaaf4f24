"""PyTorch statement of the Mega-NeRF sub-MLP shapes named by BASELINE.json /
SURVEY.md §8 A9 (the architecture itself is not part of the reference repo:
parity for this row is "unpinned"; this module is the oracle the CUDA kernel is
checked against).  `forward(x, emulate_bf16=True)` rounds GEMM operands to bf16
exactly where the tcgen05 kernel does (weights, activations entering a Linear)
and keeps fp32 accumulation / bias / activations."""
from __future__ import annotations

import numpy as np
import torch
from torch import nn


def positional_encoding(v: torch.Tensor, freqs: int) -> torch.Tensor:
    out = [v]
    for o in range(freqs):
        f = float(2 ** o)
        out += [torch.sin(f * v), torch.cos(f * v)]
    return torch.cat(out, -1)


class MegaNerfMLP(nn.Module):
    def __init__(self, basis_dim=9, n_appearance=4, appearance_dim=48, need_viewdir=False,
                 width=256, n_layers=8, skip_layer=4, pe_xyz=12, pe_dir=4, head_width=128,
                 sigma_activation="softplus"):
        super().__init__()
        self.pe_xyz, self.pe_dir, self.skip_layer, self.need_viewdir = pe_xyz, pe_dir, skip_layer, need_viewdir
        pe = 3 + 6 * pe_xyz
        self.trunk = nn.ModuleList(
            [nn.Linear(pe if i == 0 else (width + pe if i == skip_layer else width), width) for i in range(n_layers)])
        self.sigma = nn.Linear(width, 1)
        self.final = nn.Linear(width, width)
        self.embedding = nn.Embedding(n_appearance, appearance_dim) if appearance_dim > 0 else None
        head_in = width + (3 + 6 * pe_dir if need_viewdir else 0) + (appearance_dim if appearance_dim > 0 else 0)
        self.head1 = nn.Linear(head_in, head_width)
        self.head2 = nn.Linear(head_width, 3 * basis_dim)
        self.sigma_activation = sigma_activation

    @staticmethod
    def _lin(layer: nn.Linear, x: torch.Tensor, bf16: bool) -> torch.Tensor:
        if not bf16:
            return layer(x)
        w = layer.weight.to(torch.bfloat16).to(torch.float32)
        xb = x.to(torch.bfloat16).to(torch.float32)
        return xb @ w.t() + layer.bias

    def forward(self, x: torch.Tensor, emulate_bf16: bool = False) -> torch.Tensor:
        xyz = x[:, :3]
        col = 3
        h0 = positional_encoding(xyz, self.pe_xyz)
        h = h0
        for i, layer in enumerate(self.trunk):
            inp = h if i == 0 else (torch.cat([h0, h], -1) if i == self.skip_layer else h)
            h = torch.relu(self._lin(layer, inp, emulate_bf16))
        sigma = h @ self.sigma.weight.t() + self.sigma.bias  # fp32 CUDA-core dot in the kernel
        sigma = torch.nn.functional.softplus(sigma) if self.sigma_activation == "softplus" else torch.relu(sigma)
        f = self._lin(self.final, h, emulate_bf16)
        parts = [f]
        if self.need_viewdir:
            parts.append(positional_encoding(x[:, col:col + 3], self.pe_dir))
            col += 3
        if self.embedding is not None:
            idx = x[:, col].long().clamp(0, self.embedding.num_embeddings - 1)
            parts.append(self.embedding(idx))
        g = torch.relu(self._lin(self.head1, torch.cat(parts, -1), emulate_bf16))
        rgb = self._lin(self.head2, g, emulate_bf16)
        return torch.cat([rgb, sigma], -1)

    def export(self) -> dict:
        """numpy fp32 arrays in the layout mega_nerf_viewer_b200.MlpModel expects."""
        n = lambda t: t.detach().cpu().numpy().astype(np.float32)
        return dict(
            trunk_w=[n(l.weight) for l in self.trunk], trunk_b=[n(l.bias) for l in self.trunk],
            sigma_w=n(self.sigma.weight), sigma_b=n(self.sigma.bias),
            final_w=n(self.final.weight), final_b=n(self.final.bias),
            embedding=None if self.embedding is None else n(self.embedding.weight),
            head1_w=n(self.head1.weight), head1_b=n(self.head1.bias),
            head2_w=n(self.head2.weight), head2_b=n(self.head2.bias),
            skip_layer=self.skip_layer, pe_xyz_freqs=self.pe_xyz, pe_dir_freqs=self.pe_dir,
            need_viewdir=self.need_viewdir, sigma_activation=1 if self.sigma_activation == "softplus" else 0)


def flops_per_row(m: MegaNerfMLP) -> float:
    macs = sum(l.in_features * l.out_features for l in list(m.trunk) + [m.sigma, m.final, m.head1, m.head2])
    return 2.0 * macs
