"""Oracle: PosNet / NormalNet with the reference's module tree, on CPU torch.  TEST INFRASTRUCTURE.

Follows reference util/networks.py: widths :13,:74; layer order conv -> bn -> LeakyReLU x12 (:51-62,:112-123);
PosNet head :64-67 (x_pos + linear2(lrelu(linear1(x)))); NormalNet head :125-129 (tanh, then divide by
(||t||_2 + 1e-12)).  Submodule names (conv1..12, bn1..12, linear1, linear2) equal the reference's so one
``state_dict`` loads into the oracle and into ``dual_dmp_b200.util.networks``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .gcn_ref import GCNConvRef

POS_WIDTHS = [16, 32, 64, 128, 256, 256, 512, 512, 256, 256, 128, 64, 32, 16, 3]
NORM_WIDTHS = [7, 32, 64, 128, 256, 256, 512, 512, 256, 256, 128, 64, 32, 16, 3]


class _NetRef(nn.Module):
    def __init__(self, widths, device="cpu"):
        super().__init__()
        self.device = device
        h = widths
        for i in range(12):
            setattr(self, f"conv{i + 1}", GCNConvRef(h[i], h[i + 1]))
        self.linear1 = nn.Linear(h[12], h[13])
        self.linear2 = nn.Linear(h[13], h[14])
        for i in range(12):
            setattr(self, f"bn{i + 1}", nn.BatchNorm1d(h[i + 1]))
        self.l_relu = nn.LeakyReLU()

    def trunk(self, x, edge_index, taps=None):
        for i in range(1, 13):
            y = getattr(self, f"conv{i}")(x, edge_index)
            x = self.l_relu(getattr(self, f"bn{i}")(y))
            if taps is not None:
                taps.append((y, x))
        return x


class PosNetRef(_NetRef):
    def __init__(self, device="cpu"):
        super().__init__(POS_WIDTHS, device)

    def forward(self, data, taps=None):
        z1, x_pos, edge_index = data.z1, data.x_pos, data.edge_index
        # reference :50 draws an unused randn(V,3) from the global torch RNG here; it does not affect the output
        dx = self.trunk(z1, edge_index, taps)
        dx = self.l_relu(self.linear1(dx))
        dx = self.linear2(dx)
        return x_pos + dx


class NormalNetRef(_NetRef):
    def __init__(self, device="cpu"):
        super().__init__(NORM_WIDTHS, device)

    def forward(self, data, taps=None):
        z2, edge_index = data.z2, data.face_index
        dx = self.trunk(z2, edge_index, taps)
        dx = self.l_relu(self.linear1(dx))
        dx = torch.tanh(self.linear2(dx))
        inv = torch.reciprocal(torch.norm(dx, dim=1, keepdim=True).expand(-1, 3) + 1.0e-12)
        return dx * inv
