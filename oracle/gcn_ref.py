"""Oracle: torch_geometric==2.2.0 ``GCNConv`` restated in plain CPU torch.  TEST INFRASTRUCTURE (see oracle/__init__).

Third-party algorithm (not in /root/reference): torch-geometric 2.2.0 ``nn/conv/gcn_conv.py`` (``gcn_norm``,
``GCNConv.forward/message``), ``utils/loop.py`` (``add_remaining_self_loops``), ``nn/inits.py`` (``glorot``) and
torch-scatter 2.1.0 ``scatter_add``; pinned at reference requirements.txt:15,19 and called from reference
util/networks.py:15-26 (constructors, all defaults) and :51-62,:112-123 (``convK(x, edge_index)``).

Defaults in force at those call sites: improved=False, cached=False, add_self_loops=True, normalize=True,
bias=True, flow="source_to_target", aggr="add".  PARITY UNPINNED against the real package (see oracle/__init__).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


def add_remaining_self_loops_ref(edge_index: torch.Tensor, edge_weight: torch.Tensor, num_nodes: int,
                                 fill_value: float = 1.0):
    """Drop existing self loops, append one (i, i) per node AFTER the remaining edges; an existing loop keeps
    its own weight (PyG utils/loop.py add_remaining_self_loops)."""
    row, col = edge_index[0], edge_index[1]
    keep = row != col
    loop_index = torch.arange(num_nodes, dtype=edge_index.dtype)
    loop_weight = torch.full((num_nodes,), fill_value, dtype=edge_weight.dtype)
    had = ~keep
    if had.any():
        loop_weight[row[had]] = edge_weight[had]
    new_index = torch.cat([edge_index[:, keep], torch.stack([loop_index, loop_index])], dim=1)
    new_weight = torch.cat([edge_weight[keep], loop_weight])
    return new_index, new_weight


def gcn_norm_ref(edge_index: torch.Tensor, num_nodes: int, dtype=torch.float32):
    """w_e = deg(src)^-1/2 * deg(dst)^-1/2 with deg = in-degree (over the TARGET index) incl. the self loop;
    deg^-1/2 = 0 where deg = 0 (PyG gcn_norm)."""
    w = torch.ones(edge_index.shape[1], dtype=dtype)
    edge_index, w = add_remaining_self_loops_ref(edge_index, w, num_nodes, 1.0)
    row, col = edge_index[0], edge_index[1]
    deg = torch.zeros(num_nodes, dtype=dtype).scatter_add_(0, col, w)
    dis = deg.pow(-0.5)
    dis = dis.masked_fill(dis == float("inf"), 0.0)
    return edge_index, dis[row] * w * dis[col]


class _Lin(nn.Module):
    """``GCNConv.lin``: Linear(in, out, bias=False, weight_initializer='glorot'); weight is [out, in]."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin))
        a = math.sqrt(6.0 / (cin + cout))
        with torch.no_grad():
            self.weight.uniform_(-a, a)

    def forward(self, x):
        return x @ self.weight.t()


class GCNConvRef(nn.Module):
    """state_dict keys ``lin.weight`` [Cout,Cin] and ``bias`` [Cout], like PyG 2.2.0."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _Lin(in_channels, out_channels)
        self.bias = nn.Parameter(torch.zeros(out_channels))

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        n = x.shape[0]
        # cached=False: the normalisation is recomputed at every call, as the reference does 24x per step
        edge_index, w = gcn_norm_ref(edge_index, n, x.dtype)
        x = self.lin(x)                                   # transform BEFORE aggregation
        x_j = x.index_select(0, edge_index[0])            # gather source rows      [E', C]
        msg = w.view(-1, 1) * x_j                         # per-edge message        [E', C]
        out = torch.zeros(n, x.shape[1], dtype=x.dtype).index_add_(0, edge_index[1], msg)   # scatter_add
        return out + self.bias


def dense_gcn_closed_form(x: torch.Tensor, edge_index: torch.Tensor, weight: torch.Tensor,
                          bias: torch.Tensor) -> torch.Tensor:
    """Independent float64 evaluation of  D^-1/2 (A+I) D^-1/2 X W^T + b  on a dense matrix (small graphs)."""
    n = x.shape[0]
    a = torch.zeros(n, n, dtype=torch.float64)
    src, dst = edge_index[0], edge_index[1]
    keep = src != dst
    a.index_put_((dst[keep], src[keep]), torch.ones(int(keep.sum()), dtype=torch.float64), accumulate=True)
    a = a + torch.eye(n, dtype=torch.float64)
    deg = a.sum(dim=1)
    dis = deg.pow(-0.5)
    a_hat = dis[:, None] * a * dis[None, :]
    return a_hat @ (x.double() @ weight.double().t()) + bias.double()
