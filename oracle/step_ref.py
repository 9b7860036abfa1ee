"""Oracle: dataset assembly and one training iteration of the reference driver, on CPU torch.  TEST INFRASTRUCTURE.

``make_dataset`` follows reference util/datamaker.py:43-96 (z1 = N(0,1)^{V x 16} with np.random.seed(314); z2 =
[fc | fn | fa] of the noisy mesh; x_pos = smoothed vertices; edge_index = [edges^T | flipped]; face_index =
f_edges).  ``train_step`` follows the loop body reference main.py:88-110: zero_grad, both forwards, five losses,
``loss_norm2 *= 0`` while epoch <= 100, weighted sum, backward, clip_grad_norm_(normnet, 0.8), two Adam steps.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

from . import loss_ref as L


def make_dataset(n_mesh, s_mesh) -> SimpleNamespace:
    state = np.random.get_state()
    np.random.seed(314)
    z1 = np.random.normal(size=(n_mesh.vs.shape[0], 16))
    np.random.set_state(state)
    z2 = np.concatenate([n_mesh.fc, n_mesh.fn, n_mesh.fa.reshape(-1, 1)], axis=1)
    edge_index = torch.tensor(np.asarray(n_mesh.edges).T, dtype=torch.long)
    edge_index = torch.cat([edge_index, edge_index[[1, 0], :]], dim=1)
    return SimpleNamespace(
        z1=torch.tensor(z1, dtype=torch.float, requires_grad=True),
        z2=torch.tensor(z2, dtype=torch.float, requires_grad=True),
        x_pos=torch.tensor(s_mesh.vs, dtype=torch.float),
        x_norm=torch.tensor(n_mesh.fn, dtype=torch.float),
        edge_index=edge_index,
        face_index=torch.from_numpy(np.asarray(n_mesh.f_edges, dtype=np.int64)),
    )


def losses(posnet, normnet, dataset, n_mesh, k=(3.0, 4.0, 4.0, 4.0, 1.0), bnfloop=1, epoch=101):
    pos = posnet(dataset)
    l_pos1 = L.pos_rec_loss(pos, n_mesh.vs)
    l_pos2 = L.mesh_laplacian_loss(pos, n_mesh)
    norm = normnet(dataset)
    l_norm1 = L.norm_rec_loss(norm, n_mesh.fn)
    l_norm2, _ = L.fn_bnf_loss(pos, norm, n_mesh, loop=bnfloop)
    if epoch <= 100:
        l_norm2 = l_norm2 * 0.0
    l_pos3 = L.pos_norm_loss(pos, norm, n_mesh)
    total = k[0] * l_pos1 + k[1] * l_pos2 + k[2] * l_norm1 + k[3] * l_norm2 + k[4] * l_pos3
    return total, (l_pos1, l_pos2, l_norm1, l_norm2, l_pos3), pos, norm


def train_step(posnet, normnet, opt_pos, opt_norm, dataset, n_mesh, k=(3.0, 4.0, 4.0, 4.0, 1.0), bnfloop=1,
               epoch=101, grad_clip=0.8):
    posnet.train()
    normnet.train()
    opt_pos.zero_grad()
    opt_norm.zero_grad()
    total, parts, pos, norm = losses(posnet, normnet, dataset, n_mesh, k, bnfloop, epoch)
    total.backward()
    nn.utils.clip_grad_norm_(normnet.parameters(), grad_clip)
    opt_pos.step()
    opt_norm.step()
    return total.detach(), parts, pos.detach(), norm.detach()
