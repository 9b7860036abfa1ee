"""Oracle: the five training losses + MAD of reference util/loss.py, restated on CPU torch.  TEST INFRASTRUCTURE.

Each function cites the reference lines it follows; only the ``ltype`` the drivers use (the keyword defaults) is
restated.  Pinned by tests/golden/loss_*.npz generated from the real reference module (oracle/make_golden.py).
``mesh`` is any object with the reference ``Mesh`` attributes (vs, faces, f2f, edges).
"""
from __future__ import annotations

import numpy as np
import torch


def _as_tensor(x, like: torch.Tensor | None = None):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x if like is None else x.to(like.device)


def pos_rec_loss(pred_pos, real_pos: np.ndarray) -> torch.Tensor:
    """reference util/loss.py:16-35, ltype="rmse":  sqrt(mean_v ||real - pred||^2 + 1e-6).
    ``real_pos`` is float64 numpy, so the result is a float64 scalar (type promotion at :27)."""
    pred_pos = _as_tensor(pred_pos)
    real = _as_tensor(real_pos, pred_pos)
    d = (real - pred_pos).abs() ** 2
    per_v = d.sum(dim=1)
    return torch.sqrt(per_v.sum() / per_v.shape[0] + 1.0e-6)


def mesh_laplacian_loss(pred_pos: torch.Tensor, mesh) -> torch.Tensor:
    """reference util/loss.py:37-53, ltype="rmse": L = (A pos)/deg; sqrt(mean_v ||pos - L||^2 + 1e-12).
    A is the unweighted vertex adjacency (reference util/mesh.py:189-197)."""
    e = torch.from_numpy(np.asarray(mesh.edges, dtype=np.int64))
    nv = pred_pos.shape[0]
    src = torch.cat([e[:, 0], e[:, 1]])
    dst = torch.cat([e[:, 1], e[:, 0]])
    agg = torch.zeros_like(pred_pos).index_add_(0, src, pred_pos.index_select(0, dst))
    deg = torch.zeros(nv, dtype=pred_pos.dtype).index_add_(0, src, torch.ones(src.shape[0], dtype=pred_pos.dtype))
    lap = agg / deg.reshape(-1, 1)
    diff = ((pred_pos - lap) ** 2).sum(dim=1)
    return torch.sqrt(diff.sum() / diff.shape[0] + 1.0e-12)


def norm_rec_loss(pred_norm, real_norm) -> torch.Tensor:
    """reference util/loss.py:55-84, ltype="l1mae": mean_f ||pred - real||_1 (float64 when real is numpy f64)."""
    pred_norm = _as_tensor(pred_norm)
    real = _as_tensor(real_norm, pred_norm)
    d = (pred_norm - real).abs().sum(dim=1)
    return d.sum() / d.shape[0]


def fn_bnf_loss(pos, fn: torch.Tensor, mesh, loop: int = 5):
    """reference util/loss.py:86-138, ltype="l1mae".  Bilateral normal filtering, ``loop`` iterations, autograd
    through all of them.  Faithful to two quirks: ``f2f == -1`` indexes the LAST face (python negative index,
    :99-100) and those wrapped slots DO count in sigma_c (:103) while their weight is masked by ``no_neig``."""
    pos = _as_tensor(pos, fn).detach()
    faces = torch.from_numpy(np.asarray(mesh.faces, dtype=np.int64))
    p0, p1, p2 = pos[faces[:, 0]], pos[faces[:, 1]], pos[faces[:, 2]]
    fc = (p0 + p1 + p2) / 3.0
    cr = torch.linalg.cross(p1 - p0, p2 - p0, dim=1)
    fa = 0.5 * torch.sqrt((cr ** 2).sum(dim=1) + 1.0e-12)
    f2f = torch.from_numpy(np.asarray(mesh.f2f, dtype=np.int64))
    has = (f2f != -1).to(fn.dtype)
    nb_fc = fc[f2f]                                   # [F,3,3], -1 wraps
    nb_fa = fa[f2f] * has
    fc_dist = ((nb_fc - fc.reshape(-1, 1, 3)) ** 2).sum(dim=2)
    sigma_c = torch.sqrt(fc_dist + 1.0e-12).sum() / (fc_dist.shape[0] * fc_dist.shape[1])
    sigma_s = 0.3
    wc = torch.exp(-1.0 * fc_dist / (2 * sigma_c ** 2))
    new_fn = fn
    for _ in range(loop):
        nb_fn = new_fn[f2f]
        fn_dist = ((nb_fn - new_fn.reshape(-1, 1, 3)) ** 2).sum(dim=2)
        ws = torch.exp(-1.0 * fn_dist / (2 * sigma_s ** 2))
        w = (wc * ws * nb_fa).unsqueeze(2)
        acc = (w * nb_fn).sum(dim=1)
        new_fn = acc / (torch.sqrt((acc * acc).sum(dim=1, keepdim=True) + 1.0e-12) + 1.0e-12)
    d = (new_fn - fn).abs().sum(dim=1)
    return d.sum() / d.shape[0], new_fn


def pos_norm_loss(pos, norm, mesh) -> torch.Tensor:
    """reference util/loss.py:140-160, ltype="mae": sum_f sum_k |(p_fk - c_f) . n_f| / V  (divides by the number
    of VERTICES, :152)."""
    pos = _as_tensor(pos)
    norm = _as_tensor(norm, pos)
    faces = torch.from_numpy(np.asarray(mesh.faces, dtype=np.int64))
    tri = pos[faces]                                  # [F,3,3]
    fc = tri.sum(dim=1) / 3.0
    pc = tri - fc.reshape(-1, 1, 3)
    dot = (pc * norm.reshape(-1, 1, 3)).sum(dim=2).abs()
    return dot.reshape(-1).sum() / len(mesh.vs)


def mad(norm1, norm2) -> float:
    """reference util/loss.py:261-272: mean angular distance in degrees, numpy float64."""
    if isinstance(norm1, torch.Tensor):
        norm1 = norm1.detach().cpu().numpy()
    if isinstance(norm2, torch.Tensor):
        norm2 = norm2.detach().cpu().numpy()
    inner = np.sum(norm1 * norm2, 1)
    sad = np.rad2deg(np.arccos(np.clip(inner, -1.0, 1.0)))
    return float(np.sum(sad) / len(sad))
