"""Oracle: reference util/models.py restated on CPU torch.  TEST INFRASTRUCTURE.

compute_fn (:5-10): cross / ||cross||, no epsilon.   compute_vn (:12-29): normalise(sum of incident face normals).
vertex_updating (:31-44): ``loop`` sweeps; in each sweep face centroids are frozen (:35) but vertices are updated
in place one after the other; since the update of vertex i only reads the frozen centroids and its own position,
every sweep is order-independent (Jacobi).
"""
from __future__ import annotations

import numpy as np
import torch


def compute_fn(vs: torch.Tensor, faces: np.ndarray) -> torch.Tensor:
    f = torch.from_numpy(np.asarray(faces, dtype=np.int64))
    cr = torch.linalg.cross(vs[f[:, 1]] - vs[f[:, 0]], vs[f[:, 2]] - vs[f[:, 0]], dim=1)
    return cr / torch.sqrt((cr ** 2).sum(dim=1, keepdim=True))


def compute_vn(vs: torch.Tensor, fn: torch.Tensor, faces: np.ndarray) -> torch.Tensor:
    f = torch.from_numpy(np.asarray(faces, dtype=np.int64))
    vn = torch.zeros(len(vs), 3, dtype=fn.dtype)
    for k in range(3):
        vn = vn.index_add(0, f[:, k], fn)
    return vn / torch.sqrt((vn ** 2).sum(dim=1, keepdim=True))


def vertex_updating(pos: torch.Tensor, norm: torch.Tensor, mesh, loop: int = 10) -> torch.Tensor:
    new_pos = pos.detach().clone()
    norm = norm.detach().clone()
    faces = torch.from_numpy(np.asarray(mesh.faces, dtype=np.int64))
    vf = [sorted(s) for s in mesh.vf]
    for _ in range(loop):
        fc = new_pos[faces].sum(dim=1) / 3.0
        for i in range(len(new_pos)):
            idx = vf[i]
            ci, ni = fc[idx], norm[idx]
            proj = (ni * (ci - new_pos[i].reshape(1, -1))).sum(dim=1)
            new_pos[i] += (proj.reshape(-1, 1) * ni).sum(dim=0) / len(idx)
    return new_pos
