"""Drop-in for reference ``util/models.py``: ``compute_fn``, ``compute_vn``, ``vertex_updating`` on the GPU.

``vertex_updating`` (reference :31-44) is a per-vertex Python loop there; every sweep only reads centroids frozen
at the start of the sweep and the vertex's own position, so it is a Jacobi sweep and runs as one kernel per sweep.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import functional as F_
from .._lib import lib, ptr, set_device, stream_ptr
from ..graph import MeshTopology, topology_for


class _FacesOnly:
    """Minimal mesh view for the functions that only receive a ``faces`` array."""

    def __init__(self, faces, n_verts):
        faces = np.asarray(faces, dtype=np.int64)
        self.faces = faces
        self.vs = np.empty((n_verts, 3))
        self.f2f = -np.ones((len(faces), 3), dtype=np.int64)
        self.edges = np.zeros((0, 2), dtype=np.int64)


_faces_cache: dict = {}


def _topo_from_faces(faces, n_verts, device) -> MeshTopology:
    key = (id(faces), n_verts, str(device))
    hit = _faces_cache.get(key)
    if hit is not None and hit[0] is faces:
        return hit[1]
    topo = MeshTopology(_FacesOnly(faces, n_verts), device)
    if len(_faces_cache) > 16:
        _faces_cache.clear()
    _faces_cache[key] = (faces, topo)
    return topo


def compute_fn(vs: torch.Tensor, faces: np.ndarray) -> torch.Tensor:
    """face normals from tensor vertices (reference util/models.py:5-10), differentiable."""
    if not vs.is_cuda:
        raise RuntimeError("compute_fn: expected a CUDA tensor (dual_dmp_b200 has no CPU path)")
    return F_.FaceNormals.apply(vs, _topo_from_faces(faces, vs.shape[0], vs.device))


def compute_vn(vs: torch.Tensor, fn: torch.Tensor, faces: np.ndarray) -> torch.Tensor:
    """vertex normals = normalise(sum of incident face normals) (reference util/models.py:12-29)."""
    if not vs.is_cuda:
        raise RuntimeError("compute_vn: expected a CUDA tensor (dual_dmp_b200 has no CPU path)")
    topo = _topo_from_faces(faces, vs.shape[0], vs.device)
    set_device(vs.device)
    fn = fn.detach().to(vs.device, torch.float32).contiguous()
    vn = torch.empty(topo.V, 3, dtype=torch.float32, device=vs.device)
    lib.call("ddmp_vertex_normals", ptr(fn), ptr(topo.corner_ptr), ptr(topo.corner_slot), ptr(vn), topo.V,
             stream_ptr(vs.device))
    return vn


def vertex_updating(pos: torch.Tensor, norm: torch.Tensor, mesh, loop=10) -> torch.Tensor:
    """normal-guided vertex update (reference util/models.py:31-44)."""
    if not pos.is_cuda:
        raise RuntimeError("vertex_updating: expected a CUDA tensor (dual_dmp_b200 has no CPU path)")
    topo = topology_for(mesh, pos.device)
    set_device(pos.device)
    st = stream_ptr(pos.device)
    cur = pos.detach().to(torch.float32).contiguous().clone()
    nrm = norm.detach().to(pos.device, torch.float32).contiguous()
    nxt = torch.empty_like(cur)
    fc = torch.empty(topo.F, 3, dtype=torch.float32, device=pos.device)
    for _ in range(loop):
        lib.call("ddmp_face_centroids", ptr(cur), ptr(topo.faces), ptr(fc), topo.F, st)
        lib.call("ddmp_vertex_update_sweep", ptr(cur), ptr(fc), ptr(nrm), ptr(topo.corner_ptr),
                 ptr(topo.corner_slot), ptr(nxt), topo.V, st)
        cur, nxt = nxt, cur
    return cur
