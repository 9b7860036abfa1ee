"""Drop-in for reference ``util/loss.py``: the five training losses and ``mad`` with the reference signatures,
running as fused CUDA reductions (libddmp_b200) with deterministic gather-style backward passes.

Only the ``ltype`` each driver uses (the keyword defaults, reference main.py:94-104) is implemented on the GPU;
asking for another ``ltype`` raises ``NotImplementedError`` instead of silently running something else.
``real_pos`` / ``real_norm`` may be the reference's float64 numpy arrays (uploaded at every call, like the
reference does) or CUDA tensors that are already resident.
"""
from __future__ import annotations

from typing import Union

import numpy as np
import torch

from .. import functional as F_
from ..graph import topology_for


def _target64(x, device) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    x = x.to(device=device, dtype=torch.float64, non_blocking=True)
    return x.contiguous()


def _check_cuda(t: torch.Tensor, what: str):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (dual_dmp_b200 has no CPU path)")


class _TargetTopo:
    """scratch holder for losses that have no Mesh argument (pos_rec / norm_rec)."""
    _cache: dict = {}

    def __init__(self, device):
        from .._lib import lib
        self.scratch = torch.zeros(lib.query("ddmp_loss_scratch_bytes") // 8 + 1, dtype=torch.float64, device=device)

    @classmethod
    def get(cls, device):
        key = str(device)
        if key not in cls._cache:
            cls._cache[key] = cls(device)
        return cls._cache[key]


def pos_rec_loss(pred_pos: torch.Tensor, real_pos, ltype="rmse") -> torch.Tensor:
    """reference util/loss.py:16-35."""
    if ltype != "rmse":
        raise NotImplementedError("pos_rec_loss: only ltype='rmse' (the drivers' setting) runs on the GPU path")
    _check_cuda(pred_pos, "pos_rec_loss")
    return F_.PosRecLoss.apply(pred_pos, _target64(real_pos, pred_pos.device), _TargetTopo.get(pred_pos.device))


def mesh_laplacian_loss(pred_pos: torch.Tensor, mesh, ltype="rmse") -> torch.Tensor:
    """reference util/loss.py:37-53."""
    if ltype != "rmse":
        raise NotImplementedError("mesh_laplacian_loss: only ltype='rmse' runs on the GPU path")
    _check_cuda(pred_pos, "mesh_laplacian_loss")
    return F_.LaplacianLoss.apply(pred_pos, topology_for(mesh, pred_pos.device))


def norm_rec_loss(pred_norm: torch.Tensor, real_norm, ltype="l1mae") -> torch.Tensor:
    """reference util/loss.py:55-84."""
    if ltype != "l1mae":
        raise NotImplementedError("norm_rec_loss: only ltype='l1mae' runs on the GPU path")
    _check_cuda(pred_norm, "norm_rec_loss")
    return F_.NormRecLoss.apply(pred_norm, _target64(real_norm, pred_norm.device), _TargetTopo.get(pred_norm.device))


def fn_bnf_loss(pos, fn: torch.Tensor, mesh, ltype="l1mae", loop=5):
    """reference util/loss.py:86-138; returns (loss, new_fn)."""
    if ltype != "l1mae":
        raise NotImplementedError("fn_bnf_loss: only ltype='l1mae' runs on the GPU path")
    _check_cuda(fn, "fn_bnf_loss")
    if isinstance(pos, np.ndarray):
        pos = torch.from_numpy(pos).to(fn.device, dtype=torch.float32)
    return F_.BnfLoss.apply(pos.detach(), fn, topology_for(mesh, fn.device), int(loop))


def pos_norm_loss(pos: torch.Tensor, norm: torch.Tensor, mesh, ltype="mae") -> torch.Tensor:
    """reference util/loss.py:140-160."""
    if ltype != "mae":
        raise NotImplementedError("pos_norm_loss: only ltype='mae' runs on the GPU path")
    _check_cuda(pos, "pos_norm_loss")
    if isinstance(norm, np.ndarray):
        norm = torch.from_numpy(norm).to(pos.device, dtype=torch.float32)
    return F_.PosNormLoss.apply(pos, norm, topology_for(mesh, pos.device))


def dual_loss(pos: torch.Tensor, norm: torch.Tensor, mesh, real_pos, real_norm, k=(3.0, 4.0, 4.0, 4.0, 1.0), bnfloop=1,
              bnf_scale=1.0):
    """The five loss calls and the weighted sum of the reference loop body (main.py:94-106) as ONE fused kernel:
    ``total, parts = dual_loss(pos, norm, n_mesh, n_mesh.vs, n_mesh.fn, (k1..k5), bnfloop, 0.0 if epoch <= 100 else 1.0)``
    equals ``k1*pos_rec_loss(pos, real_pos) + k2*mesh_laplacian_loss(pos, mesh) + k3*norm_rec_loss(norm, real_norm) +
    k4*(fn_bnf_loss(pos, norm, mesh, loop=bnfloop)[0] * bnf_scale) + k5*pos_norm_loss(pos, norm, mesh)`` (same
    arithmetic, same dtypes: ``total`` is float64); ``parts`` holds the five terms.  Not part of the reference API --
    the drivers' individual calls keep working -- it is what ``dual_dmp_b200.step.DualStep`` runs."""
    _check_cuda(pos, "dual_loss")
    _check_cuda(norm, "dual_loss")
    dev = pos.device
    return F_.DualLoss.apply(pos, norm, _target64(real_pos, dev), _target64(real_norm, dev), topology_for(mesh, dev),
                             tuple(float(x) for x in k), int(bnfloop), float(bnf_scale))


def mad(norm1: Union[np.ndarray, torch.Tensor], norm2: Union[np.ndarray, torch.Tensor]):
    """reference util/loss.py:261-272.  numpy inputs follow the reference's float64 numpy expression (host-side
    evaluation bookkeeping, reference main.py:60,123); when both are CUDA tensors the fused device reduction is
    used and a python float is returned."""
    if isinstance(norm1, torch.Tensor) and isinstance(norm2, torch.Tensor) and norm1.is_cuda and norm2.is_cuda:
        return float(F_.mad_device(norm1.detach(), norm2.detach(), _TargetTopo.get(norm1.device)).item())
    if isinstance(norm1, torch.Tensor):
        norm1 = norm1.to("cpu").detach().numpy().copy()
    if isinstance(norm2, torch.Tensor):
        norm2 = norm2.to("cpu").detach().numpy().copy()
    inner = np.sum(norm1 * norm2, 1)
    sad = np.rad2deg(np.arccos(np.clip(inner, -1.0, 1.0)))
    return np.sum(sad) / len(sad)


def _not_on_the_hot_path(name, where):
    def f(*a, **kw):
        raise NotImplementedError(
            f"{name} ({where}) is not called by main.py / main4real.py and is outside the accelerated hot path "
            "(SURVEY.md §2.1 row 2); use the reference implementation for it")
    f.__name__ = name
    return f


# present in reference util/loss.py but never called by its drivers: refuse loudly instead of an AttributeError
weighted_norm_rec_loss = _not_on_the_hot_path("weighted_norm_rec_loss", "reference util/loss.py:162-172")
weighted_pos_norm_loss = _not_on_the_hot_path("weighted_pos_norm_loss", "reference util/loss.py:174-192")
bnf = _not_on_the_hot_path("bnf", "reference util/loss.py:195-259, numpy bilateral filter with a per-vertex Python loop")
distance_from_reference_mesh = _not_on_the_hot_path("distance_from_reference_mesh",
                                                    "reference util/loss.py:279-284, pymeshlab Hausdorff")


def angular_difference(norm1, norm2):
    """reference util/loss.py:274-277 (used by check/mad_checker.py)."""
    inner = np.sum(norm1 * norm2, 1)
    return np.rad2deg(np.arccos(np.clip(inner, -1.0, 1.0)))
