"""Shared builders for the parity tests: seeded meshes, datasets, paired (oracle, product) networks."""
from types import SimpleNamespace

import numpy as np
import torch

from dual_dmp_b200 import synth
from dual_dmp_b200.util.mesh import Mesh


def small_case(kind="ico", n=8, seed=314):
    """(n_mesh, s_mesh, gt_mesh) with the reference's data conventions."""
    if kind == "ico":
        case = synth.make_case(n, seed=seed)
        faces, gt, noisy, smooth = case.faces, case.gt_vs, case.noise_vs, case.smooth_vs
    else:   # open (boundary) mesh
        vs, faces = synth.open_patch(n, 0.3)
        edges = synth.unique_edges(faces, len(vs))
        gt = vs / (np.linalg.norm(vs[edges[:, 0]] - vs[edges[:, 1]], axis=1).mean())
        fn, _ = synth.face_normals_areas(gt, faces)
        vn = synth.vertex_normals(gt, faces, fn)
        noisy = gt + vn * np.random.RandomState(seed).normal(0, 0.2, size=(len(gt), 1))
        smooth = synth.laplacian_smooth(noisy, edges, 30)
    return Mesh(vs=noisy, faces=faces), Mesh(vs=smooth, faces=faces), Mesh(vs=gt, faces=faces)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (both moved to CPU float64)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def to_device(ds, device):
    return SimpleNamespace(z1=ds.z1.detach().to(device), z2=ds.z2.detach().to(device), x_pos=ds.x_pos.to(device),
                           x_norm=ds.x_norm.to(device), edge_index=ds.edge_index, face_index=ds.face_index)


def report(name: str, value) -> None:
    """append a measured parity error to gpurun_out/parity.log (brought back from the GPU box)"""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    with open(os.path.join(root, "gpurun_out", "parity.log"), "a") as f:
        f.write(f"{name}\t{value}\n")
