"""Drop-in for reference ``util/datamaker.py``: ``create_dataset(dir) -> (mesh_dic, Dataset)``.

Same files, same tensors, same field names (reference util/datamaker.py:16-21,23-107); the torch_geometric
``Data`` wrapper (:95) is replaced by a plain object because nothing downstream needs PyG.  ``dataset_from_meshes``
builds the same thing from in-memory meshes (synthetic benchmarks, no OBJ round trip).
"""
from __future__ import annotations

import glob
from typing import Tuple

import numpy as np
import torch

from .mesh import Mesh


class Dataset:
    def __init__(self, z1, z2, x_pos, x_norm, edge_index, face_index):
        self.keys = ["x", "z1", "z2", "x_pos", "x_norm", "edge_index", "face_index"]
        self.x = z1
        self.z1, self.z2 = z1, z2
        self.x_pos, self.x_norm = x_pos, x_norm
        self.edge_index, self.face_index = edge_index, face_index
        self.num_nodes = z1.shape[0]
        self.num_edges = edge_index.shape[1]
        self.num_node_features = z1.shape[1]
        self.contains_isolated_nodes = bool(
            np.setdiff1d(np.arange(self.num_nodes), edge_index.numpy().reshape(-1)).size > 0)
        self.contains_self_loops = bool((edge_index[0] == edge_index[1]).any())

    def pin_memory(self):
        """page-lock the per-step inputs so the uploads the networks do at every forward are asynchronous"""
        for k in ("z1", "z2", "x_pos", "x_norm"):
            t = getattr(self, k)
            if not t.is_pinned():
                setattr(self, k, t.detach().pin_memory().requires_grad_(t.requires_grad))
        self.x = self.z1
        return self

    def to(self, device):
        """keep the per-step inputs resident on ``device`` (the reference re-uploads them at every forward)"""
        for k in ("z1", "z2", "x_pos", "x_norm"):
            t = getattr(self, k)
            setattr(self, k, t.detach().to(device).requires_grad_(t.requires_grad))
        self.x = self.z1
        return self


def dataset_from_meshes(n_mesh: Mesh, s_mesh: Mesh) -> Dataset:
    """reference util/datamaker.py:43-96 with pos_initialization="rand16", norm_initialization="pos_norm_area"."""
    state = np.random.get_state()
    np.random.seed(314)
    z1 = np.random.normal(size=(n_mesh.vs.shape[0], 16))
    np.random.set_state(state)
    z2 = np.concatenate([n_mesh.fc, n_mesh.fn, n_mesh.fa.reshape(-1, 1)], axis=1)
    z1 = torch.tensor(z1, dtype=torch.float, requires_grad=True)
    z2 = torch.tensor(z2, dtype=torch.float, requires_grad=True)
    x_pos = torch.tensor(s_mesh.vs, dtype=torch.float)
    x_norm = torch.tensor(n_mesh.fn, dtype=torch.float)
    edge_index = torch.tensor(n_mesh.edges.T, dtype=torch.long)
    edge_index = torch.cat([edge_index, edge_index[[1, 0], :]], dim=1)
    face_index = torch.from_numpy(n_mesh.f_edges)
    return Dataset(z1, z2, x_pos, x_norm, edge_index, face_index)


def create_dataset(file_path: str) -> Tuple[dict, Dataset]:
    n_file = glob.glob(file_path + "/*_noise.obj")[0]
    s_file = glob.glob(file_path + "/*_smooth.obj")[0]
    mesh_name = n_file.split("/")[-2]
    gt_file = glob.glob(file_path + "/*_gt.obj")
    if len(gt_file) != 0:
        gt_file = gt_file[0]
        gt_mesh = Mesh(gt_file)
    else:
        gt_mesh = None
    n_mesh = Mesh(n_file)
    o1_mesh = Mesh(n_file)
    s_mesh = Mesh(s_file)
    dataset = dataset_from_meshes(n_mesh, s_mesh)
    mesh_dic = {"gt_file": gt_file, "n_file": n_file, "s_file": s_file, "mesh_name": mesh_name, "gt_mesh": gt_mesh,
                "n_mesh": n_mesh, "o1_mesh": o1_mesh, "s_mesh": s_mesh}
    return mesh_dic, dataset
