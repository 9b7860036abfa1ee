"""Mesh preprocessing on the GPU: the reference's offline conventions (preprocess/noisemaker.py,
preprocess/preprocess.py -- both need pymeshlab) as float64 device kernels, so a 16M-face case is prepared in seconds
instead of minutes of numpy (30 smoothing sweeps at 16M faces: 30 x 3 bincounts over 48M half-edges).

``make_case_device(n)`` is the device twin of ``synth.make_case(n)`` (same conventions, same noise stream drawn on the
host with ``np.random.seed``); ``noisemaker`` / ``smooth`` / ``normalize`` / ``edge_based_scaling`` mirror the
reference tool functions of those names on device arrays.  Results agree with the numpy versions to float64 rounding
(the summation order inside a vertex's neighbourhood differs); tests/test_gpu_preprocess.py.
"""
from __future__ import annotations

import numpy as np
import torch

from . import synth
from ._lib import lib, ptr, set_device, stream_ptr


class DeviceMesh:
    """vertex array (float64, device) + the index tables the preprocessing kernels need"""

    def __init__(self, vs, faces, device, edges=None):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("dual_dmp_b200.preprocess runs on CUDA only (no CPU fallback)")
        self.device = dev
        faces = np.ascontiguousarray(faces, dtype=np.int64)
        V, F = len(vs), len(faces)
        self.V, self.F = V, F
        self.vs = torch.as_tensor(np.ascontiguousarray(vs, dtype=np.float64)).to(dev)
        self.faces_host = faces
        self.faces = torch.from_numpy(faces.astype(np.int32)).to(dev)
        e = synth.unique_edges(faces, V) if edges is None else np.asarray(edges, dtype=np.int64)
        self.edges_host = e
        self.edges = torch.from_numpy(e.astype(np.int32)).to(dev)
        rows = np.concatenate([e[:, 0], e[:, 1]])
        cols = np.concatenate([e[:, 1], e[:, 0]])
        order = np.lexsort((cols, rows))
        rp = np.zeros(V + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows, minlength=V), out=rp[1:])
        self.rowptr = torch.from_numpy(rp.astype(np.int32)).to(dev)
        self.col = torch.from_numpy(cols[order].astype(np.int32)).to(dev)
        flat = faces.reshape(-1)
        corder = np.argsort(flat, kind="stable")
        cp = np.zeros(V + 1, dtype=np.int64)
        np.cumsum(np.bincount(flat, minlength=V), out=cp[1:])
        self.corner_ptr = torch.from_numpy(cp.astype(np.int32)).to(dev)
        self.corner_slot = torch.from_numpy(corder.astype(np.int32)).to(dev)
        self.scratch = torch.zeros(int(lib.query("ddmp_prep_scratch_bytes")) // 8 + 1, dtype=torch.float64, device=dev)

    # ---- kernels ------------------------------------------------------------------------------------------------
    def _st(self):
        set_device(self.device)
        return stream_ptr(self.device)

    def face_normals(self, vs=None):
        vs = self.vs if vs is None else vs
        fn = torch.empty(self.F, 3, dtype=torch.float64, device=self.device)
        lib.call("ddmp_prep_face_geometry", ptr(vs), ptr(self.faces), ptr(fn), None, None, self.F, self._st())
        return fn

    def vertex_normals(self, vs=None):
        fn = self.face_normals(vs)
        vn = torch.empty(self.V, 3, dtype=torch.float64, device=self.device)
        lib.call("ddmp_prep_vertex_normals", ptr(fn), ptr(self.corner_ptr), ptr(self.corner_slot), ptr(vn), self.V,
                 self._st())
        return vn

    def mean_edge_length(self, vs=None) -> torch.Tensor:
        vs = self.vs if vs is None else vs
        out = torch.empty((), dtype=torch.float64, device=self.device)
        lib.call("ddmp_prep_edge_length_sum", ptr(vs), ptr(self.edges), ptr(out), ptr(self.scratch), len(self.edges_host),
                 self._st())
        return out / len(self.edges_host)

    def bbox(self, vs=None) -> torch.Tensor:
        vs = self.vs if vs is None else vs
        out = torch.empty(6, dtype=torch.float64, device=self.device)
        lib.call("ddmp_prep_bbox", ptr(vs), ptr(out), ptr(self.scratch), self.V, self._st())
        return out

    def _affine(self, a, b=None, t=None, shift=None, scale=1.0):
        out = torch.empty_like(a)
        lib.call("ddmp_prep_affine_rows", ptr(a), ptr(b), ptr(t), ptr(shift), float(scale), ptr(out), self.V, self._st())
        return out


def edge_based_scaling(mesh: DeviceMesh, vs=None) -> torch.Tensor:
    """rescale so the mean edge length is 1 (reference preprocess/noisemaker.py:32-36, preprocess.py:68-72)"""
    vs = mesh.vs if vs is None else vs
    return mesh._affine(vs, scale=1.0 / float(mesh.mean_edge_length(vs)))


def normalize(mesh: DeviceMesh, vs=None) -> torch.Tensor:
    """unit bounding box (largest side 1) and bounding-box centre at the origin (MeshLab transform_scale_normalize +
    transform_translate_center_set_origin as the reference tools apply them, preprocess/noisemaker.py:28-30)"""
    vs = mesh.vs if vs is None else vs
    bb = mesh.bbox(vs)
    side = float((bb[3:] - bb[:3]).max())
    centre = (bb[3:] + bb[:3]) * 0.5
    return mesh._affine(vs, shift=(-centre).contiguous(), scale=1.0 / side if side > 0 else 1.0)


def gaussian_noise(mesh: DeviceMesh, vs=None, level=0.2, seed=314) -> torch.Tensor:
    """vs + vn * N(0, level), one draw per vertex with np.random.seed(seed) (reference noisemaker.py:38-42)"""
    vs = mesh.vs if vs is None else vs
    state = np.random.get_state()
    np.random.seed(seed)
    noise = np.random.normal(loc=0, scale=level, size=(mesh.V, 1))
    np.random.set_state(state)
    t = torch.from_numpy(noise.reshape(-1)).to(mesh.device)
    return mesh._affine(vs, b=mesh.vertex_normals(vs), t=t)


def smooth(mesh: DeviceMesh, vs=None, steps=30) -> torch.Tensor:
    """``steps`` uniform-weight Laplacian sweeps (reference noisemaker.py:25-26: MeshLab laplacian_smooth,
    stepsmoothnum=30, cotangentweight=False; same sweep as synth.laplacian_smooth)"""
    cur = (mesh.vs if vs is None else vs).clone()
    nxt = torch.empty_like(cur)
    st = mesh._st()
    for _ in range(steps):
        lib.call("ddmp_prep_smooth_sweep", ptr(cur), ptr(mesh.rowptr), ptr(mesh.col), ptr(nxt), mesh.V, st)
        cur, nxt = nxt, cur
    return cur


def make_case_device(n: int, device="cuda:0", noise_level: float = 0.2, seed: int = 314, smooth_steps: int = 30):
    """device twin of ``synth.make_case``: icosphere n -> (gt, noisy, smoothed) with the reference's conventions"""
    vs, faces = synth.icosphere(n)
    mesh = DeviceMesh(vs, faces, device)
    gt = edge_based_scaling(mesh)
    noisy = gaussian_noise(mesh, gt, noise_level, seed)
    smoothed = smooth(mesh, noisy, smooth_steps)
    return synth.SyntheticCase(n=n, faces=faces, gt_vs=gt.cpu().numpy(), noise_vs=noisy.cpu().numpy(),
                               smooth_vs=smoothed.cpu().numpy())
