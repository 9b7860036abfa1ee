"""Synthetic inputs for the Dual-DMP hot path.

Class-I ("frequency n") geodesic icosphere: F = 20 n^2, V = 10 n^2 + 2, E = 30 n^2, plus the reference's
data conventions restated from its offline tools (which need pymeshlab and are out of scope):

* rescale so the mean edge length is 1        (reference preprocess/noisemaker.py:32-36, preprocess.py:68-72)
* Gaussian noise along the vertex normal, np.random.seed(314), sigma = level (0.2)
                                              (reference preprocess/noisemaker.py:38-42)
* ``*_smooth`` mesh = 30 steps of uniform (non-cotangent) Laplacian smoothing of the noisy mesh
                                              (reference preprocess/noisemaker.py:25-26,76; MeshLab filter)

Everything is vectorised numpy with a deterministic, combinatorial vertex numbering (no float de-duplication),
so the same (n, seed) gives bit-identical meshes on every box.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

_PHI = (1.0 + 5.0 ** 0.5) / 2.0

# 12 icosahedron corners and 20 outward-oriented faces.
_ICO_V = np.array(
    [[-1, _PHI, 0], [1, _PHI, 0], [-1, -_PHI, 0], [1, -_PHI, 0],
     [0, -1, _PHI], [0, 1, _PHI], [0, -1, -_PHI], [0, 1, -_PHI],
     [_PHI, 0, -1], [_PHI, 0, 1], [-_PHI, 0, -1], [-_PHI, 0, 1]], dtype=np.float64)
_ICO_F = np.array(
    [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
     [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
     [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
     [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)


def icosphere(n: int) -> tuple[np.ndarray, np.ndarray]:
    """Unit icosphere of frequency ``n``: returns (vs [V,3] float64, faces [F,3] int64).

    Vertex ids are combinatorial: 12 corners, then (n-1) points on each of the 30 icosahedron edges, then the
    interior lattice points of each of the 20 icosahedron faces.
    """
    assert n >= 1
    corners = _ICO_V / np.linalg.norm(_ICO_V, axis=1, keepdims=True)
    # edge table of the icosahedron: sorted (lo, hi) -> edge id, in first-appearance order
    he = np.stack([_ICO_F[:, [0, 1]], _ICO_F[:, [1, 2]], _ICO_F[:, [2, 0]]], axis=1).reshape(-1, 2)
    he_s = np.sort(he, axis=1)
    ekey = he_s[:, 0] * 12 + he_s[:, 1]
    uniq, first = np.unique(ekey, return_index=True)
    order = np.argsort(first, kind="stable")
    uniq = uniq[order]
    eid_of_key = {int(k): i for i, k in enumerate(uniq)}
    n_e_pts = n - 1
    n_i_pts = (n - 1) * (n - 2) // 2
    base_edge = 12
    base_int = 12 + 30 * n_e_pts
    V = base_int + 20 * n_i_pts
    assert V == 10 * n * n + 2

    # lattice of one face: (i, j, k) with i + j + k = n are the weights of corners (a, b, c)
    ii, jj = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    keep = (ii + jj) <= n
    li = ii[keep]
    lj = jj[keep]
    lk = n - li - lj
    # local lattice index -> position in the (li, lj) list
    lat_index = -np.ones((n + 1, n + 1), dtype=np.int64)
    lat_index[li, lj] = np.arange(li.size)
    # interior numbering inside a face (i, j, k all >= 1), row-major in (i, j)
    interior = (li >= 1) & (lj >= 1) & (lk >= 1)
    int_rank = np.cumsum(interior) - 1

    vs = np.zeros((V, 3), dtype=np.float64)
    vs[:12] = corners
    gid_all = np.empty((20, li.size), dtype=np.int64)
    for f in range(20):
        a, b, c = (int(x) for x in _ICO_F[f])
        gid = np.empty(li.size, dtype=np.int64)
        # corners
        gid[(lj == 0) & (lk == 0)] = a
        gid[(li == 0) & (lk == 0)] = b
        gid[(li == 0) & (lj == 0)] = c
        # edges: the point between corners p (weight wp) and q (weight wq); parameter = weight of max(p, q)
        for (p, wp, q, wq, zero) in ((a, li, b, lj, lk), (b, lj, c, lk, li), (c, lk, a, li, lj)):
            m = (zero == 0) & (wp > 0) & (wq > 0)
            lo, hi = (p, q) if p < q else (q, p)
            t = (wq if q == hi else wp)[m]                 # weight of the hi corner, 1..n-1
            e = eid_of_key[lo * 12 + hi]
            gid[m] = base_edge + e * n_e_pts + (t - 1)
        gid[interior] = base_int + f * n_i_pts + int_rank[interior]
        gid_all[f] = gid
        p3 = (li[:, None] * corners[a] + lj[:, None] * corners[b] + lk[:, None] * corners[c]) / float(n)
        # canonical position: computed from the face that owns the first appearance; edge points are computed
        # from the sorted corner pair below so every face agrees bit-for-bit.
        vs[gid[interior]] = p3[interior]
    # edge points from their sorted corner pair (bit-identical regardless of the face that touches them)
    if n_e_pts > 0:
        t = np.arange(1, n, dtype=np.float64)[:, None]
        for key, e in eid_of_key.items():
            lo, hi = key // 12, key % 12
            vs[base_edge + e * n_e_pts: base_edge + (e + 1) * n_e_pts] = ((n - t) * corners[lo] + t * corners[hi]) / n
    vs /= np.linalg.norm(vs, axis=1, keepdims=True)

    # small triangles of one face in lattice coordinates. Corner a is at (i=n), b at (j=n), c at (k=n);
    # the lattice step i->i+1 moves toward a, j->j+1 toward b. Orientation follows (a, b, c).
    ui, uj = li[(li + lj) <= n - 1], lj[(li + lj) <= n - 1]
    up = np.stack([lat_index[ui + 1, uj], lat_index[ui, uj + 1], lat_index[ui, uj]], axis=1)
    di, dj = li[(li + lj) <= n - 2], lj[(li + lj) <= n - 2]
    down = np.stack([lat_index[di + 1, dj], lat_index[di + 1, dj + 1], lat_index[di, dj + 1]], axis=1)
    local = np.concatenate([up, down], axis=0)
    faces = np.concatenate([gid_all[f][local] for f in range(20)], axis=0)
    assert faces.shape[0] == 20 * n * n
    # make sure orientation is outward
    fnrm = np.cross(vs[faces[:, 1]] - vs[faces[:, 0]], vs[faces[:, 2]] - vs[faces[:, 0]])
    cen = vs[faces].mean(axis=1)
    flip = (fnrm * cen).sum(axis=1) < 0
    faces[flip] = faces[flip][:, [0, 2, 1]]
    return vs, faces


def unique_edges(faces: np.ndarray, n_verts: int) -> np.ndarray:
    """Undirected edges [E,2] (lo, hi) in first-appearance order over faces (v0v1, v1v2, v2v0).
    Same order as reference util/mesh.py:54-72,82 (``build_gemm``)."""
    he = np.stack([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=1).reshape(-1, 2)
    he = np.sort(he, axis=1).astype(np.int64)
    key = he[:, 0] * np.int64(n_verts) + he[:, 1]
    _, first = np.unique(key, return_index=True)
    first.sort()
    return he[first]


def face_normals_areas(vs: np.ndarray, faces: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """float64 face normals / areas exactly as reference util/mesh.py:87-92."""
    fnrm = np.cross(vs[faces[:, 1]] - vs[faces[:, 0]], vs[faces[:, 2]] - vs[faces[:, 0]])
    norm = np.linalg.norm(fnrm, axis=1, keepdims=True) + 1e-24
    fa = 0.5 * np.sqrt((fnrm ** 2).sum(axis=1))
    return fnrm / norm, fa


def vertex_normals(vs: np.ndarray, faces: np.ndarray, fn: np.ndarray) -> np.ndarray:
    """Normalised unweighted sum of incident face normals (reference util/mesh.py:94-107)."""
    vn = np.zeros_like(vs)
    idx = faces.reshape(-1)
    rep = np.repeat(fn, 3, axis=0)
    for c in range(3):
        vn[:, c] = np.bincount(idx, weights=rep[:, c], minlength=vs.shape[0])
    nrm = np.linalg.norm(vn, axis=1, keepdims=True)
    nrm[nrm == 0] = 1.0          # sklearn.normalize leaves all-zero rows untouched
    return vn / nrm


def open_patch(n: int, zcut: float = 0.5) -> tuple[np.ndarray, np.ndarray]:
    """Icosphere ``n`` with every face whose centroid has z > zcut removed and vertices compacted: an open
    (boundary) manifold mesh, so ``f2f`` has -1 entries."""
    vs, faces = icosphere(n)
    keep = vs[faces].mean(axis=1)[:, 2] <= zcut
    faces = faces[keep]
    used = np.unique(faces)
    remap = -np.ones(len(vs), dtype=np.int64)
    remap[used] = np.arange(used.size)
    return vs[used], remap[faces]


def laplacian_smooth(vs: np.ndarray, edges: np.ndarray, steps: int = 30) -> np.ndarray:
    """Uniform-weight Laplacian smoothing, ``steps`` Jacobi sweeps: x_i <- (x_i + sum_j x_j) / (deg_i + 1).
    Stands in for MeshLab ``laplacian_smooth(stepsmoothnum=30, cotangentweight=False)`` used by the reference's
    offline tool (preprocess/noisemaker.py:25-26)."""
    V = vs.shape[0]
    deg = np.bincount(edges.reshape(-1), minlength=V).astype(np.float64)
    x = vs.copy()
    src = np.concatenate([edges[:, 0], edges[:, 1]])
    dst = np.concatenate([edges[:, 1], edges[:, 0]])
    for _ in range(steps):
        acc = x.copy()
        for c in range(3):
            acc[:, c] += np.bincount(src, weights=x[dst, c], minlength=V)
        x = acc / (deg + 1.0)[:, None]
    return x


@dataclass
class SyntheticCase:
    """gt / noisy / smoothed vertex arrays that share one face list."""
    n: int
    faces: np.ndarray          # [F,3] int64
    gt_vs: np.ndarray          # [V,3] float64, mean edge length 1
    noise_vs: np.ndarray
    smooth_vs: np.ndarray


def make_case(n: int, noise_level: float = 0.2, seed: int = 314, smooth_steps: int = 30) -> SyntheticCase:
    vs, faces = icosphere(n)
    edges = unique_edges(faces, vs.shape[0])
    ave = np.linalg.norm(vs[edges[:, 0]] - vs[edges[:, 1]], axis=1).sum() / edges.shape[0]
    gt = vs / ave
    fn, _ = face_normals_areas(gt, faces)
    vn = vertex_normals(gt, faces, fn)
    rng_state = np.random.get_state()
    np.random.seed(seed)
    noise = np.random.normal(loc=0, scale=noise_level, size=(gt.shape[0], 1))
    np.random.set_state(rng_state)
    noisy = gt + vn * noise
    smooth = laplacian_smooth(noisy, edges, smooth_steps)
    return SyntheticCase(n=n, faces=faces, gt_vs=gt, noise_vs=noisy, smooth_vs=smooth)


def write_obj(path: str, vs: np.ndarray, faces: np.ndarray) -> None:
    """OBJ writer with the reference's format (util/mesh.py:267-285: float32 cast, 8 decimals, 1-based)."""
    v32 = np.asarray(vs, dtype=np.float32)
    with open(path, "w") as fp:
        fp.write("".join("v {0:.8f} {1:.8f} {2:.8f}\n".format(*row) for row in v32.tolist()))
        fp.write("".join("f {0} {1} {2}\n".format(*row) for row in (np.asarray(faces) + 1).tolist()))


def write_case(dir_path: str, case: SyntheticCase, name: str | None = None) -> str:
    """Write <name>_gt.obj / _noise.obj / _smooth.obj the way ``create_dataset`` expects them
    (reference util/datamaker.py:26-35)."""
    name = name or os.path.basename(os.path.normpath(dir_path))
    os.makedirs(dir_path, exist_ok=True)
    write_obj(os.path.join(dir_path, name + "_gt.obj"), case.gt_vs, case.faces)
    write_obj(os.path.join(dir_path, name + "_noise.obj"), case.noise_vs, case.faces)
    write_obj(os.path.join(dir_path, name + "_smooth.obj"), case.smooth_vs, case.faces)
    return dir_path
