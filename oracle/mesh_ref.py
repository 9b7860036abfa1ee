"""Oracle: the reference Mesh graph builder restated with plain Python loops.  TEST INFRASTRUCTURE.

Small meshes only (O(F) Python).  Follows reference util/mesh.py: edges = first appearance scanning faces with
per-face edge order (v0v1, v1v2, v2v0), rows sorted (lo, hi) (:54-72,82); vf = set of incident faces per vertex
(:153-158); f2f(i) = faces that share exactly two vertices with face i, padded with -1 (:176-186); f_edges
(:182-183,187); v_dims = vertex degree (:193-197).  The reference's f2f row order is CPython-set dependent, so
the canonical form returned here is ascending with the -1 padding last.  Pinned against the real reference
``Mesh`` by tests/golden/mesh_*.npz.
"""
from __future__ import annotations

import numpy as np


class MeshRef:
    def __init__(self, vs: np.ndarray, faces: np.ndarray):
        self.vs = np.asarray(vs, dtype=np.float64)
        self.faces = np.asarray(faces, dtype=np.int64)
        nv, nf = len(self.vs), len(self.faces)
        # edges, first-appearance order
        seen: dict[tuple[int, int], int] = {}
        edges: list[tuple[int, int]] = []
        for f in self.faces.tolist():
            for k in range(3):
                a, b = f[k], f[(k + 1) % 3]
                key = (a, b) if a < b else (b, a)
                if key not in seen:
                    seen[key] = len(edges)
                    edges.append(key)
        self.edges = np.array(edges, dtype=np.int32).reshape(-1, 2)
        deg = np.zeros(nv, dtype=np.float32)
        for a, b in edges:
            deg[a] += 1
            deg[b] += 1
        self.v_dims = deg
        # vertex -> incident faces
        vf: list[set[int]] = [set() for _ in range(nv)]
        for i, f in enumerate(self.faces.tolist()):
            for v in f:
                vf[v].add(i)
        self.vf = vf
        # face adjacency: exactly two shared vertices
        f2f = -np.ones((nf, 3), dtype=np.int64)
        src, dst = [], []
        for i, f in enumerate(self.faces.tolist()):
            count: dict[int, int] = {}
            for v in f:
                for g in vf[v]:
                    count[g] = count.get(g, 0) + 1
            nb = sorted(g for g, c in count.items() if c == 2)
            f2f[i, :len(nb)] = nb
            src += [i] * len(nb)
            dst += nb
        self.f2f = f2f
        self.f_edges = np.array([src, dst], dtype=np.int64).reshape(2, -1)
        # float64 geometry (reference :87-112)
        cr = np.cross(self.vs[self.faces[:, 1]] - self.vs[self.faces[:, 0]],
                      self.vs[self.faces[:, 2]] - self.vs[self.faces[:, 0]])
        self.fa = 0.5 * np.sqrt((cr ** 2).sum(axis=1))
        self.fn = cr / (np.linalg.norm(cr, axis=1, keepdims=True) + 1e-24)
        self.fc = self.vs[self.faces].sum(axis=1) / 3.0
        vn = np.zeros((nv, 3))
        for i, f in enumerate(self.faces.tolist()):
            for v in f:
                vn[v] += self.fn[i]
        nrm = np.linalg.norm(vn, axis=1, keepdims=True)
        nrm[nrm == 0] = 1.0
        self.vn = vn / nrm


def canonical_f2f(f2f: np.ndarray) -> np.ndarray:
    """Row-wise ascending neighbours with the -1 padding last (the comparison form for f2f)."""
    big = np.iinfo(np.int64).max
    a = np.where(f2f < 0, big, f2f)
    a = np.sort(a, axis=1)
    return np.where(a == big, -1, a)


def canonical_pairs(pairs: np.ndarray) -> np.ndarray:
    """Lexicographically sorted directed pairs [2, M] (the comparison form for f_edges / edge_index)."""
    o = np.lexsort((pairs[1], pairs[0]))
    return pairs[:, o]
