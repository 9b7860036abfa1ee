"""Drop-in for reference ``util/mesh.py`` (class ``Mesh``) — the graph / static-geometry builder of the hot path.

Same constructor (``Mesh(path, build_mat=False)``) and the same attributes the training path reads
(``vs, faces, fn, fa, fc, vn, edges, f2f, f_edges, v2v_mat, v_dims, vf``), but built with sorted
array passes instead of per-face Python dict/set/Counter loops (reference util/mesh.py:45-85,152-197 is
~0.1 ms/face; this is ~1 µs/face).

Bit-exactness contract (SURVEY.md §7 hard part 5): integer outputs are identical to the reference in canonical
form —
  * ``edges``   : identical array, identical order (first appearance scanning faces; reference :54-72,82);
  * ``f2f``     : identical as a set per row. The reference's per-row order comes from CPython ``set``
                  iteration (:176-186) and is not defined; here each row is ascending with the -1 padding last;
  * ``f_edges`` : identical as a multiset of directed pairs, grouped by source face like the reference;
  * ``v_dims``  : identical values; ``v2v_mat`` identical COO (edges then flipped edges, uncoalesced; :189-197);
  * float64 geometry (``fn, fa, fc, vn``) uses the same numpy expressions as the reference (:87-112).
Only manifold meshes are supported, like the reference (a third face on an edge raises there at :75-76; here it
raises ``ValueError``).
"""
from __future__ import annotations

import warnings

import numpy as np
import torch


def _fromtext(text, dtype):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)     # text-mode fromstring: the C tokeniser is the point
        return np.fromstring(text, dtype=dtype, sep=" ")


def _coo(inds, vals, size):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # "Sparse invariant checks are implicitly disabled"
        return torch.sparse_coo_tensor(inds, vals, size=size, check_invariants=False)


class Mesh:
    def __init__(self, path=None, build_mat: bool = False, *, vs=None, faces=None):
        self.path = path
        if path is not None:
            self.vs, self.faces = self.fill_from_file(path)
        else:  # in-memory construction (synthetic meshes; no OBJ round trip)
            self.vs = np.ascontiguousarray(vs, dtype=np.float64)
            self.faces = np.ascontiguousarray(faces, dtype=np.int64)
            assert self.faces.ndim == 2 and self.faces.shape[1] == 3
            assert np.logical_and(self.faces >= 0, self.faces < len(self.vs)).all()
        self.compute_face_normals()
        self.compute_face_center()
        self.device = "cpu"
        self.build_gemm()
        self.compute_vert_normals()
        self.build_v2v()
        self.build_vf()
        if build_mat:
            raise NotImplementedError(
                "build_uni_lap / build_mesh_lap are never called by the reference drivers (out of scope, "
                "SURVEY.md §2.1 row 4)")

    # ---- OBJ ingest (reference util/mesh.py:23-43) -----------------------------------------------------
    def fill_from_file(self, path):
        """Same grammar as the reference parser (``v x y z ...``, ``f a[/..] b[/..] c[/..]``, 1-based or negative
        indices, everything else ignored).  The payloads of all ``v`` / ``f`` lines are converted by one C-level
        ``np.fromstring`` each when the file is plain (exactly three numbers per ``v`` line, no ``/`` in ``f`` lines);
        otherwise a per-line tokeniser handles extra columns and ``a/b/c`` references."""
        with open(path) as f:
            lines = f.read().split("\n")
        v_lines, f_lines, f_pos = [], [], []
        for ln in lines:
            if ln.startswith("v ") or ln.startswith("v\t"):
                v_lines.append(ln[2:])
            elif ln.startswith("f ") or ln.startswith("f\t"):
                f_lines.append(ln[2:])
                f_pos.append(len(v_lines))       # vertices defined so far (negative indices are relative to it)
            elif ln[:1] in (" ", "\t"):          # indented keyword: take the slow, fully general route
                sp = ln.split(None, 1)
                if len(sp) == 2 and sp[0] == "v":
                    v_lines.append(sp[1])
                elif len(sp) == 2 and sp[0] == "f":
                    f_lines.append(sp[1])
                    f_pos.append(len(v_lines))
        vs = np.zeros((0, 3), dtype=np.float64)
        if v_lines:
            flat = _fromtext(" ".join(v_lines), np.float64)
            if flat.size == 3 * len(v_lines):
                vs = flat.reshape(-1, 3)
            else:                                # extra columns (w, colours): keep the first three like the reference
                vs = np.array([ln.split()[:3] for ln in v_lines]).astype(np.float64).reshape(-1, 3)
        faces = np.zeros((0, 3), dtype=int)
        if f_lines:
            text = " ".join(f_lines)
            if "/" not in text:
                flat = _fromtext(text, np.int64)
                assert flat.size == 3 * len(f_lines)
                flat = flat.reshape(-1, 3)
            else:
                toks = [ln.split() for ln in f_lines]
                assert all(len(t) == 3 for t in toks)
                flat = np.array([c.split("/", 1)[0] for t in toks for c in t]).astype(np.int64).reshape(-1, 3)
            rel = np.asarray(f_pos, dtype=np.int64)[:, None]
            faces = np.where(flat >= 0, flat - 1, rel + flat).astype(int)
        assert np.logical_and(faces >= 0, faces < len(vs)).all()
        return vs, faces

    # ---- edges (reference build_gemm :45-85, edges part only) --------------------------------------------
    def build_gemm(self):
        faces = self.faces
        nv = len(self.vs)
        he = np.stack([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=1).reshape(-1, 2)
        he = np.sort(he, axis=1).astype(np.int64)
        key = he[:, 0] * np.int64(nv) + he[:, 1]
        uniq, first, inv, cnt = np.unique(key, return_index=True, return_inverse=True, return_counts=True)
        if cnt.size and cnt.max() > 2:
            raise ValueError("non-manifold mesh: an edge is shared by more than two faces")
        order = np.argsort(first, kind="stable")           # first-appearance order
        rank = np.empty_like(order)
        rank[order] = np.arange(order.size)
        self.edges = he[first[order]].astype(np.int32)
        self.edges_count = int(order.size)
        # half-edge slot (3*f + k) -> edge id in first-appearance numbering
        self.face_edges = rank[inv.reshape(-1)].reshape(-1, 3)

    # ---- float64 static geometry: same expressions as the reference (:87-112) -----------------------------
    def compute_face_normals(self):
        face_normals = np.cross(self.vs[self.faces[:, 1]] - self.vs[self.faces[:, 0]],
                                self.vs[self.faces[:, 2]] - self.vs[self.faces[:, 0]])
        norm = np.linalg.norm(face_normals, axis=1, keepdims=True) + 1e-24
        face_areas = 0.5 * np.sqrt((face_normals ** 2).sum(axis=1))
        face_normals /= norm
        self.fn, self.fa = face_normals, face_areas

    def compute_vert_normals(self):
        nv = len(self.vs)
        vn = np.zeros((nv, 3), dtype=np.float64)
        idx = self.faces.reshape(-1)
        rep = np.repeat(self.fn, 3, axis=0)
        for c in range(3):
            vn[:, c] = np.bincount(idx, weights=rep[:, c], minlength=nv)
        nrm = np.sqrt((vn * vn).sum(axis=1, keepdims=True))
        nrm[nrm == 0.0] = 1.0
        self.vn = vn / nrm

    def compute_face_center(self):
        self.fc = np.sum(self.vs[self.faces], 1) / 3.0

    # ---- vertex adjacency (reference build_v2v :189-197) ------------------------------------------------
    def build_v2v(self):
        v2v_inds = self.edges.T
        v2v_inds = torch.from_numpy(np.concatenate([v2v_inds, v2v_inds[[1, 0]]], axis=1)).long()
        v2v_vals = torch.ones(v2v_inds.shape[1]).float()
        nv = len(self.vs)
        self.v2v_mat = _coo(v2v_inds, v2v_vals, (nv, nv))
        self.v_dims = torch.from_numpy(
            np.bincount(self.edges.reshape(-1).astype(np.int64), minlength=nv).astype(np.float32))

    # ---- vertex->faces and face adjacency (reference build_vf :152-187) -----------------------------------
    def build_vf(self):
        faces = self.faces
        nf, nv = len(faces), len(self.vs)
        # v->f CSR: incident faces of each vertex, ascending face id
        flat_v = faces.reshape(-1)
        flat_f = np.repeat(np.arange(nf, dtype=np.int64), 3)
        order = np.lexsort((flat_f, flat_v))
        sv, sf = flat_v[order], flat_f[order]
        if sv.size:
            keep = np.ones(sv.size, dtype=bool)
            keep[1:] = (sv[1:] != sv[:-1]) | (sf[1:] != sf[:-1])   # a degenerate face lists a vertex twice
            sv, sf = sv[keep], sf[keep]
        self.vf_ptr = np.zeros(nv + 1, dtype=np.int64)
        np.cumsum(np.bincount(sv, minlength=nv), out=self.vf_ptr[1:])
        self.vf_idx = sf
        self._vf = None

        # face adjacency through shared edges: an edge with two faces makes them neighbours
        fe = self.face_edges.reshape(-1)
        half = np.arange(3 * nf, dtype=np.int64)
        o = np.argsort(fe, kind="stable")
        fe_s, half_s = fe[o], half[o]
        pair = np.nonzero(fe_s[1:] == fe_s[:-1])[0]
        fa_, fb_ = half_s[pair] // 3, half_s[pair + 1] // 3
        ok = fa_ != fb_
        fa_, fb_ = fa_[ok], fb_[ok]
        src = np.concatenate([fa_, fb_])
        dst = np.concatenate([fb_, fa_])
        # unique directed pairs (two faces sharing two edges would otherwise repeat; the reference's Counter
        # keys are unique too)
        pk = np.unique(src * np.int64(nf) + dst)
        src, dst = pk // nf, pk % nf
        deg = np.bincount(src, minlength=nf)
        if deg.size and deg.max() > 3:
            raise ValueError("non-manifold mesh: a face has more than three edge neighbours")
        ptr = np.zeros(nf + 1, dtype=np.int64)
        np.cumsum(deg, out=ptr[1:])
        f2f = -np.ones((nf, 3), dtype=np.int64)
        slot = np.arange(src.size, dtype=np.int64) - ptr[src]
        f2f[src, slot] = dst
        self.f2f = f2f
        self.f_edges = np.stack([src, dst]).astype(np.int64)

    @property
    def vf(self):
        """list of ``set`` per vertex like the reference (:153-158). Built lazily: only
        ``models.vertex_updating`` and the numpy ``bnf`` read it, never the training step."""
        if self._vf is None:
            p, idx = self.vf_ptr, self.vf_idx
            self._vf = [set(idx[p[i]:p[i + 1]].tolist()) for i in range(len(self.vs))]
        return self._vf

    @property
    def v2f_mat(self):
        rows = np.repeat(np.arange(len(self.vs), dtype=np.int64), np.diff(self.vf_ptr))
        inds = torch.from_numpy(np.stack([rows, self.vf_idx]))
        return _coo(inds, torch.ones(inds.shape[1]), (len(self.vs), len(self.faces)))

    # ---- writers (reference :267-285) -------------------------------------------------------------------
    def save(self, filename):
        assert len(self.vs) > 0
        from ..synth import write_obj
        write_obj(filename, self.vs, self.faces)

    def save_as_ply(self, filename, fn):
        """ASCII PLY with one RGBA colour per face, ``fn`` [F,3] in [0,1] -> 0..255 (same header and number formats as
        reference util/mesh.py:287-318, which check/mad_checker.py:48 calls; host-side output, one bulk write)."""
        assert len(self.vs) > 0
        v32 = np.asarray(self.vs, dtype=np.float32)
        faces = np.asarray(self.faces, dtype=np.uint32)
        col = np.clip((255 * np.asarray(fn, dtype=np.float32)).astype(np.int64), 0, 255)     # int() truncates toward 0
        with open(filename, "w") as fp:
            fp.write("ply\nformat ascii 1.0\nelement vertex {}\n".format(len(v32)))
            fp.write("property float x\nproperty float y\nproperty float z\n")
            fp.write("element face {}\n".format(len(faces)))
            fp.write("property list uchar int vertex_indices\n")
            fp.write("property uchar red\nproperty uchar green\nproperty uchar blue\nproperty uchar alpha\n")
            fp.write("end_header\n")
            fp.write("".join("{0:.6f} {1:.6f} {2:.6f}\n".format(*row) for row in v32.tolist()))
            fp.write("".join("3 {0} {1} {2} {3} {4} {5} 255\n".format(*f, *c) for f, c in zip(faces.tolist(), col.tolist())))
