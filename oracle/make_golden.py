"""Generate tests/golden/*.npz from the REAL reference modules.  Run in the build container only:

    python -m oracle.make_golden          (needs /root/reference; writes tests/golden/)

It imports the unmodified reference ``util/mesh.py``, ``util/loss.py`` and ``util/models.py`` from /root/reference
(``pymeshlab`` is stubbed: it is imported at util/loss.py:4 but only used by the out-of-scope
``distance_from_reference_mesh`` :279-284), runs them on small seeded inputs and stores inputs and outputs.  The
fixtures pin oracle/loss_ref.py, oracle/models_ref.py, oracle/mesh_ref.py and the product's ``Mesh`` builder;
nothing on the GPU box reads /root/reference.  ``util/networks.py`` cannot be imported (torch_geometric /
torch_scatter are not installable here) — the GCN part of the oracle is anchored differently, see oracle/__init__.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _import_reference():
    stub = types.ModuleType("pymeshlab")
    stub.MeshSet = object
    sys.modules["pymeshlab"] = stub
    sys.path.insert(0, REF)
    import util.loss as ref_loss          # noqa: E402
    import util.mesh as ref_mesh          # noqa: E402
    import util.models as ref_models      # noqa: E402
    sys.path.pop(0)
    return ref_mesh, ref_loss, ref_models


def _canon_f2f(f2f):
    big = np.iinfo(np.int64).max
    a = np.sort(np.where(f2f < 0, big, f2f), axis=1)
    return np.where(a == big, -1, a)


def _canon_pairs(p):
    return p[:, np.lexsort((p[1], p[0]))]


def main():
    sys.path.insert(0, os.path.dirname(OUT.rstrip("/")).rsplit("/tests", 1)[0])
    from dual_dmp_b200 import synth

    ref_mesh, ref_loss, ref_models = _import_reference()
    os.makedirs(OUT, exist_ok=True)
    cases = {
        "ico3": synth.icosphere(3),
        "ico6": synth.icosphere(6),
        "open4": synth.open_patch(4, 0.5),
        "tetra": (np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], dtype=np.float64),
                  np.array([[0, 1, 2], [0, 3, 1], [0, 2, 3], [1, 3, 2]], dtype=np.int64)),
        "strip2": (np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.2]], dtype=np.float64),
                   np.array([[0, 1, 2], [1, 3, 2]], dtype=np.int64)),
    }
    for name, (vs, faces) in cases.items():
        vs = vs * (1.0 if name in ("tetra", "strip2") else 5.0)
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, name + ".obj")
            synth.write_obj(path, vs, faces)
            m = ref_mesh.Mesh(path)
        g = torch.Generator().manual_seed(1234)
        V, F = len(m.vs), len(m.faces)
        # seeded network-like outputs
        pos = torch.from_numpy(m.vs).float() + 0.05 * torch.randn(V, 3, generator=g)
        nrm = torch.from_numpy(m.fn).float() + 0.2 * torch.randn(F, 3, generator=g)
        nrm = nrm / nrm.norm(dim=1, keepdim=True)
        out = dict(vs=m.vs, faces=m.faces, edges=m.edges, f2f_canon=_canon_f2f(m.f2f), f2f_raw=m.f2f,
                   f_edges_canon=_canon_pairs(m.f_edges), v_dims=m.v_dims.numpy(), fn=m.fn, fa=m.fa, fc=m.fc,
                   vn=m.vn, pos=pos.numpy(), nrm=nrm.numpy())
        # noisy targets (the reference compares against the noisy mesh's own vs / fn)
        tgt_vs = m.vs + 0.03 * np.random.RandomState(7).randn(V, 3)
        tgt_fn = m.fn
        out["tgt_vs"] = tgt_vs
        for loop in (1, 3, 5):
            p = pos.clone().requires_grad_(True)
            n = nrm.clone().requires_grad_(True)
            l1 = ref_loss.pos_rec_loss(p, tgt_vs)
            l2 = ref_loss.mesh_laplacian_loss(p, m)
            l3 = ref_loss.norm_rec_loss(n, tgt_fn)
            l4, new_fn = ref_loss.fn_bnf_loss(p, n, m, loop=loop)
            l5 = ref_loss.pos_norm_loss(p, n, m)
            total = 3.0 * l1 + 4.0 * l2 + 4.0 * l3 + 4.0 * l4 + 1.0 * l5
            total.backward()
            out[f"loss_loop{loop}"] = np.array([l1.item(), l2.item(), l3.item(), l4.item(), l5.item()],
                                               dtype=np.float64)
            out[f"loss_dtypes_loop{loop}"] = np.array([str(x.dtype) for x in (l1, l2, l3, l4, l5)])
            out[f"gpos_loop{loop}"] = p.grad.numpy()
            out[f"gnrm_loop{loop}"] = n.grad.numpy()
            out[f"bnf_fn_loop{loop}"] = new_fn.detach().numpy()
        # individual loss gradients (unit upstream gradient)
        for key, fn_ in (("pos_rec", lambda p, n: ref_loss.pos_rec_loss(p, tgt_vs)),
                         ("lap", lambda p, n: ref_loss.mesh_laplacian_loss(p, m)),
                         ("norm_rec", lambda p, n: ref_loss.norm_rec_loss(n, tgt_fn)),
                         ("pos_norm", lambda p, n: ref_loss.pos_norm_loss(p, n, m))):
            p = pos.clone().requires_grad_(True)
            n = nrm.clone().requires_grad_(True)
            val = fn_(p, n)
            val.backward()
            out[f"g_{key}_pos"] = p.grad.numpy() if p.grad is not None else np.zeros((0,))
            out[f"g_{key}_nrm"] = n.grad.numpy() if n.grad is not None else np.zeros((0,))
        out["mad"] = np.float64(ref_loss.mad(nrm.numpy(), m.fn))
        out["mad_self"] = np.float64(ref_loss.mad(m.fn, m.fn))
        # util/models.py
        out["compute_fn"] = ref_models.compute_fn(pos, m.faces).numpy()
        out["compute_vn"] = ref_models.compute_vn(pos, torch.from_numpy(m.fn).float(), m.faces).numpy()
        if F <= 200:
            out["vertex_updating"] = ref_models.vertex_updating(pos, nrm, m, loop=3).numpy()
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
        print(name, "V", V, "F", F, "E", len(m.edges), {k: out[k] for k in ("loss_loop1", "mad")})


if __name__ == "__main__":
    main()
