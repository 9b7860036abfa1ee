"""The GCNConv restatement (third-party torch_geometric 2.2.0; parity unpinned against the package itself) against
hand-computed tiny graphs and an independent dense float64 closed form (SURVEY.md §8c KAT list)."""
import numpy as np
import torch

from dual_dmp_b200 import synth
from oracle.gcn_ref import GCNConvRef, dense_gcn_closed_form, gcn_norm_ref
from oracle.mesh_ref import MeshRef


def _sym(edges):
    e = torch.tensor(edges, dtype=torch.long).t()
    return torch.cat([e, e[[1, 0]]], dim=1)


def test_single_node_graph_is_identity():
    ei = torch.zeros(2, 0, dtype=torch.long)
    idx, w = gcn_norm_ref(ei, 1)
    assert idx.tolist() == [[0], [0]] and w.tolist() == [1.0]


def test_k4_all_quarter():
    ei = _sym([(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)])
    idx, w = gcn_norm_ref(ei, 4)
    assert idx.shape[1] == 16
    assert torch.allclose(w, torch.full((16,), 0.25))
    # self loops are appended AFTER the original edges
    assert idx[:, -4:].tolist() == [[0, 1, 2, 3], [0, 1, 2, 3]]


def test_path_graph_weights():
    ei = _sym([(0, 1), (1, 2)])
    idx, w = gcn_norm_ref(ei, 3)
    d = {(int(a), int(b)): float(x) for a, b, x in zip(idx[0], idx[1], w)}
    s = 1 / np.sqrt(6.0)
    assert abs(d[(0, 1)] - s) < 1e-7 and abs(d[(1, 2)] - s) < 1e-7
    assert abs(d[(0, 0)] - 0.5) < 1e-7 and abs(d[(1, 1)] - 1 / 3) < 1e-7


def test_existing_self_loop_is_replaced_and_isolated_node():
    ei = torch.tensor([[0, 1, 1], [1, 0, 1]], dtype=torch.long)       # node 2 isolated, (1,1) already a loop
    idx, w = gcn_norm_ref(ei, 3)
    assert idx.shape[1] == 2 + 3
    d = {(int(a), int(b)): float(x) for a, b, x in zip(idx[0], idx[1], w)}
    assert abs(d[(2, 2)] - 1.0) < 1e-7 and abs(d[(1, 1)] - 0.5) < 1e-7


def test_conv_matches_dense_closed_form_on_meshes():
    torch.manual_seed(0)
    for vs, faces in (synth.icosphere(4), synth.open_patch(4)):
        m = MeshRef(vs, faces)
        for ei, n in ((_sym(m.edges.tolist()), len(vs)), (torch.from_numpy(m.f_edges), len(faces))):
            conv = GCNConvRef(7, 12)
            with torch.no_grad():
                conv.bias.normal_()
            x = torch.randn(n, 7)
            y = conv(x, ei)
            ref = dense_gcn_closed_form(x, ei, conv.lin.weight.detach(), conv.bias.detach())
            assert (y.double() - ref).abs().max() / ref.abs().max() < 1e-6


def test_a_hat_symmetric_so_backward_is_forward():
    vs, faces = synth.icosphere(3)
    m = MeshRef(vs, faces)
    ei = torch.from_numpy(m.f_edges)
    idx, w = gcn_norm_ref(ei, len(faces))
    a = torch.zeros(len(faces), len(faces), dtype=torch.float64)
    a.index_put_((idx[1], idx[0]), w.double(), accumulate=True)
    assert torch.equal(a, a.t())


def test_glorot_bounds_and_state_dict_keys():
    conv = GCNConvRef(16, 32)
    assert set(conv.state_dict().keys()) == {"bias", "lin.weight"}
    assert conv.lin.weight.shape == (32, 16)
    assert conv.lin.weight.abs().max() <= np.sqrt(6.0 / 48) + 1e-7
    assert torch.count_nonzero(conv.bias) == 0
