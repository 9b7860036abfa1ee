"""GPU preprocessing (dual_dmp_b200/preprocess.py) against the numpy restatement of the reference's
offline conventions (dual_dmp_b200/synth.py: noisemaker.py:25-42, preprocess.py:68-72)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("n", [3, 16, 48])
def test_make_case_device_matches_numpy(n):
    from dual_dmp_b200 import preprocess, synth
    ref = synth.make_case(n)
    got = preprocess.make_case_device(n, DEV)
    assert np.array_equal(got.faces, ref.faces)
    for a, b, tol in ((got.gt_vs, ref.gt_vs, 1e-13), (got.noise_vs, ref.noise_vs, 1e-12), (got.smooth_vs, ref.smooth_vs, 1e-11)):
        assert a.dtype == np.float64 and a.shape == b.shape
        assert np.abs(a - b).max() <= tol * np.abs(b).max(), np.abs(a - b).max()


def test_pieces_open_mesh():
    """open (boundary) mesh: vertex normals / areas / centroids, mean edge length, bounding-box normalisation"""
    from dual_dmp_b200 import preprocess, synth
    vs, faces = synth.open_patch(10, 0.3)
    vs = vs * np.array([3.0, 1.0, 0.5]) + np.array([10.0, -2.0, 0.25])
    m = preprocess.DeviceMesh(vs, faces, DEV)
    fn, fa = synth.face_normals_areas(vs, faces)
    vn = synth.vertex_normals(vs, faces, fn)
    assert np.abs(m.face_normals().cpu().numpy() - fn).max() < 1e-13
    assert np.abs(m.vertex_normals().cpu().numpy() - vn).max() < 1e-12
    e = synth.unique_edges(faces, len(vs))
    ave = np.linalg.norm(vs[e[:, 0]] - vs[e[:, 1]], axis=1).sum() / len(e)
    assert abs(float(m.mean_edge_length()) - ave) < 1e-12 * ave
    bb = m.bbox().cpu().numpy()
    assert np.array_equal(bb[:3], vs.min(axis=0)) and np.array_equal(bb[3:], vs.max(axis=0))
    nv = preprocess.normalize(m).cpu().numpy()
    assert abs((nv.max(axis=0) - nv.min(axis=0)).max() - 1.0) < 1e-12
    assert np.abs(nv.max(axis=0) + nv.min(axis=0)).max() < 1e-12
    sv = preprocess.edge_based_scaling(m).cpu().numpy()
    assert np.abs(sv - vs / ave).max() < 1e-12 * np.abs(vs).max()
    sm = preprocess.smooth(m, steps=5).cpu().numpy()
    assert np.abs(sm - synth.laplacian_smooth(vs, e, 5)).max() < 1e-12 * np.abs(vs).max()
    assert torch.equal(preprocess.smooth(m, steps=5), torch.from_numpy(sm).to(DEV))        # deterministic
