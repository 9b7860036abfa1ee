"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/ddmp_b200.h
declares; the product fails loudly without a GPU instead of falling back."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from dual_dmp_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def test_header_declares_and_library_exports(built):
    decls = _lib.parse_header()
    names = [d[0] for d in decls]
    assert len(names) == len(set(names)) and len(names) >= 40
    text = open(_lib.HEADER_PATH).read()
    assert set(re.findall(r"\b(ddmp_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", text, flags=re.S))) == set(names)
    for n in names:
        assert hasattr(built, n), n
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (ddmp_\w+)", out))
    assert exported == set(names)


def test_host_only_queries(built):
    assert _lib.lib.query("ddmp_version") >= 100
    assert _lib.lib.query("ddmp_rows_per_block", 512) == 128
    assert _lib.lib.query("ddmp_rows_per_block", 32) == 256
    assert _lib.lib.query("ddmp_num_row_blocks", 1000, 512) == 8
    assert _lib.lib.query("ddmp_loss_scratch_bytes") >= 8192
    assert _lib.lib.query("ddmp_gemm_dw_workspace_bytes", 100000, 256, 512) > 0


def test_binary_is_sm100a_only(built):
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(built):
    with pytest.raises(RuntimeError):
        _lib.lib.call("ddmp_device_info", None, None, None)
    from dual_dmp_b200.util.networks import PosNet
    from dual_dmp_b200.util import loss as L
    with pytest.raises(RuntimeError, match="no CPU"):
        L.pos_rec_loss(torch.zeros(4, 3), torch.zeros(4, 3).numpy().astype("float64"))
    net = PosNet("cpu")
    from types import SimpleNamespace
    ds = SimpleNamespace(z1=torch.zeros(4, 16), x_pos=torch.zeros(4, 3), edge_index=torch.zeros(2, 0, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA only"):
        net(ds)


def test_state_dict_layout_matches_oracle():
    from dual_dmp_b200.util.networks import NormalNet, PosNet
    from oracle.networks_ref import NormalNetRef, PosNetRef
    for a, b in ((PosNet("cpu"), PosNetRef()), (NormalNet("cpu"), NormalNetRef())):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        assert all(sa[k].shape == sb[k].shape for k in sa)
        a.load_state_dict(sb)


def test_util_shim():
    """`import util.networks` etc. resolve to the drop-in when dual_dmp_b200/dropin is on the path"""
    import sys
    code = ("import util.loss as Loss, util.models as Models, util.datamaker as Datamaker\n"
            "from util.mesh import Mesh\nfrom util.networks import PosNet, NormalNet\n"
            "import dual_dmp_b200.util.networks as N\nassert PosNet is N.PosNet and NormalNet is N.NormalNet\n"
            "assert Loss.fn_bnf_loss and Models.vertex_updating and Datamaker.create_dataset and Mesh\n")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "dual_dmp_b200", "dropin"))
    subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd="/tmp")
