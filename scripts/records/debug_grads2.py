import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
from tests.helpers import small_case, rel_err
from tests.test_gpu_nets import _pair
from dual_dmp_b200.util.datamaker import dataset_from_meshes

for kind, n in (("ico", 8), ("open", 12)):
    n_mesh, s_mesh, _ = small_case(kind, n)
    ds = dataset_from_meshes(n_mesh, s_mesh)
    pr, nr, pd, nd = _pair()
    for net_r, net_d, gshape in ((pr, pd, len(n_mesh.vs)), (nr, nd, len(n_mesh.faces))):
        net_r.train(); net_d.train(); net_d.reorder = False
        taps_d = []; net_d.taps = taps_d
        out_d = net_d(ds)
        net_64 = copy.deepcopy(net_r).double()
        ds64 = SimpleNamespace(z1=ds.z1.detach().double(), z2=ds.z2.detach().double(), x_pos=ds.x_pos.double(),
                               edge_index=ds.edge_index, face_index=ds.face_index)
        t64 = []; net_64(ds64, t64)
        t32 = []; net_r(ds, t32)
        print(kind, n, type(net_d).__name__, "rows", gshape)
        for l, ((y64, x64), (y32, x32), (yd, st)) in enumerate(zip(t64, t32, taps_d)):
            zd = (yd * st[2] + st[3]).cpu().double()
            bn = getattr(net_64, f"bn{l+1}")
            z64 = torch.nn.functional.batch_norm(y64, None, None, bn.weight, bn.bias, True, 0.1, 1e-5)
            bn32 = getattr(net_r, f"bn{l+1}")
            z32 = torch.nn.functional.batch_norm(y32, None, None, bn32.weight, bn32.bias, True, 0.1, 1e-5).double()
            flips_d = int(((zd > 0) != (z64 > 0)).sum()); flips_32 = int(((z32 > 0) != (z64 > 0)).sum())
            print(f"   layer {l+1:2d} C={yd.shape[1]:3d} flips dev/ref32 vs 64: {flips_d}/{flips_32}  min|z64| {float(z64.abs().min()):.1e}  zerr dev {float((zd-z64).abs().max()):.1e} ref32 {float((z32-z64).abs().max()):.1e}")
