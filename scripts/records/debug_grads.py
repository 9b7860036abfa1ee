import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
from tests.helpers import small_case, rel_err
from tests.test_gpu_nets import _pair
from dual_dmp_b200.util.datamaker import dataset_from_meshes

n_mesh, s_mesh, _ = small_case("ico", 8)
ds = dataset_from_meshes(n_mesh, s_mesh)
pr, nr, pd, nd = _pair()
for net_r, net_d, gshape in ((pr, pd, len(n_mesh.vs)), (nr, nd, len(n_mesh.faces))):
    net_r.train(); net_d.train(); net_d.reorder = False
    out_r = net_r(ds); out_d = net_d(ds)
    g = torch.randn(gshape, 3, generator=torch.Generator().manual_seed(5))
    out_r.backward(g); out_d.backward(g.to("cuda:0"))
    net_64 = copy.deepcopy(net_r).double(); net_64.zero_grad()
    ds64 = SimpleNamespace(z1=ds.z1.detach().double(), z2=ds.z2.detach().double(), x_pos=ds.x_pos.double(),
                           edge_index=ds.edge_index, face_index=ds.face_index)
    net_64(ds64).backward(g.double())
    p64 = dict(net_64.named_parameters())
    print(type(net_d).__name__)
    for (name, a), (_, b) in zip(net_d.named_parameters(), net_r.named_parameters()):
        print(f"  {name:22s} dev-vs-64 {rel_err(a.grad, p64[name].grad):.2e}  ref32-vs-64 {rel_err(b.grad, p64[name].grad):.2e}  max|g| {float(p64[name].grad.abs().max()):.2e}")
    if net_d is nd:
        a, b = nd.conv1.lin.weight.grad.cpu().double(), p64["conv1.lin.weight"].grad
        print("per-column abs err", (a - b).abs().max(dim=0).values.tolist())
        print("per-column max|g|", b.abs().max(dim=0).values.tolist())
