import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import torch, copy
from helpers import rel_err, small_map
from oracle import graph as ograph, conv as oconv
from hic_gnn_b200 import layers as glayers, utils as gutils
for n, density, mul in [(114,1.0,3.0),(300,0.95,3.0),(500,0.95,3.0),(700,0.95,3.0),(700,0.95,1.0),(700,0.3,3.0),(1500,0.95,3.0)]:
    adj = small_map(n, density, seed=2)
    g = torch.Generator().manual_seed(102)
    x = 0.25*torch.randn(n,512,generator=g)
    odata = ograph.load_input(adj.numpy().copy(), x.numpy())
    gdata = gutils.load_input(adj.numpy().copy(), x.numpy())
    torch.manual_seed(1)
    oc = oconv.GATConv(512,256,heads=2)
    with torch.no_grad():
        oc.bias.uniform_(-0.1,0.1); oc.att_l.mul_(mul); oc.att_r.mul_(mul)
    od = copy.deepcopy(oc).double(); od.lin_r = od.lin_l
    gc = glayers.GATConv(512,256,heads=2).cuda(); gc.load_state_dict(oc.state_dict())
    # oracle f64 with xl as a leaf
    xl64 = od.lin_l(x.double()).detach().requires_grad_(True)
    H,C=2,256
    def oracle_from_xl(m, xl):
        from oracle.graph import set_diag
        import torch.nn.functional as F
        x_l = xl.view(-1,H,C)
        al = (x_l*m.att_l).sum(-1); ar=(x_l*m.att_r).sum(-1)
        gg = set_diag(odata.edge_index)
        mask = torch.zeros(n,n,dtype=torch.bool); mask[gg.row,gg.col]=True
        e = al.t().unsqueeze(1)+ar.t().unsqueeze(2)
        e = F.leaky_relu(e,0.2).masked_fill(~mask.unsqueeze(0), float("-inf"))
        mm = e.max(dim=2,keepdim=True).values
        p=(e-mm).exp(); a=p/(p.sum(dim=2,keepdim=True)+1e-16)
        return torch.einsum("hij,jhc->ihc",a,x_l).reshape(n,H*C)+m.bias, e
    yo, e = oracle_from_xl(od, xl64)
    torch.manual_seed(5); w = torch.randn(n,512)
    go = torch.autograd.grad((yo*w.double()).sum(), [xl64, od.att_l, od.att_r, od.bias])
    xlg = gc.lin_l(x.cuda()).detach().requires_grad_(True)
    yg = glayers._GatAttend.apply(xlg, gc.att_l, gc.att_r, gc.bias, gdata.edge_index, 2, 256, 0.2)
    gg = torch.autograd.grad((yg*w.cuda()).sum(), [xlg, gc.att_l, gc.att_r, gc.bias])
    errs = [rel_err(yg, yo)] + [rel_err(a,b) for a,b in zip(gg,go)]
    d = (gg[0].double().cpu()-go[0]).abs()
    worst_row = int(d.max(1).values.argmax())
    fin = e[torch.isfinite(e)]
    print(n, density, mul, "y,dxl,datt_l,datt_r,dbias:", " ".join(f"{v:.2e}" for v in errs), "worst row", worst_row, "min|z|", float(fin.abs().min()), "n(|z|<1e-6)", int((fin.abs()<1e-6).sum()))
