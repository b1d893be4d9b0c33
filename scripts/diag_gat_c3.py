"""Diagnostic: per-tensor error of the GAT backward at the C3 size against the f64 oracle (CSR / dense paths, f32 oracle)."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_conv import _gat_case
from helpers import rel_err
from hic_gnn_b200 import layers as glayers

for n, dens, sharpen in ((1500, 0.5, True), (2493, 0.95, True), (2493, 0.95, False), (2493, 0.3, True)):
    x, odata, gdata, oc = _gat_case(n, dens, min_kink_gap=0.0)
    if not sharpen:
        with torch.no_grad():
            oc.att_l.div_(3.0); oc.att_r.div_(3.0)
    w = torch.randn(n, 512, generator=torch.Generator().manual_seed(7))
    def og(dtype):
        m = copy.deepcopy(oc).to(dtype)
        xo = x.to(dtype).clone().requires_grad_(True)
        yo = m(xo, odata.edge_index, dense=True)
        return yo.detach(), torch.autograd.grad((yo * w.to(dtype)).sum(), [xo, m.lin_l.weight, m.att_l, m.att_r, m.bias])
    y64, g64 = og(torch.float64); y32, g32 = og(torch.float32)
    print(f"n={n} dens={dens} sharpen={sharpen}  oracle32: y {rel_err(y32,y64):.2e} " + " ".join(f"{k} {rel_err(a,b):.2e}" for k,a,b in zip("x W al ar b".split(), g32, g64)))
    for path in ("csr", "dense"):
        for bw in ("fused", "split"):
            if path == "dense" and bw == "split": continue
            glayers.GAT_BACKWARD = bw
            gc = glayers.GATConv(512, 256, heads=2).cuda(); gc.path = path; gc.load_state_dict(oc.state_dict())
            xg = x.cuda().requires_grad_(True)
            yg = gc(xg, gdata.edge_index)
            gg = torch.autograd.grad((yg * w.cuda()).sum(), [xg, gc.lin_l.weight, gc.att_l, gc.att_r, gc.bias])
            print(f"   {path:5s} {bw:5s}: y {rel_err(yg,y64):.2e} " + " ".join(f"{k} {rel_err(a,b):.2e}" for k,a,b in zip("x W al ar b".split(), gg, g64)))
    glayers.GAT_BACKWARD = "fused"
