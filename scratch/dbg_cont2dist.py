import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hic_gnn_b200 import utils
g = np.load("tests/golden/reference_golden.npz")
for tag in ("1mb", "500kb"):
    y = torch.tensor(g[f"{tag}_y"], device="cuda")
    for f in (0.5, 1.0):
        got = utils.cont2dist(y, f).cpu().numpy()
        want = g[f"{tag}_wish_{f}"]
        bad = np.argwhere(got.view(np.uint64) != want.view(np.uint64))
        print(tag, f, "mismatch", len(bad), "of", got.size)
        for i, j in bad[:5]:
            a = g[f"{tag}_y"][i, j]
            print("  ", i, j, repr(a), got[i, j].hex(), want[i, j].hex(), "r", (1.0 / a).hex(), "max", got.max().hex(), want.max().hex())
