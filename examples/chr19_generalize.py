#!/usr/bin/env python
"""BASELINE.json configs[1]: GM12878 chr19 500 kb GAT generalisation from the 1 Mb model
(HiC_GAT_generalize_directly.py), end to end on one B200 with the drop-in modules:

    1 Mb and 500 kb contact lists          utils.convert_to_matrix -> kr.kr_norm         utils.py:10-26, r_utils.R:1-93
    train on the 1 Mb map                  utils.load_input / wish_target / train.fit    HiC_GAT_generalize_directly.py:182-260
    save / reload the weights              state_dict round trip (reference keys)        :282, :312-314
    align the 500 kb embeddings            utils.domain_alignment                        :316, utils.py:83-108
    500 kb structure from the 1 Mb model   utils.load_input + model.get_model            :317-323
    dSCC of the generalised structure      metrics.dscc                                  :335
    PDB file                               utils.WritePDB(coords * 100, ...)             :365

The contact lists come from tests/golden (copies of the reference's Data/ fixtures recorded by make_golden.py); node2vec
embeddings are replaced by seeded stand-ins (node2vec / gensim are not part of the hot path): the 500 kb stand-ins are a
rotated, noisy interleaving of the 1 Mb ones, so that the alignment has something to recover.

    python examples/chr19_generalize.py [--steps 300] [--out /tmp/chr19_500kb_generalized_structure.pdb]
"""
import argparse
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from hic_gnn_b200 import kr, metrics, models, train, utils


def normed_map(contacts):
    raw = utils.convert_to_matrix(contacts)
    raw.fill_diagonal_(0)                               # HiC_GAT_generalize_directly.py:118
    return kr.kr_norm(raw)                              # replaces the Rscript subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("-lr", type=float, default=1e-3)
    ap.add_argument("-thresh", type=float, default=1e-8)
    ap.add_argument("-conversion", type=float, default=1.0)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
    list_trained, list_untrained = g["1mb_list"], g["500kb_list"]
    normed_trained, normed_untrained = normed_map(list_trained), normed_map(list_untrained)
    n1, n2 = normed_trained.shape[0], normed_untrained.shape[0]
    rng = np.random.default_rng(42)
    emb_trained = 0.25 * rng.standard_normal((n1, 512))
    q, _ = np.linalg.qr(rng.standard_normal((512, 512)))
    emb_untrained = np.repeat(emb_trained, 2, axis=0)[:n2] @ q + 0.01 * rng.standard_normal((min(2 * n1, n2), 512))
    if emb_untrained.shape[0] < n2:
        emb_untrained = np.vstack([emb_untrained, 0.25 * rng.standard_normal((n2 - emb_untrained.shape[0], 512))])

    # ---- train on the 1 Mb map
    data = utils.load_input(normed_trained, emb_trained)
    target = utils.wish_target(data.y, args.conversion)
    torch.manual_seed(42)
    model = models.GATNetSelectiveResidualsUpdated().cuda()
    hist = train.fit(model, data.x.float(), data.edge_index, target, mode="mse_pearson", lr=args.lr, thresh=args.thresh, max_steps=args.steps,
                     use_cuda_graph=True, check_every=10)
    with torch.no_grad():
        dscc_trained = metrics.dscc(model.get_model(data.x.float(), data.edge_index), target)
    buf = io.BytesIO()
    torch.save(model.state_dict(), buf)                 # reference: torch.save(model.state_dict(), ..._weights.pt)

    # ---- generalise to the 500 kb map
    model = models.GATNetSelectiveResidualsUpdated().cuda()
    buf.seek(0)
    model.load_state_dict(torch.load(buf))
    model.eval()
    fitembed = utils.domain_alignment(list_trained, list_untrained, emb_trained, emb_untrained)
    data_fit = utils.load_input(normed_untrained, fitembed)
    target_fit = utils.wish_target(data_fit.y, args.conversion)
    with torch.no_grad():
        coords = model.get_model(data_fit.x.float(), data_fit.edge_index)
    dscc_generalised = metrics.dscc(coords, target_fit)
    if args.out:
        utils.WritePDB(coords * 100, args.out)
    print(f"trained on {n1} loci: steps={len(hist)} loss {hist[0]:.5f} -> {hist[-1]:.5f} dSCC={dscc_trained:.4f}; "
          f"generalised to {n2} loci: dSCC={dscc_generalised:.4f}" + (f"; wrote {args.out}" if args.out else ""))
    return hist, coords, dscc_trained, dscc_generalised


if __name__ == "__main__":
    main()
