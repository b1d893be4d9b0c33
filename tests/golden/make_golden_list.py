"""Golden vectors for the list -> graph path: outputs of the REFERENCE's own
``utils.load_input`` fed 3-column contact lists (``utils.py:29-31`` -> ``convert_to_matrix``, ``utils.py:10-26``),
imported unmodified from /root/reference with the absent third-party modules stubbed exactly as in make_golden.py.

    python tests/golden/make_golden_list.py        (this container only: needs /root/reference)

Cases: the two shipped chr19 lists; a sparse upper-triangular list with gaps (bins that only carry zero counts are
dropped, utils.py:22-24), duplicate records ("last record wins"), diagonal records and records below the first
sub-diagonal (dropped by ``triu(mat) + tril(mat.T, 1)``, utils.py:21); and a list with first-sub-diagonal records, which the
reference turns into an ASYMMETRIC matrix (the direct path must refuse it).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF = "/root/reference"


def sparse_list(seed, nbins, nrec, res=5000, lower=False, subdiag=False):
    rng = np.random.default_rng(seed)
    bins = np.sort(rng.choice(np.arange(0, 4 * nbins) * res, size=nbins, replace=False)).astype(np.float64)
    i = rng.integers(0, nbins, nrec)
    d = np.minimum(rng.geometric(0.08, nrec), nbins - 1)
    j = np.minimum(i + d - 1, nbins - 1)              # includes diagonal records (d == 1)
    v = np.ceil(rng.random(nrec) * 50.0)
    v[rng.random(nrec) < 0.05] = 0.0                  # zero counts
    rec = np.stack([bins[i], bins[j], v], axis=1)
    rec = np.concatenate([rec, rec[: nrec // 20] * np.array([1.0, 1.0, 3.0])])  # duplicates: the later record wins
    dead = bins[rng.choice(nbins, 5, replace=False)]  # bins whose every record carries a zero count: dropped (utils.py:22-24)
    rec[np.isin(rec[:, 0], dead) | np.isin(rec[:, 1], dead), 2] = 0.0
    if lower:                                         # records well below the diagonal: dropped by triu + tril(.T, 1)
        k = rng.integers(0, nbins - 5, nrec // 10)
        rec = np.concatenate([rec, np.stack([bins[k + 4], bins[k], np.full(k.size, 7.0)], axis=1)])
    if subdiag:                                       # first sub-diagonal records: added onto the super-diagonal only
        k = rng.integers(0, nbins - 1, 10)
        rec = np.concatenate([rec, np.stack([bins[k + 1], bins[k], np.full(k.size, 11.0)], axis=1)])
    return rec[rng.permutation(len(rec))]


def main():
    import make_golden

    make_golden._install_stubs()
    sys.path.insert(0, REF)
    import utils as ref_utils  # the reference's utils.py, unmodified

    out = {}
    cases = {
        "1mb": np.loadtxt(f"{REF}/Data/GM12878_1mb_chr19_list.txt"),
        "500kb": np.loadtxt(f"{REF}/Data/GM12878_500kb_chr19_list.txt"),
        "sparse": sparse_list(1, 400, 6000),
        "sparse_lower": sparse_list(2, 300, 3000, lower=True),
        "sparse_subdiag": sparse_list(3, 120, 1500, subdiag=True),
    }
    for tag, lst in cases.items():
        mat = ref_utils.convert_to_matrix(lst)
        n = mat.shape[0]
        data = ref_utils.load_input(lst.copy(), np.zeros((n, 4), dtype=np.float32))  # reference output
        st = data.edge_index
        out[f"{tag}_list"] = lst
        out[f"{tag}_n"] = np.array([n])
        out[f"{tag}_rowptr"] = st.storage.rowptr().numpy()
        out[f"{tag}_col"] = st.storage.col().numpy()
        out[f"{tag}_val"] = st.storage.value().numpy()
        y = data.y
        out[f"{tag}_asym"] = np.array([float((y - y.t()).abs().max())])
        for f in (0.5, 1.0):
            w = ref_utils.cont2dist(y.clone(), f)     # reference output (dense; the direct path must reproduce it implicitly)
            out[f"{tag}_wish_{f}"] = w.numpy().astype(np.float32) if n <= 500 else np.zeros(0, dtype=np.float32)
        print(tag, "records", len(lst), "n", n, "nnz", st.storage.col().numel(), "asym", out[f"{tag}_asym"][0])
    np.savez_compressed(os.path.join(HERE, "reference_golden_list.npz"), **out)
    print("wrote", os.path.getsize(os.path.join(HERE, "reference_golden_list.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
