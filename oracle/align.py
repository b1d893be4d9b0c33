"""CPU restatement of the generalisation helpers either side of the hot path (test infrastructure only).

* ``domain_alignment`` / ``domain_alignment_filtered`` -- utils.py:83-146: Procrustes alignment of the
  embeddings of a finer-resolution map onto those of the map the network was trained on.
* ``write_pdb`` -- utils.py:149-192.

Pinned against the reference's own outputs (tests/golden/make_golden_io.py -> reference_golden_io.npz).
scipy.linalg.orthogonal_procrustes (scipy 1.7.3, requirements.txt:29) is restated from its published
algorithm: ``u, w, vt = svd(B.T @ A).T)``, ``R = u @ vt``.
"""
from __future__ import annotations

import numpy as np


def _bin_ids(lst):
    idx = np.unique(np.asarray(lst)[:, 0]).astype(int)   # utils.py:84,87
    return idx, int(np.min(idx[1:] - idx[:-1]))          # utils.py:85,88


def orthogonal_procrustes(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """R minimising ||A R - B||_F over orthogonal R (scipy.linalg.orthogonal_procrustes)."""
    u, _, vt = np.linalg.svd(b.T.dot(a).T)
    return u.dot(vt)


def matched_rows(list1, list2, n1=None, n2=None):
    """Row indices (into embeddings2, embeddings1) of the bin pairs the reference aligns on
    (utils.py:90-100; with ``n1``/``n2`` the bounds filter of ``domain_alignment_filtered``, utils.py:126-131)."""
    idx1, diff1 = _bin_ids(list1)
    idx2, diff2 = _bin_ids(list2)
    bins = int(diff1 / (2 * diff2))                      # utils.py:90 (float division, truncated)
    a_rows, b_rows = [], []
    for i in range(bins + 1):
        aidx = np.where(np.isin(idx2 + i * diff2, idx1))[0]
        bidx = np.where(np.isin(idx1, idx2 + i * diff2))[0]
        if n2 is not None:
            aidx, bidx = aidx[aidx < n2], bidx[bidx < n1]
            if not (len(aidx) > 0 and len(bidx) > 0):
                continue
        a_rows.append(aidx)
        b_rows.append(bidx)
    if not a_rows:
        raise ValueError("No valid alignment indices found. Check your input data.")
    return np.concatenate(a_rows), np.concatenate(b_rows)


def domain_alignment(list1, list2, embeddings1, embeddings2, filtered: bool = False):
    e1, e2 = np.asarray(embeddings1), np.asarray(embeddings2)
    a_rows, b_rows = matched_rows(list1, list2, e1.shape[0], e2.shape[0]) if filtered else matched_rows(list1, list2)
    transform = orthogonal_procrustes(e2[a_rows], e1[b_rows])   # utils.py:105 / 143
    return np.matmul(e2, transform)                             # utils.py:106 / 144


def write_pdb(positions, ctype: str = "0") -> str:
    """The text ``utils.WritePDB`` writes (utils.py:149-192), as one string."""
    lines = ["\n"]
    n = len(positions)
    for i in range(1, n + 1):
        c2 = str(i)
        c4 = "B" + c2
        c5, c6, c7 = ("%.3f" % positions[i - 1][k] for k in range(3))
        lines.append("%s  %s   %s %s   %s%s%s  %s\n" % ("ATOM", c2.rjust(5), "CA MET", c4.ljust(6), c5.rjust(8), c6.rjust(8), c7.rjust(8), "0.20 10.00"))
    for i in range(1, n + 1):
        j = i + 1
        if j > n and ctype == "1":
            continue
        lines.append("%s%s%s\n" % ("CONECT", str(i).rjust(5), str(j).rjust(5)))
    lines.append("END")
    return "".join(lines)
