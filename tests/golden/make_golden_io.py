#!/usr/bin/env python
"""Golden vectors for the rows either side of the hot path, FROM THE REFERENCE ITSELF:

* ``utils.domain_alignment`` / ``domain_alignment_filtered`` (utils.py:83-146) on the shipped chr19 1 Mb / 500 kb
  contact lists with seeded stand-in embeddings (node2vec is not available; the function is agnostic to their origin);
* ``utils.WritePDB`` (utils.py:149-192) on the shipped structures: the bytes the reference writes for coordinates
  parsed from its own ``Outputs/*_structure.pdb`` -- which must reproduce those shipped files -- and for a seeded
  random structure with both ``ctype`` values.

Run in the build container only (needs /root/reference); the committed ``reference_golden_io.npz`` is what tests read.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (stand-ins for torch_geometric / torch_sparse + helpers)


def main():
    mg._install_stubs()
    sys.path.insert(0, mg.REF)
    import utils as ref_utils  # the reference's utils.py, unmodified

    out = {}
    l1 = np.loadtxt(f"{mg.REF}/Data/GM12878_1mb_chr19_list.txt")
    l2 = np.loadtxt(f"{mg.REF}/Data/GM12878_500kb_chr19_list.txt")
    n1, n2 = len(np.unique(l1[:, 0])), len(np.unique(l2[:, 0]))
    rng = np.random.default_rng(11)
    e1 = rng.standard_normal((n1, 48))
    q, _ = np.linalg.qr(rng.standard_normal((48, 48)))
    # the finer map's embeddings: a rotated, noisy, interleaved copy, so that the Procrustes problem is well posed
    e2 = np.repeat(e1, 2, axis=0)[:n2] @ q + 0.05 * rng.standard_normal((n2, 48))
    if e2.shape[0] < n2:
        e2 = np.vstack([e2, rng.standard_normal((n2 - e2.shape[0], 48))])
    out["align_emb1"], out["align_emb2"] = e1, e2
    out["align_fit"] = ref_utils.domain_alignment(l1, l2, e1, e2)
    out["align_fit_filtered"] = ref_utils.domain_alignment_filtered(l1, l2, e1, e2)
    # trained on the finer map, applied to the coarser one (bins = 0 branch)
    out["align_fit_swapped"] = ref_utils.domain_alignment_filtered(l2, l1, e2, e1)

    def written(pos, ctype="0"):
        with tempfile.NamedTemporaryFile("r", suffix=".pdb", delete=False) as f:
            path = f.name
        ref_utils.WritePDB(pos, path, ctype) if ctype != "0" else ref_utils.WritePDB(pos, path)
        data = open(path, "rb").read()
        os.unlink(path)
        return np.frombuffer(data, dtype=np.uint8)

    for tag, fname in (("1mb", "GM12878_1mb_chr19_list_structure.pdb"), ("500kb", "GM12878_500kb_chr19_list_generalized_structure.pdb")):
        shipped = open(f"{mg.REF}/Outputs/{fname}", "rb").read()
        xyz = mg.read_pdb(f"{mg.REF}/Outputs/{fname}")
        again = written(xyz)
        assert again.tobytes() == shipped, "WritePDB(read(shipped)) does not reproduce the shipped file"
        out[f"pdb_text_{tag}"] = again
    pos = rng.standard_normal((1203, 3)) * 400.0
    pos[5] = [0.0004, -0.0004, 99999.9996]  # rounding / width edge cases of '%.3f' in an 8-wide field
    pos[6] = [-12345.6789, 1e-9, -0.0005]
    out["pdb_rand_pos"] = pos
    out["pdb_rand_text_c0"] = written(pos, "0")
    out["pdb_rand_text_c1"] = written(pos, "1")
    np.savez_compressed(os.path.join(HERE, "reference_golden_io.npz"), **out)
    print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "reference_golden_io.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
