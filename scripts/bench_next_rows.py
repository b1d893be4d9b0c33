#!/usr/bin/env python
"""Timings of the 'next' rows on one B200 next to their CPU oracle: KR normalisation
(f-1), contact-list ingest (f-2), dSCC (f-3).  Writes one JSON object to stdout.  Not part of bench.py's
contract line; the numbers are recorded in profiles/r1_next_rows.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from hic_gnn_b200 import _native as N
from hic_gnn_b200 import kr, metrics, synth, utils
from hic_gnn_b200.ops import _stream


def gpu_time(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def main():
    out = {}
    # ---- f-1: gemv roofline + full KR
    for n in (9970, 30000):
        A = torch.rand(n, n, dtype=torch.float64, device="cuda")
        x = torch.rand(n, dtype=torch.float64, device="cuda")
        y = torch.empty_like(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lib = N.lib()
        for _ in range(3):
            lib.hicgat_gemv_f64(A.data_ptr(), A.stride(0), n, x.data_ptr(), y.data_ptr(), _stream())
        e0.record()
        for _ in range(10):
            lib.hicgat_gemv_f64(A.data_ptr(), A.stride(0), n, x.data_ptr(), y.data_ptr(), _stream())
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out[f"gemv_f64_n{n}"] = {"ms": ms, "GBps": n * n * 8 / ms / 1e6}
        del A
    raw = synth.raw_block(2493, 0, 2493, synth.solve_c0(2493, 0.95), 7, "cuda")
    t_gpu = gpu_time(lambda: kr.kr_norm(raw), reps=3)
    from oracle import kr as okr

    raw_np = raw.cpu().numpy()
    t0 = time.perf_counter()
    want = okr.kr_norm(raw_np)
    t_cpu = time.perf_counter() - t0
    got = kr.kr_norm(raw).cpu().numpy()
    out["kr_norm_n2493"] = {"gpu_s": t_gpu, "oracle_numpy_s": t_cpu, "max_abs_diff": float(np.abs(got - want).max()),
                            "frac_entries_differing": float((got != want).mean())}
    raw = synth.raw_block(9970, 0, 9970, synth.solve_c0(9970, 0.07), 7, "cuda")
    out["kr_norm_n9970"] = {"gpu_s": gpu_time(lambda: kr.kr_norm(raw), reps=2)}
    del raw
    # ---- f-2: list ingest on the shipped chr19 list and on a synthetic 10k-locus list
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
    lst = g["500kb_list"]
    out["convert_to_matrix_chr19_500kb"] = {"records": int(lst.shape[0]), "gpu_s": gpu_time(lambda: utils.convert_to_matrix(lst))}
    m = synth.synthetic_map(3000, 0.3, seed=1)
    iu = torch.triu_indices(3000, 3000, 1)
    v = m[iu[0], iu[1]]
    nz = v != 0
    big = torch.stack([iu[0][nz].double() * 5000, iu[1][nz].double() * 5000, v[nz]], 1).numpy()
    from oracle import graph as ograph

    t0 = time.perf_counter()
    want = ograph.convert_to_matrix(big)
    t_cpu = time.perf_counter() - t0
    got = utils.convert_to_matrix(big).cpu().numpy()
    out["convert_to_matrix_3000_loci"] = {"records": int(big.shape[0]), "gpu_s": gpu_time(lambda: utils.convert_to_matrix(big)),
                                          "oracle_numpy_s": t_cpu, "bit_exact": bool(np.array_equal(got, want))}
    # ---- f-3: dSCC
    from scipy.stats import spearmanr

    from oracle import loss as oloss
    from oracle import wish as owish

    for n, dens in ((2493, 0.95), (9970, 0.07)):
        adj = synth.synthetic_map_chunked(n, dens, device="cuda")
        truth = utils.cont2dist(adj, 1.0)
        coords = (0.3 * torch.randn(n, 3, generator=torch.Generator().manual_seed(3))).cuda()
        t_gpu = gpu_time(lambda: metrics.dscc(coords, truth), reps=3)
        rec = {"pairs": n * (n - 1) // 2, "gpu_s": t_gpu, "dscc": metrics.dscc(coords, truth)}
        if n <= 3000:
            tt, dd = oloss.triu_pairs(truth.cpu(), coords.cpu())
            t0 = time.perf_counter()
            rec["scipy"] = float(spearmanr(tt.numpy(), dd.numpy())[0])
            rec["scipy_s"] = time.perf_counter() - t0
        out[f"dscc_n{n}"] = rec
        del adj, truth
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
