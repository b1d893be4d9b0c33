#!/usr/bin/env python
"""A/B of the pair-loss kernel variants on one GPU over the row-block shapes that matter:
the full 50k / 10k / 2.5k-locus maps and the per-rank blocks of 2/4/8/16-way row sharding.

    python scripts/bench_pairloss_variants.py [--variants 0,0:512,1] [--iters 100] [--out gpurun_out/variants.json]

Per (shape, variant): mean / min CUDA-event duration of single launches, algorithmic GB/s (4 B per
ordered pair) and, across variants, the relative difference of moments and gradient.  Target values
are uniform random numbers (the arithmetic is data-independent); not a bench line.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from hic_gnn_b200 import _native as N
from hic_gnn_b200 import ops


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp(min=1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="0", help="comma list of variant[:rows_per_cta[:tail_depth[:tail_min_rows]]], e.g. 0,0:512,0:1024:2:256,1")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--mode", default="mse_moments")
    ap.add_argument("--sizes", default="49850,9970,2493")
    ap.add_argument("--worlds", default="1,2,4,8,16")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    variants = []
    for v in args.variants.split(","):
        f = [int(x) for x in v.split(":")] + [0, -1, 256]
        variants.append((f[0], f[1] if len(v.split(":")) > 1 else 0, f[2] if len(v.split(":")) > 2 else -1, f[3] if len(v.split(":")) > 3 else 256))
    mode = ops._MODES[args.mode]
    dev = torch.device("cuda", 0)
    rows_out = []
    for n in [int(s) for s in args.sizes.split(",")]:
        pitch = ops.WishTarget.pitch_for(n)
        data = torch.rand(n, pitch, device=dev)
        coords = 0.3 * torch.randn(n, 3, device=dev)
        c_mse, c_l1 = 4.0 / (float(n) * n), 0.1 / (n * (n - 1) / 2.0)
        for w in [int(x) for x in args.worlds.split(",")]:
            rows = -(-n // w)
            if rows < 512:
                continue
            tgt = ops.WishTarget(data[:rows], n, 0, rows)
            ref = None
            for v, rb, td, tm in variants:
                N.set_pairloss_tuning(rb, v)
                N.set_pairloss_schedule(td, tm)
                for _ in range(10):
                    m, g = ops.pairloss_raw(coords, tgt, mode, c_mse, c_l1)
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
                torch.cuda.synchronize()
                for a, b in ev:
                    a.record()
                    m, g = ops.pairloss_raw(coords, tgt, mode, c_mse, c_l1, m, g)
                    b.record()
                torch.cuda.synchronize()
                ts = sorted(a.elapsed_time(b) for a, b in ev)
                mean, lo, med = sum(ts) / len(ts), ts[0], ts[len(ts) // 2]
                gbs = rows * n * 4 / (med * 1e-3) / 1e9
                rec = {"n": n, "rows": rows, "variant": v, "rb": rb, "tail": [td, tm], "ms_mean": mean, "ms_med": med, "ms_min": lo, "gbs_med": gbs}
                if ref is None:
                    ref = (m.clone(), g.clone())
                else:
                    rec["dm"], rec["dg"] = rel(m, ref[0]), rel(g, ref[1])
                rows_out.append(rec)
                print(json.dumps(rec), flush=True)
            N.set_pairloss_tuning(0, 0)
            N.set_pairloss_schedule(-1, 256)
        del data
        torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows_out, f, indent=1)


if __name__ == "__main__":
    main()
