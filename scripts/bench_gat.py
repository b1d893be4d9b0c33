"""GAT gather kernels (csrc/conv.cu) alone on the synthetic maps: forward and fused backward, per `rows_per_cta` tuning.

    python scripts/bench_gat.py [--workloads c5,c4] [--rows 8,16] [--out profiles/r2_gat_rows.json]

CUDA events around 10 launches after 3 warm-ups; H = 2, C = 256 (every live GAT net of the reference)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="c5,c4")
    ap.add_argument("--rows", default="8,16", help="rows_per_cta[:l2_persist_mb] variants")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import bench
    from hic_gnn_b200 import _native as N, layers, ops, synth
    from hic_gnn_b200.graph import CSRGraph

    res = {}
    for name in args.workloads.split(","):
        w = bench.WORKLOADS[name]
        n = w["n"]
        adj = synth.synthetic_map_chunked(n, w["density"], device="cuda")
        rowptr, col, val = ops.csr_from_dense(adj)
        del adj
        graph = CSRGraph(rowptr, col, val, n)
        g = torch.Generator(device="cuda").manual_seed(1)
        conv = layers.GATConv(512, 256, heads=2).cuda()
        xl = torch.randn(n, 512, device="cuda", generator=g).requires_grad_(True)
        gout = torch.randn(n, 512, device="cuda", generator=g)
        entry = {"n": n, "edges_with_self_loops": int(graph.with_self_loops()[1].numel())}
        ref = None
        for spec in args.rows.split(","):
            rows, l2mb = (int(v) for v in (spec.split(":") + ["0"])[:2])
            N.check(N.lib().hicgat_gat_set_tuning(rows, l2mb), "hicgat_gat_set_tuning")
            fwd = lambda: layers._GatAttend.apply(xl, conv.att_l, conv.att_r, conv.bias, graph, 2, 256, 0.2)
            out = fwd()
            t_f = timed(lambda: fwd())
            t_fb = timed(lambda: torch.autograd.grad(fwd(), [xl, conv.att_l, conv.att_r, conv.bias], gout))
            grads = torch.autograd.grad(out, [xl, conv.att_l, conv.att_r, conv.bias], gout)
            if ref is None:
                ref = (out.detach().clone(), [t.clone() for t in grads])
                same = True
            else:  # the row -> warp mapping does not change any summation order: bit-identical
                same = bool(torch.equal(out, ref[0])) and all(bool(torch.equal(a, b)) for a, b in zip(grads, ref[1]))
            entry[f"rows{rows}" + (f"_l2persist{l2mb}mb" if l2mb else "")] = {"fwd_ms": round(t_f, 4), "fwd_bwd_ms": round(t_fb, 4), "bwd_ms": round(t_fb - t_f, 4), "bit_identical_to_first": same}
        N.check(N.lib().hicgat_gat_set_tuning(8, 0), "hicgat_gat_set_tuning")
        res[name] = entry
        print(name, json.dumps(entry), flush=True)
        del graph, xl, gout
        torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
