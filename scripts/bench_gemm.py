#!/usr/bin/env python
"""The Linear GEMMs of the C5 training step (49 850 loci): hand-written tcgen05 3xTF32 path (split kernels + GEMM) next to the cuBLAS
fp32 GEMM it replaces, forward and both backward GEMMs, with the error of each against f64.  One JSON object on stdout."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from hic_gnn_b200 import ops


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps, out


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max())


def main():
    res = {}
    m = 49850
    g = torch.Generator(device="cuda").manual_seed(0)
    for nin, nout in ((512, 512), (512, 256), (256, 128), (128, 64)):
        x = 0.5 * torch.randn(m, nin, generator=g, device="cuda")
        w = torch.randn(nout, nin, generator=g, device="cuda") / nin**0.5
        gy = torch.randn(m, nout, generator=g, device="cuda")
        flops = 2.0 * m * nin * nout
        want = {"fwd": x.double() @ w.double().t(), "dx": gy.double() @ w.double(), "dW": gy.double().t() @ x.double()}
        ours = {
            "fwd": lambda: ops.gemm_tf32_tn(ops.split_tf32(x, 0b100), ops.split_tf32(w, 0b010)),
            "dx": lambda: ops.gemm_tf32_tn(ops.split_tf32(gy, 0b100), ops.split_tf32(w, 0b010, transpose=True)),
            "dW": lambda: ops.gemm_tf32_tn(ops.split_tf32(gy, 0b100, transpose=True), ops.split_tf32(x, 0b010, transpose=True)),
        }
        cublas = {"fwd": lambda: x @ w.t(), "dx": lambda: gy @ w, "dW": lambda: gy.t() @ x}
        xs, ws = ops.split_tf32(x, 0b100), ops.split_tf32(w, 0b010)
        gemm_only, _ = timed(lambda: ops.gemm_tf32_tn(xs, ws))
        for k in ("fwd", "dx", "dW"):
            t_o, y_o = timed(ours[k])
            t_c, y_c = timed(cublas[k])
            res[f"{nin}x{nout}_{k}"] = {"ours_ms": t_o, "cublas_fp32_ms": t_c, "ours_tflops_fp32_equiv": flops / t_o / 1e9, "cublas_tflops": flops / t_c / 1e9,
                                        "err_ours": rel(y_o, want[k]), "err_cublas": rel(y_c, want[k])}
        res[f"{nin}x{nout}_fwd"]["gemm_kernel_only_ms"] = gemm_only
        res[f"{nin}x{nout}_fwd"]["gemm_kernel_tf32_tflops"] = 3 * flops / gemm_only / 1e9
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
