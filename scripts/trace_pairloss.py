"""Per-CTA %globaltimer timeline of the pair-loss kernels: first-data latency, streaming loop, epilogue, tail, and the
entry / wait / end stamps of the combine kernel.  Needs the trace build of the library:

    python -m hic_gnn_b200.build --trace      # -> hic_gnn_b200/libhicgat_trace.so (-DHICGAT_TRACE)
    python scripts/trace_pairloss.py          # -> gpurun_out/trace_pairloss.json + raw per-CTA stamps (.npy)
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["HICGAT_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hic_gnn_b200", "libhicgat_trace.so")
import torch

from hic_gnn_b200 import _native as N
from hic_gnn_b200 import ops

dev = torch.device("cuda", 0)
lib = N.lib()
lib.hicgat_debug_set_trace.restype = C.c_int
lib.hicgat_debug_set_trace.argtypes = [C.c_void_p]
mode = ops._MODES["mse_moments"]
out = {}
for n, rows in [(9970, 9970), (49850, 6232), (9970, 1247)]:
    pitch = ops.WishTarget.pitch_for(n)
    data = torch.rand(rows, pitch, device=dev)
    coords = 0.3 * torch.randn(n, 3, device=dev)
    tgt = ops.WishTarget(data, n, 0, rows)
    c_mse = 4.0 / (float(n) * n)
    trace = torch.zeros(8192 * 8, dtype=torch.int64, device=dev)
    for _ in range(5):
        m, g = ops.pairloss_raw(coords, tgt, mode, c_mse, 0.0)
    torch.cuda.synchronize()
    assert lib.hicgat_debug_set_trace(trace.data_ptr()) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m, g = ops.pairloss_raw(coords, tgt, mode, c_mse, 0.0, m, g)
    e1.record()
    torch.cuda.synchronize()
    assert lib.hicgat_debug_set_trace(None) == 0
    t = trace.view(-1, 8).cpu()
    comb = t[8190:8192].clone()
    t = t[:8190]
    t = t[(t[:, 0] > 0) & (t[:, 5] > 0)]
    base = int(t[:, 0].min())
    ev_us = e0.elapsed_time(e1) * 1e3
    rel = lambda c: (t[:, c] - base).double() / 1e3
    start, first, loop_end, pub, strip_done, end = (rel(c) for c in range(6))
    kind = t[:, 7]
    def q(x):
        return [round(float(x.min()), 2), round(float(x.median()), 2), round(float(x.max()), 2)] if x.numel() else None
    rec = {
        "n": n, "rows": rows, "ctas": int(t.shape[0]), "event_us": ev_us, "span_us": float(end.max()),
        "first_data_minus_start_us": q(first - start), "loop_us": q(loop_end - first), "loop_end_us": q(loop_end),
        "publish_us": q(pub - loop_end), "after_ticket_us_strip_last": q((end - pub)[kind == 1]), "after_ticket_us_grid_last": q((end - pub)[kind == 2]),
        "after_ticket_us_others": q((end - pub)[kind == 0]), "tail_us (last loop end -> kernel end)": float(end.max() - loop_end.max()),
        "second_wave_start_us": float(start.sort().values[296]) if t.shape[0] > 296 else None,
    }
    import numpy as np
    np.save(f"gpurun_out/trace_raw_{n}x{rows}.npy", trace.view(-1, 8).cpu().numpy())
    rec["combine_grad_blk0 (entry, wait_done, end) us"] = [round((int(comb[0, k]) - base) / 1e3, 2) for k in range(3)]
    rec["combine_moments_blk (entry, wait_done, end) us"] = [round((int(comb[1, k]) - base) / 1e3, 2) for k in range(3)]
    out[f"{n}x{rows}"] = rec
    print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/trace_pairloss.json", "w"), indent=1)
