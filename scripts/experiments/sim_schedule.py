"""Fluid list-scheduling model of the pair-loss grid, fitted against measured kernel times."""
import ctypes as C, heapq, sys, itertools
import numpy as np
sys.path.insert(0, "/root/repo")
from hic_gnn_b200 import _native as N
lib = N.lib()

def sched(n, r0, r1):
    buf = (C.c_int32 * 600)()
    k = lib.hicgat_pairloss_describe_schedule(n, r0, r1, C.addressof(buf), 600)
    assert k > 0
    a = list(buf[:k]); ns, stg, c0, c1 = a[:4]
    return ns, stg, a[4:4 + c0 + 1], a[4 + c0 + 1:4 + c0 + 1 + c1 + 1]

def items_of(n, rows):
    ns, stg, b0, b1 = sched(n, 0, rows)
    nch = max(len(b0), len(b1)) - 1
    out = []
    for ch in range(nch):
        for st in range(ns):
            b = b1 if (stg and (st // 148) & 1) else b0
            if ch < len(b) - 1:
                out.append(b[ch + 1] - b[ch])
    return out  # dispatch order

def simulate(items, row_bytes=512.0, B=6.45e6, cap=60e3, w=(1.2, 0.8), p0=1.0, p1=1.5, fixed=6.0, slots=296):
    """B, cap in bytes/us.  Each slot: idle p0 -> stream (fluid share) -> idle p1 -> next item."""
    # state per slot: phase (0 pre,1 stream,2 post), remaining (us or bytes)
    nxt = 0
    t = 0.0
    phase = np.full(slots, -1); rem = np.zeros(slots); wt = np.array([w[0] if s < slots // 2 else w[1] for s in range(slots)])
    def start(s):
        nonlocal nxt
        if nxt < len(items):
            phase[s] = 0; rem[s] = p0; cur[s] = items[nxt] * row_bytes; nxt += 1
        else:
            phase[s] = -1
    cur = np.zeros(slots)
    for s in range(slots): start(s)
    while (phase >= 0).any():
        act = phase == 1
        rate = np.zeros(slots)
        if act.any():
            # water-filling with caps
            wsum = wt[act].sum()
            r = B * wt / wsum
            r = np.where(act, r, 0.0)
            over = r > cap
            # one round of redistribution is enough for our purpose
            if (over & act).any():
                excess = (r[over & act] - cap).sum()
                r[over & act] = cap
                free = act & ~over
                if free.any():
                    r[free] += excess * wt[free] / wt[free].sum()
                    r[free] = np.minimum(r[free], cap)
            rate = r
        # time to next event
        dt = np.inf
        timed = (phase == 0) | (phase == 2)
        if timed.any(): dt = min(dt, rem[timed].min())
        if act.any(): dt = min(dt, (rem[act] / rate[act]).min())
        t += dt
        rem[timed] -= dt
        rem[act] -= rate[act] * dt
        for s in np.nonzero((phase >= 0) & (rem <= 1e-9))[0]:
            if phase[s] == 0: phase[s] = 1; rem[s] = cur[s]
            elif phase[s] == 1: phase[s] = 2; rem[s] = p1
            else: start(s)
    return t + fixed

if __name__ == "__main__":
    meas = []
    def cfg(n, rows, rb, td, tm, ms):
        N.set_pairloss_tuning(rb, 0); N.set_pairloss_schedule(td, tm)
        meas.append(((n, rows, rb, td, tm), items_of(n, rows), ms))
    for rb, td, tm, ms in [(0, -1, 256, 91.5), (0, 1, 256, 88.6), (0, 2, 256, 91.5), (0, 3, 256, 87.2), (1024, 2, 256, 87.2), (1024, 3, 128, 85.3), (2048, 3, 256, 101.7), (2048, 4, 128, 98.6)]:
        cfg(9970, 9970, rb, td, tm, ms)
    for rb, td, tm, ms in [(0, -1, 256, 225.4), (0, 1, 256, 222.6), (0, 2, 256, 228.7), (0, 3, 256, 232.7), (1536, 3, 256, 238.9), (3136, 4, 256, 226.7)]:
        cfg(49850, 6232, rb, td, tm, ms)
    for kw in [dict(), dict(w=(1.0, 1.0)), dict(p0=2.0, p1=3.0), dict(cap=40e3), dict(cap=100e3), dict(w=(1.3, 0.7), p0=1.5, p1=2.5, cap=45e3)]:
        print("params", kw)
        for (key, items, ms) in meas:
            rb_bytes = 512.0
            t = simulate(items, **kw)
            print("   ", key, "items", len(items), "sim %.1f meas %.1f" % (t, ms))
