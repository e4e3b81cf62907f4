"""Host vs device neighbourhood timing (mh_neighbourhood, k = 31, radius 200)."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multih_b200 as m
ctx = m.Context()
for n in (1197, 5000, 100_000):
    sc = m.scenes.make_scene(n, 8, seed=3)
    for name, kw in (("host", {}), ("device", {"ctx": ctx, "backend": 2})):
        m.capi.neighbourhood(sc.pts, 200.0, 31, **kw)
        t = time.perf_counter()
        for _ in range(3):
            o, a = m.capi.neighbourhood(sc.pts, 200.0, 31, **kw)
        print(f"N {n} {name}: {(time.perf_counter() - t) / 3 * 1e3:.2f} ms per (count + fill) pair of calls, {len(a)} entries", flush=True)
