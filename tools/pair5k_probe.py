"""Stage timers of mh_process on one 5000-correspondence synthetic pair (the cfg5 unit of work), host expansion single-threaded."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MH_GC_THREADS", "1")
import numpy as np, multih_b200 as m
for i in (0, 1, 2, 3):
    sc = m.scenes.make_scene(5000, 3 + (i % 6), seed=0xB200 + 4 + i)
    ctx = m.Context()
    ctx.process(sc.pts, sc.aff, sc.F)
    t = time.perf_counter()
    lab, H, K = ctx.process(sc.pts, sc.aff, sc.F)
    dt = (time.perf_counter() - t) * 1e3
    print(f"pair {i}: {dt:.0f} ms, K={K}, iterations {ctx.iterations}, stages {ctx.stage_ms()}, alternating "
          f"{dict((k, round(v, 1)) for k, v in ctx.alternating_ms().items())}", flush=True)
