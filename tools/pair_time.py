"""ms per image pair of mh_process on the bundled barrsmith fixture (and its alternating-stage breakdown)."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multih_b200 as m
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "barrsmith_hotpath_input.npz"))
ctx = m.Context()
ctx.process(g["pts"], g["aff"], g["F"])
best = 1e9
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    ctx.rng_state = 1
    t = time.perf_counter()
    lab, H, K = ctx.process(g["pts"], g["aff"], g["F"])
    best = min(best, (time.perf_counter() - t) * 1e3)
print(f"MH_GC_THREADS={os.environ.get('MH_GC_THREADS', 'default')}: best {best:.1f} ms/pair, planes {K}, iterations {ctx.iterations}, "
      f"labels checksum {int((lab.astype(np.int64) * np.arange(1, len(lab) + 1)).sum())}", ctx.stage_ms(), ctx.alternating_ms())
