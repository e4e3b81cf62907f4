"""cfg3 (BASELINE configs[2]): 100k correspondences x 20 planes through mh_process, stage timings."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multih_b200 as m
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
sc = m.scenes.make_scene(n, 20, seed=0xB200 + 2)
ctx = m.Context()
t = time.perf_counter()
lab, H, K = ctx.process(sc.pts, sc.aff, sc.F)
dt = time.perf_counter() - t
ok = lab >= 0
purity = np.mean([np.bincount(sc.gt[(lab == k) & (sc.gt >= 0)], minlength=20).max() / max(1, ((lab == k) & (sc.gt >= 0)).sum()) for k in range(K)]) if K else 0
print(f"N={n}: {dt:.2f} s, K={K}, iterations={ctx.iterations}, outliers={np.mean(~ok):.3f} (generated 0.5), inlier purity={purity:.3f}", ctx.stage_ms(), ctx.alternating_ms(), flush=True)
