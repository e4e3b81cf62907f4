"""K3 probe (run on the GPU box): mean-shift against the oracle at several sizes, with device timings.
usage: python tools/k3_probe.py [sizes...]   e.g. 146:6 1197:10 4096:10 20000:10"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multih_b200 as m
from oracle import oracle as orc

specs = sys.argv[1:] or ["146:6", "1197:10", "4096:10", "9000:10", "20000:10"]
metric = int(os.environ.get("MS_METRIC", "0"))
check = os.environ.get("MS_CHECK", "1") == "1"
ctx = m.Context(**({"meanshift_metric": metric} if metric else {}))
for spec in specs:
    n, d = (int(v) for v in spec.split(":"))
    if d == 10:
        sc = m.scenes.make_scene(n, 8, seed=0xB200) if n != 1197 else None
        if sc is None:
            g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "barrsmith_hotpath_input.npz"))
            pts, aff, F = g["pts"], g["aff"], g["F"]
        else:
            pts, aff, F = sc.pts, sc.aff, sc.F
        feat = orc.features10(orc.haf_hypotheses(pts, aff, F), pts, 0.005)
    else:
        sc = m.scenes.make_scene(max(8 * n, 4096), 8, seed=0xB200 + 1)
        H = orc.haf_hypotheses(sc.pts, sc.aff, sc.F)[:n]
        feat = orc.features6(H)
    d_feat = torch.from_numpy(feat).cuda()
    cen, asg, st = ctx.meanshift(d_feat, 2.2)   # warm-up (allocations)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        ctx.rng_state = 1
        cen, asg, st = ctx.meanshift(d_feat, 2.2)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    line = f"N={n} D={d}: C={cen.shape[0]} traj={st[0]} iters={st[1]}  {e0.elapsed_time(e1) / reps:.3f} ms/call (wall {wall:.3f})"
    if check and n <= 30000:
        ctx.rng_state = 1
        cen, asg, st = ctx.meanshift(d_feat, 2.2)
        t0 = time.perf_counter()
        co, ao, _, sto = orc.meanshift(feat, 2.2, metric=metric)
        tcpu = (time.perf_counter() - t0) * 1e3
        ok = st == sto and cen.shape[0] == co.shape[0]
        err = float(np.abs(cen.cpu().numpy() - co).max()) if ok and co.size else float("nan")
        agree = float((asg.cpu().numpy() == ao).mean())
        line += f" | oracle {sto} {tcpu:.0f} ms  same={ok} centre err {err:.2e} assign agree {agree:.5f}"
    print(line, flush=True)
