"""K3 L2 tensor-core member (meanshift_metric = 2) against the oracle's sequential L2 mean-shift (metric = 1): mode sets."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multih_b200 as m
from oracle import oracle as orc

def match(a, b, r):
    """fraction of the rows of a that have a row of b within r (L2)"""
    if len(a) == 0 or len(b) == 0:
        return 0.0
    hit = 0
    for i in range(0, len(a), 512):
        d = ((a[i:i + 512, None, :] - b[None, :, :]) ** 2).sum(-1)
        hit += int((d.min(1) < r * r).sum())
    return hit / len(a)

specs = sys.argv[1:] or ["146:6", "1197:10", "4096:10", "20000:10"]
ctx2 = m.Context(meanshift_metric=2)
ctx1 = m.Context(meanshift_metric=1)
for spec in specs:
    n, d = (int(v) for v in spec.split(":"))
    sc = m.scenes.make_scene(max(n, 4096) if d == 6 else n, 8, seed=0xB200)
    H = orc.haf_hypotheses(sc.pts, sc.aff, sc.F, threads=8)
    feat = orc.features10(H, sc.pts, 0.005) if d == 10 else orc.features6(H[:n])
    d_feat = torch.from_numpy(feat).cuda()
    cen, asg, st = ctx2.meanshift(d_feat, 2.2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cen, asg, st = ctx2.meanshift(d_feat, 2.2)
    torch.cuda.synchronize()
    t2 = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    c1, a1, st1 = ctx1.meanshift(d_feat, 2.2)
    torch.cuda.synchronize()
    t1 = (time.perf_counter() - t0) * 1e3
    cen, c1, asg, a1 = cen.cpu().numpy(), c1.cpu().numpy(), asg.cpu().numpy(), a1.cpu().numpy()
    big1 = np.bincount(a1[a1 >= 0], minlength=len(c1)) >= 3
    big2 = np.bincount(asg[asg >= 0], minlength=len(cen)) >= 3
    print(f"N={n} D={d}: gram C={len(cen)} iters={st[1]} {t2:.2f} ms | sequential L2 C={len(c1)} {t1:.2f} ms | "
          f"sequential centres with a gram centre within bw/2: {match(c1, cen, 1.1):.4f} (bw: {match(c1, cen, 2.2):.4f}), "
          f"gram centres with a sequential one within bw/2: {match(cen, c1, 1.1):.4f}; clusters >= 3 members: gram {int(big2.sum())} "
          f"sequential {int(big1.sum())}, of those matched within bw/2: {match(c1[big1], cen[big2], 1.1):.4f}", flush=True)
