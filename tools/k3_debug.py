"""Debug aid: the 6-D mean-shift input of test_k3_meanshift_and_k4_3pt_vs_oracle, GPU vs a step-by-step numpy emulation."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multih_b200 as m
from oracle import oracle as orc

sc = m.scenes.make_scene(3000, 6, seed=7)
fo = orc.features10(orc.haf_hypotheses(sc.pts, sc.aff, sc.F), sc.pts, 0.005)
co, ao, _, sto = orc.meanshift(fo, 2.2)
C = co.shape[0]
order = np.argsort(ao, kind="stable"); order = order[ao[order] >= 0]
offs = np.concatenate([[0], np.cumsum(np.bincount(ao[ao >= 0], minlength=C))]).astype(np.int32)
H3o, keep_o = orc.cluster_3pt(sc.pts, offs, order.astype(np.int32), sc.F)
f6 = orc.features6(H3o[keep_o])
if len(sys.argv) > 1:
    f6 = f6[: int(sys.argv[1])]
N, D = f6.shape
B, stop = 2.2 ** 2, 2.2e-3
ctx = m.Context()
cen, asg, st = ctx.meanshift(torch.from_numpy(f6).cuda(), 2.2)
cen = cen.cpu().numpy(); asg = asg.cpu().numpy()
c6o, a6o, _, st6o = orc.meanshift(f6, 2.2)
print("gpu", st, cen.shape, "oracle", st6o, c6o.shape)


def traj(s):
    mean = f6[s].copy(); it = 0; votes = np.zeros(N, int)
    while True:
        mem = np.abs(mean - f6).sum(1) < B; votes += mem; it += 1
        nm = f6[mem].sum(0) / mem.sum()
        d = np.sqrt(((nm - mean) ** 2).sum()); mean = nm
        if d < stop or it >= 200:
            break
    return mean, it, votes


hold = 1; visited = np.zeros(N, bool); cents = []; log = []
while (~visited).any():
    hold = (hold * 214013 + 2531011) & 0xffffffff; r = (hold >> 16) & 0x7fff
    un = np.where(~visited)[0]; idx = un[int(round(r / 32767.0 * (len(un) - 1)))]
    mean, it, votes = traj(idx); visited |= votes > 0
    mw = -1
    for k, c in enumerate(cents):
        if np.sqrt(((mean - c) ** 2).sum()) < 1.1:
            mw = k; break
    if mw >= 0:
        cents[mw] = 0.5 * (cents[mw] + mean)
    else:
        cents.append(mean)
    log.append((idx, it, int((votes > 0).sum()), mw if mw >= 0 else len(cents) - 1, len(un)))
cents = np.array(cents)
print("emulation", len(log), sum(l[1] for l in log), cents.shape)
k = 0
while k < min(len(cents), len(cen)) and np.abs(cents[k] - cen[k]).max() < 1e-6:
    k += 1
print("first differing centre", k)
if k < len(cents):
    print("emulated centre", cents[k]); print("gpu centre     ", cen[k] if k < len(cen) else None)
    steps = [i for i, l in enumerate(log) if l[3] == k]
    print("steps that wrote it (step, seed, iters, members, cid, remaining):", [(i,) + log[i] for i in steps][:5])
    s0 = steps[0]
    print("log around:", log[max(0, s0 - 2): s0 + 3])
    seed = log[s0][0]
    print("seed row", f6[seed], "sorted pos", int(np.argsort(f6[:, 0], kind="stable").tolist().index(seed)))
print("assign equal", (asg == a6o).mean())
