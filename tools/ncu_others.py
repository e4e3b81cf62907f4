"""One launch of every other kernel family at a representative size, for ncu captures (K1, K2 dense, K3, K4)."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, multih_b200 as m
n = 1 << 20
sc, pick = bench.make_workload(n)
ctx = m.Context(); ctx.set_geometry(sc.F, sc.pts)
d_pts, d_aff = ctx.upload(sc.pts, sc.aff)
for _ in range(2):
    d_h = ctx.haf_hypotheses(d_pts, d_aff)                                   # K1
d_hyp = torch.cat([ctx.hypotheses_from_host(sc.planes), d_h[torch.from_numpy(pick[:824] % n).cuda()]]).contiguous()  # K = 1024
od = torch.empty((n, 1025), dtype=torch.int32, device="cuda")
for _ in range(2):
    ctx.data_cost_dense(d_pts, d_hyp, out=od)                                # K2 dense int32
od16 = torch.empty((n, 1025), dtype=torch.int16, device="cuda")
for _ in range(2):
    ctx.data_cost_dense(d_pts, d_hyp, elem_bytes=2, out=od16)               # K2 dense int16
lab = torch.from_numpy(sc.gt).cuda()
for _ in range(2):
    ctx.refit_haf(d_pts, d_aff, lab, 200)                                    # K4
f10 = ctx.features10(d_h[:20000].contiguous(), d_pts[:20000].contiguous())
cen, asg, st = ctx.meanshift(f10, 2.2)                                       # K3 (20k points, 10-D)
ctx.refit_3pt(d_pts[:20000].contiguous(), asg, cen.shape[0])                 # K4 3PT
torch.cuda.synchronize(); print("done", st)
