"""One launch pair of the dense cost kernels at 1M x 1025 for ncu captures."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, multih_b200 as m
n = 1 << 20
sc, pick = bench.make_workload(n)
ctx = m.Context(); ctx.set_geometry(sc.F, sc.pts)
d_pts, d_aff = ctx.upload(sc.pts, sc.aff)
d_h = ctx.haf_hypotheses(d_pts, d_aff)
d_hyp = torch.cat([ctx.hypotheses_from_host(sc.planes), d_h[torch.from_numpy(pick[:824] % n).cuda()]]).contiguous()  # K = 1024
od = torch.empty((n, 1025), dtype=torch.int32, device="cuda")
od16 = torch.empty((n, 1025), dtype=torch.int16, device="cuda")
for _ in range(2):
    ctx.data_cost_dense(d_pts, d_hyp, out=od)
    ctx.data_cost_dense(d_pts, d_hyp, elem_bytes=2, out=od16)
torch.cuda.synchronize(); print("done")
