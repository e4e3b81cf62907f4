"""Single launch of the bench's dominant kernel at its bench size (4M x 8192), for ncu captures."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, multih_b200 as m
n = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_TOTAL
sc, pick = bench.make_workload(n)
ctx = m.Context(); ctx.set_geometry(sc.F, sc.pts)
d_pts, d_aff = ctx.upload(sc.pts, sc.aff)
p_pts, p_aff = ctx.upload(sc.pts[pick], sc.aff[pick])
d_hyp = torch.cat([ctx.hypotheses_from_host(sc.planes), ctx.haf_hypotheses(p_pts, p_aff)]).contiguous()
if os.environ.get('MH_FAST_CONFIG'): ctx.set_fast_config(int(os.environ['MH_FAST_CONFIG']))
o = {}
for _ in range(3):
    o = ctx.data_cost_fused(d_pts, d_hyp, kmax=0, want_list=False, out=o)
torch.cuda.synchronize()
print("done", o["inliers"].sum().item())
