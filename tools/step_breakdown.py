"""Stage-by-stage CUDA-event breakdown of one bench step at a given per-rank size (no NCCL): where a small step's time goes."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, multih_b200 as m
n = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
sc, pick = bench.make_workload(n)
ctx = m.Context(); ctx.set_geometry(sc.F, sc.pts)
d_pts, d_aff = ctx.upload(sc.pts, sc.aff)
p_pts, p_aff = ctx.upload(sc.pts[pick % n], sc.aff[pick % n])
K = bench.K_HYP
d_hyp = torch.cat([ctx.hypotheses_from_host(sc.planes), ctx.haf_hypotheses(p_pts, p_aff)]).contiguous()
d_hyp_pt = torch.empty((n, 12), dtype=torch.float32, device="cuda")
fused = {"best": torch.empty(n, dtype=torch.int64, device="cuda"), "inliers": torch.empty(K, dtype=torch.int32, device="cuda")}
labels = torch.empty(n, dtype=torch.int32, device="cuda"); acc = torch.empty((K, 12), dtype=torch.float64, device="cuda")
d_ref = torch.empty((K, 12), dtype=torch.float32, device="cuda")
stages = [("K1 haf", lambda: ctx.haf_hypotheses(d_pts, d_aff, out=d_hyp_pt)),
          ("K2 fused", lambda: ctx.data_cost_fused(d_pts, d_hyp, kmax=0, want_list=False, out=fused)),
          ("labels", lambda: ctx.labels_from_best(fused["best"], labels)),
          ("copy hyp", lambda: d_ref.copy_(d_hyp)),
          ("K4 accumulate", lambda: ctx.refit_haf_accumulate(d_pts, d_aff, labels, K, out=acc)),
          ("pack", lambda: ctx.pack_inlier_counts(fused["inliers"], acc)),
          ("unpack", lambda: ctx.pack_inlier_counts(fused["inliers"], acc, unpack=True)),
          ("K4 solve", lambda: ctx.refit_haf_solve(acc, d_ref))]
def step():
    for _, f in stages: f()
for _ in range(3): step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)]
reps = 20; tot = np.zeros(len(stages)); whole = 0.0
for _ in range(reps):
    ev[0].record()
    for i, (_, f) in enumerate(stages):
        f(); ev[i + 1].record()
    torch.cuda.synchronize()
    tot += [ev[i].elapsed_time(ev[i + 1]) for i in range(len(stages))]
    whole += ev[0].elapsed_time(ev[-1])
for (name, _), t in zip(stages, tot / reps): print("%-16s %.3f ms" % (name, t))
print("whole step       %.3f ms (sum of stages %.3f)" % (whole / reps, tot.sum() / reps))
import time
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(reps): step()
cpu = (time.perf_counter() - t) / reps * 1e3; torch.cuda.synchronize()
print("host time to enqueue one step: %.3f ms" % cpu)
