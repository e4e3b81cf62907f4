import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multih_b200 as m
def ev_time(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
N = int(os.environ.get("TUNE_N", str(1 << 20)))
scb = m.scenes.make_scene(N, 200, seed=0xB200 + 3)
ctx = m.Context(); ctx.set_geometry(scb.F, scb.pts); pb, ab = ctx.upload(scb.pts, scb.aff)
hb = ctx.haf_hypotheses(pb, ab)
idx = torch.randint(0, N, (7992,), device="cuda", generator=torch.Generator("cuda").manual_seed(int(os.environ.get("TUNE_SEED", "4"))))
hyp = torch.cat([ctx.hypotheses_from_host(scb.planes), hb[idx]]).contiguous()
ctx.set_fused_variant(0)
lst = ctx.data_cost_fused(pb, hyp, kmax=1)   # list kernel (scalar) as the cross-check
ref = (lst["best"].clone(), lst["inliers"].clone())
ctx.set_fused_variant(1)
for cfg in map(int, sys.argv[1:] or ["0", "1", "2", "3"]):
    ctx.set_fast_config(cfg)
    o = ctx.data_cost_fused(pb, hyp, kmax=0, want_list=False, out={})
    ms = ev_time(lambda: ctx.data_cost_fused(pb, hyp, kmax=0, want_list=False, out=o))
    ms2 = ev_time(lambda: ctx.data_cost_fused(pb, hyp, kmax=0, want_list=False, want_inliers=False, out={"best": o["best"]}))
    if ref is None: ref = (o["best"].clone(), o["inliers"].clone())
    print("cfg %d: argmin+inliers %.3f ms %.3e res/s (%.1f%% of 73.2 TF) | argmin only %.3f ms %.3e | same best %s inl maxdiff %d" % (
        cfg, ms, N * 8192 / ms * 1e3, N * 8192 * 20 / ms * 1e3 / 73.2e12 * 100, ms2, N * 8192 / ms2 * 1e3, torch.equal(o["best"], ref[0]), (o["inliers"] - ref[1]).abs().max().item()), flush=True)
