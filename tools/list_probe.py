"""Throughput of the list-producing K2 member (mh_data_cost_fused with d_list) on the bench scene slice: tensor-core list kernel
vs the first-generation emit-on-every-hit kernel (set_fused_variant(0)).  usage: python tools/list_probe.py [kmax...]"""
import os, sys
import numpy as np, torch
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
for planes, K in ((200, 8192), (20, 256)):
    sc = m.scenes.make_scene(N, planes, seed=0xB200 + 3)
    ctx = m.Context(); ctx.set_geometry(sc.F, sc.pts); pb, ab = ctx.upload(sc.pts, sc.aff)
    hb = ctx.haf_hypotheses(pb, ab)
    idx = torch.randint(0, N, (K - planes,), device="cuda", generator=torch.Generator("cuda").manual_seed(4))
    hyp = torch.cat([ctx.hypotheses_from_host(sc.planes), hb[idx]]).contiguous()
    for kmax in map(int, sys.argv[1:] or ["16", "32"]):
        ctx.set_fused_variant(0)
        ref = ctx.data_cost_fused(pb, hyp, kmax=kmax)
        ms0 = ev_time(lambda: ctx.data_cost_fused(pb, hyp, kmax=kmax, out=ref), reps=2, warm=1)
        ctx.set_fused_variant(1)
        o = ctx.data_cost_fused(pb, hyp, kmax=kmax)
        ms = ev_time(lambda: ctx.data_cost_fused(pb, hyp, kmax=kmax, out=o))
        same_cnt = torch.equal(o["count"], ref["count"]); same_best = torch.equal(o["best"], ref["best"])
        fit = ref["count"] <= kmax
        a = torch.sort(o["list"][fit].masked_fill(torch.arange(kmax, device="cuda")[None, :] >= o["count"][fit][:, None], -1), 1).values
        b = torch.sort(ref["list"][fit].masked_fill(torch.arange(kmax, device="cuda")[None, :] >= ref["count"][fit][:, None], -1), 1).values
        print(f"{planes} planes x K={K}, kmax {kmax}: tensor-core list kernel {ms:.3f} ms = {N * K / ms * 1e3:.3e} res/s | first generation {ms0:.3f} ms = "
              f"{N * K / ms0 * 1e3:.3e} | mean entries/site {ref['count'].float().mean().item():.1f}, sites that fit {fit.float().mean().item():.3f} | "
              f"same counts {same_cnt} same argmin {same_best} same lists (fitting sites) {torch.equal(a, b)}", flush=True)
