"""How often does the tensor-core fast path (filter on 3xTF32 products) disagree with the FP32 kernels, and on which hypotheses?"""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multih_b200 as m
N = 1 << 20
sc = m.scenes.make_scene(N, 200, seed=0xB200 + 3)
ctx = m.Context(); ctx.set_geometry(sc.F, sc.pts); pb, ab = ctx.upload(sc.pts, sc.aff)
hb = ctx.haf_hypotheses(pb, ab)
cfgs = [int(c) for c in sys.argv[1:]] or [61]
for seed in range(int(os.environ.get("PROBE_SEEDS", "12"))):
    g = torch.Generator("cuda").manual_seed(seed)
    idx = torch.randint(0, N, (7992,), device="cuda", generator=g)
    hyp = torch.cat([ctx.hypotheses_from_host(sc.planes), hb[idx]]).contiguous()
    ctx.set_fast_config(5)
    ref = ctx.data_cost_fused(pb, hyp, kmax=0, want_list=False, out={})
    rb, ri = ref["best"].clone(), ref["inliers"].clone()
    hmax = hyp[:, :9].abs().max(1).values
    for cfg in cfgs:
        ctx.set_fast_config(cfg)
        o = ctx.data_cost_fused(pb, hyp, kmax=0, want_list=False, out={})
        bad = (o["best"] != rb).nonzero().flatten()
        print(f"seed {seed} cfg {cfg}: {bad.numel()} rows differ; inlier maxdiff {(o['inliers'] - ri).abs().max().item()}; hyp |h|max: median {hmax.median().item():.3g} p99 {hmax.float().quantile(0.99).item():.3g} max {hmax.max().item():.3g}")
        for r in bad[:6].tolist():
            a, b = int(rb[r]), int(o["best"][r])
            la, lb = a & 0xffffffff, b & 0xffffffff
            rr = ctx.residuals(pb[r:r + 1].contiguous(), hyp[[max(la - 1, 0), max(lb - 1, 0)]].contiguous()).cpu().numpy()
            print(f"   row {r}: v3 (cost {a >> 32}, label {la}) tc (cost {b >> 32}, label {lb}); d2[px^2] {rr.ravel()}; |h|max {hmax[max(la-1,0)].item():.3g} {hmax[max(lb-1,0)].item():.3g}; h[v3 label] {hyp[max(la-1,0), :9].cpu().numpy()}")
