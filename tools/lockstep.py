"""Lock-step comparison GPU stage vs oracle stage on the oracle's trajectory (same inputs at every stage)."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import multih_b200 as m
from oracle import oracle as orc

def rel(H, Ho):
    H = H / H[:, 8:9]; Ho = Ho / Ho[:, 8:9]
    return np.abs(H - Ho).max(1) / np.abs(Ho).max(1)

which = sys.argv[1] if len(sys.argv) > 1 else "synthetic"
if which == "synthetic":
    sc = m.scenes.make_scene(3000, 6, seed=7); pts, aff, F = sc.pts, sc.aff, sc.F; locality = 1 / 20.0
else:
    g = np.load("tests/golden/barrsmith_hotpath_input.npz"); pts, aff, F = g["pts"], g["aff"], g["F"]; locality = 0.005
thr, lam = 2.2, 0.5
N = len(pts)
ctx = m.Context(m.capi.default_params(locality=locality)); ctx.set_geometry(F, pts)
d_pts, d_aff = ctx.upload(pts, aff)
e2 = orc.epipole2(F)
H_pt = orc.haf_hypotheses(pts, aff, F, e2)
Hg = ctx.hypotheses_to_host(ctx.haf_hypotheses(d_pts, d_aff), True)
print("K1 rel p99/max", np.percentile(rel(Hg, H_pt), 99), rel(Hg, H_pt).max())
f10 = orc.features10(H_pt, pts, locality)
f10g = ctx.features10(ctx.haf_hypotheses(d_pts, d_aff), d_pts).cpu().numpy()
print("feat10 abs p99/max", np.percentile(np.abs(f10 - f10g).max(1), 99), np.abs(f10 - f10g).max())
rng = 1
centres, assign, rng_o, st_o = orc.meanshift(f10, thr, 0, rng)
ctx.rng_state = 1
cg, ag, st_g = ctx.meanshift(torch.from_numpy(f10).cuda(), thr)
print("MS10 same-features: stats", st_g, st_o, "C", cg.shape[0], len(centres), "assign agree", (ag.cpu().numpy() == assign).mean())
ctx2 = m.Context(m.capi.default_params(locality=locality)); ctx2.set_geometry(F, pts)
cg2, ag2, st_g2 = ctx2.meanshift(torch.from_numpy(f10g).cuda(), thr)
print("MS10 gpu-features : stats", st_g2, st_o, "C", cg2.shape[0], len(centres), "assign agree", (ag2.cpu().numpy() == assign).mean() if cg2.shape[0] == len(centres) else None)
C = len(centres)
order = np.argsort(assign, kind="stable"); order = order[assign[order] >= 0]
offs = np.concatenate([[0], np.cumsum(np.bincount(assign[assign >= 0], minlength=C))]).astype(np.int32)
Hc, keep = orc.cluster_3pt(pts, offs, order.astype(np.int32), F)
d_h3, keep_g = ctx.refit_3pt(d_pts, torch.from_numpy(assign).cuda(), C)
print("3PT keep agree", (keep_g.cpu().numpy().astype(bool) == keep).mean(), "rel p50/p99/max", np.percentile(rel(ctx.hypotheses_to_host(d_h3)[keep], Hc[keep]), [50, 99, 100]))
hyp = Hc[keep]
off, adj = orc.radius_neighbours(pts, 1.0 / locality, 31)
labeling = np.full(N, -1, dtype=np.int32); last_e = 2.0 ** 31; not_changed = 0
rng = rng_o
for it in range(1, 60):
    K = len(hyp)
    f6 = orc.features6(hyp)
    d_hyp = ctx.hypotheses_from_host(hyp)
    f6g = ctx.features6(d_hyp).cpu().numpy()
    modes, _, rng2, st = orc.meanshift(f6, thr, 0, rng)
    ctx.rng_state = rng
    mg, _, stg = ctx.meanshift(torch.from_numpy(f6).cuda(), thr)
    ctx.rng_state = rng
    mg2, _, stg2 = ctx.meanshift(torch.from_numpy(f6g).cuda(), thr)
    rng = rng2
    Hm = np.stack([orc.mode_to_homography(mo, F).ravel() for mo in modes])
    Hmg = ctx.hypotheses_to_host(ctx.modes_to_hypotheses(torch.from_numpy(modes).cuda()))
    cnt, _, lmin, keepm = orc.inlier_stats(pts, Hm, thr, 0.005)
    scg, lming, keepg = ctx.inlier_stats(d_pts, ctx.hypotheses_from_host(Hm))
    merged = Hm[keepm]; changed = len(merged) != K
    if changed: hyp = merged
    not_changed = 0 if changed else not_changed + 1
    K = len(hyp)
    msg = f"it {it}: K_in {len(f6)} feat6 maxdiff {np.abs(f6 - f6g).max():.2e} | MS same-feat C {mg.shape[0]}/{len(modes)} stats {stg}/{st} | gpu-feat C {mg2.shape[0]} | modes->H rel max {rel(Hmg, Hm).max():.1e} | inl cnt maxdiff {np.abs(scg[:,5]-cnt).max():.0f} keep agree {(keepg==keepm).mean():.3f} | K {K}"
    if K <= 1: print(msg); break
    cost = orc.data_cost_dense(pts, hyp, lam, thr, threads=8)
    costg = ctx.data_cost_dense(d_pts, ctx.hypotheses_from_host(hyp)).cpu().numpy()
    init = None if changed else np.clip(labeling + 1, 0, K)
    e, gl = orc.gco_ref_expansion(cost, 50, off, adj, init)
    glg, eg = m.capi.alpha_expansion(costg, 50, off, adj, init)
    labeling = (gl - 1).astype(np.int32)
    hyp_new, _, cnt_o = orc.refit_haf(pts, aff, labeling, K, F, e2, H_init=hyp)
    d_hr, _ = ctx.refit_haf(d_pts, d_aff, torch.from_numpy(labeling).cuda(), K, d_hyp=ctx.hypotheses_from_host(hyp))
    ok = cnt_o > 0
    r = rel(ctx.hypotheses_to_host(d_hr)[ok], hyp_new[ok])
    print(msg + f" | cost mismatch {(cost != costg).sum()} | label agree {(gl == glg).mean():.4f} E {e}/{eg} | refit rel max {r.max():.1e} (min cnt {cnt_o[ok].min()})")
    hyp = hyp_new
    if (not changed and abs(last_e - e) < 1e-5) or not_changed > 10: break
    last_e = e
