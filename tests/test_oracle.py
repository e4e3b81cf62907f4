"""CPU tests of the oracle (the FP64 restatement of the reference hot path): known-answer tests, the reference's integer
cost constants, and the golden vectors produced by tests/golden/make_golden.py (an independent transliteration on top of
the real OpenCV routines)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def scene(mh):
    return mh.scenes.make_scene(4000, 5, outlier_ratio=0.0, noise_px=0.0, noise_aff=0.0, seed=11)


def test_scene_is_epipolar_consistent(scene):
    x1 = np.c_[scene.pts[:, :2], np.ones(len(scene.pts))]
    x2 = np.c_[scene.pts[:, 2:], np.ones(len(scene.pts))]
    assert np.abs(np.einsum("ij,jk,ik->i", x2, scene.F, x1)).max() < 1e-9  # x2^T F x1 = 0


def test_sym_eigen_matches_numpy(orc):
    rng = np.random.default_rng(0)
    for n in (3, 4):
        A = rng.normal(size=(n, n)); A = A @ A.T
        w, V = orc.sym_eigen(A)
        wn = np.linalg.eigvalsh(A)[::-1]
        assert np.allclose(w, wn, rtol=1e-12, atol=1e-12)            # descending, like cv::eigen
        assert np.allclose(V @ A @ V.T, np.diag(w), atol=1e-10)       # eigenvectors in rows


def test_haf_known_answer(orc, scene):
    # noise-free plane: GetHomographyHAF (MultiH.cpp:850-911) must return the generating homography
    H = orc.haf_hypotheses(scene.pts, scene.aff, scene.F)
    ref = scene.planes[scene.gt]
    assert (np.abs(H - ref).max(1) / np.abs(ref).max(1)).max() < 1e-6


def test_3pt_and_refit_known_answer(orc, scene):
    for p in range(5):
        idx = np.where(scene.gt == p)[0][:12]
        H = orc.homography_3pt(scene.pts[idx, :2], scene.pts[idx, 2:], scene.F).ravel()
        assert np.abs(H / H[8] - scene.planes[p]).max() < 1e-7
    Hr, M10, cnt = orc.refit_haf(scene.pts, scene.aff, scene.gt, 5, scene.F)
    assert np.abs(Hr / Hr[:, 8:9] - scene.planes).max() < 1e-6
    assert cnt.sum() == len(scene.pts)


def test_cost_constants(orc, scene):
    # default parameters (lambda 0.5, thr 2.2): outlier label 4901, beyond truncation 9802, inliers 0..200 and
    # DEcreasing with distance; Potts 50 (MultiH.cpp:473-511, MultiH.h:41-44)
    c = orc.data_cost_dense(scene.pts[:50], scene.planes)
    assert (c[:, 0] == 4901).all()
    own = c[np.arange(50), scene.gt[:50] + 1]
    assert (own == 200).all()
    other = np.delete(c, 0, axis=1)
    assert set(np.unique(other)) <= set(range(0, 201)) | {9802}
    assert orc.smooth_cost(1, 2) == 50 and orc.smooth_cost(3, 3) == 0
    p = np.array([[0.0, 0.0, 3.0, 0.0]])
    I = np.eye(3).ravel()[None]
    d2, T = 9.0, 2.2 ** 2 * 81 / 16
    assert orc.data_cost_dense(p, I)[0, 1] == int(np.floor(200 * (1 - d2 / T) + 0.5))


def test_features_layout(orc, scene):
    H = scene.planes[:3]
    f6 = orc.features6(H)
    f10 = orc.features10(H, scene.pts[:3], 0.005)
    for k in range(3):
        h = H[k].reshape(3, 3)
        im = [h @ np.array(v) for v in ([0, 0, 1.0], [1, 0, 1.0], [0, 1, 1.0])]
        im = [v[:2] / v[2] for v in im]
        assert np.allclose(f6[k], np.concatenate(im))                                   # x1 y1 x2 y2 x3 y3
        assert np.allclose(f10[k, :6], [im[0][0], im[1][0], im[2][0], im[0][1], im[1][1], im[2][1]])  # x1 x2 x3 y1 y2 y3
        assert np.allclose(f10[k, 6:], 0.005 * scene.pts[k])


def test_meanshift_semantics(orc):
    # two well separated blobs; L1 window vs bw^2 (MeanShiftClustering.h:76-85); merge at bw/2 by averaging
    rng = np.random.default_rng(3)
    a = rng.normal(0, 0.05, size=(40, 6)); b = rng.normal(0, 0.05, size=(30, 6)) + 10.0
    X = np.concatenate([a, b])
    centres, assign, state, (traj, iters) = orc.meanshift(X, 2.2)
    assert centres.shape[0] == 2 and traj == 2
    assert len(set(assign[:40])) == 1 and len(set(assign[40:])) == 1 and assign[0] != assign[-1]
    assert np.allclose(sorted(centres[:, 0]), [a[:, 0].mean(), b[:, 0].mean()], atol=1e-6)
    # MSVC rand() restatement: first values of the unseeded generator are 41, 18467, 6334
    hold, vals = 1, []
    for _ in range(3):
        hold = (hold * 214013 + 2531011) & 0xFFFFFFFF
        vals.append((hold >> 16) & 0x7FFF)
    assert vals == [41, 18467, 6334]


def test_inlier_stats_straightness(orc, scene):
    # points of one plane restricted to a line in image 1 => lambda_min ~ 0 => rejected (MultiH.cpp:445-463)
    pts = scene.pts[scene.gt == 0][:200].copy()
    H = scene.planes[:1]
    cnt, sc, lmin, keep = orc.inlier_stats(pts, H)
    assert cnt[0] == 200 and keep[0]
    t = np.linspace(0, 1, 50)
    line = np.stack([100 + 300 * t, 50 + 200 * t], 1)
    h = H[0].reshape(3, 3)
    q = (h @ np.c_[line, np.ones(50)].T).T
    lp = np.c_[line, q[:, :2] / q[:, 2:]]
    cnt, sc, lmin, keep = orc.inlier_stats(lp, H)
    assert cnt[0] == 50 and lmin[0] < 0.005 and not keep[0]


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLD, "golden_small.npz")), reason="golden vectors not generated")
def test_golden_vectors(orc):
    """tests/golden/golden_small.npz was produced by make_golden.py with cv2.eigen / cv2.invert(DECOMP_SVD) — the
    OpenCV routines the reference calls (MultiH.cpp:893, 973, 1015, 1038)."""
    g = np.load(os.path.join(GOLD, "golden_small.npz"))
    pts, aff, F = g["pts"], g["aff"], g["F"]
    assert np.allclose(orc.epipole2(F), g["e2"], rtol=1e-9)
    H = orc.haf_hypotheses(pts, aff, F)
    rel = np.abs(H - g["haf_H"]).max(1) / np.abs(g["haf_H"]).max(1)
    assert np.percentile(rel, 99) < 1e-7 and rel.max() < 1e-4, (np.percentile(rel, 99), rel.max())
    assert np.allclose(orc.features10(g["haf_H"], pts, 0.005), g["feat10"], rtol=1e-12, atol=1e-12)
    assert np.allclose(orc.features6(g["haf_H"][:16]), g["feat6"], rtol=1e-12, atol=1e-12)
    assert np.array_equal(orc.data_cost_dense(pts, g["cost_H"]), g["cost"])
    for k in range(len(g["pt3_idx_padded"])):
        idx = g["pt3_idx_padded"][k]
        idx = idx[idx >= 0]
        H3 = orc.homography_3pt(pts[idx, :2], pts[idx, 2:], F).ravel()
        ref = g["pt3_H"][k]
        assert np.abs(H3 / H3[8] - ref / ref[8]).max() / np.abs(ref / ref[8]).max() < 1e-6
    Hr, _, cnt = orc.refit_haf(pts, aff, g["labels"], int(g["labels"].max()) + 1, F)
    ref = g["refit_H"]
    ok = cnt > 0
    rel = np.abs(Hr[ok] / Hr[ok][:, 8:9] - ref[ok] / ref[ok][:, 8:9]).max(1) / np.abs(ref[ok] / ref[ok][:, 8:9]).max(1)
    assert rel.max() < 1e-6
    Hm = np.stack([orc.mode_to_homography(m, F).ravel() for m in g["modes"]])
    ref = g["modes_H"]
    assert (np.abs(Hm / Hm[:, 8:9] - ref / ref[:, 8:9]).max(1) / np.abs(ref / ref[:, 8:9]).max(1)).max() < 1e-6
    cnt, sc, lmin, keep = orc.inlier_stats(pts, g["cost_H"])
    assert np.array_equal(cnt, g["inl_count"]) and np.array_equal(keep, g["inl_keep"])
    assert np.allclose(lmin, g["inl_lmin"], rtol=1e-6, atol=1e-9)


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLD, "golden_prefilter.npz")), reason="golden vectors not generated")
def test_prefilter_golden(orc):
    """Pre-filter (MultiH.cpp:786-838) vs the cv2.solvePoly-based transliteration on all 2903 bundled barrsmith rows and
    a noisy synthetic scene."""
    g = np.load(os.path.join(GOLD, "golden_prefilter.npz"))
    for name in ("barr", "syn"):
        po, ao, keep = orc.prefilter(g[f"{name}_pts"], g[f"{name}_aff"], g[f"{name}_F"])
        assert np.array_equal(keep, g[f"{name}_keep"])
        assert np.abs(po - g[f"{name}_out_pts"]).max() < 1e-8
        assert np.abs(ao - g[f"{name}_out_aff"]).max() < 1e-8
    # the hot-path input fixture is exactly the survivors of the F-RANSAC inliers
    assert g["barr_keep"].sum() >= 1197


def test_meanshift_terminates_on_a_cycling_trajectory(orc):
    """The reference's mean-shift loops `while (1)` (MeanShiftClustering.h:62) and never returns when the mean cycles; the
    oracle (and the kernel) cap a trajectory at MH_MS_MAX_WINDOW_ITERS window iterations.  This scene cycles."""
    import multih_b200 as m

    sc = m.scenes.make_scene(5000, 3 + (29 % 6), seed=0xB200 + 4 + 29)
    fo = orc.features10(orc.haf_hypotheses(sc.pts, sc.aff, sc.F), sc.pts, 0.005)
    cen, asg, _, st = orc.meanshift(fo, 2.2)
    assert st[0] > 0 and st[1] >= 200 and len(cen) > 10 and (asg >= 0).all()


def test_compatibility_check_golden(orc):
    """orc_compatibility_check against tests/golden/golden_compat.npz — an independent transliteration of
    HomographyCompatibilityCheck (MultiH.cpp:100-222) with Python lists for the std::vectors and the cv2-based 3-point fit
    (make_golden.py): same clusters removed, same rand() consumption, medians to 1e-6 (own SVD vs cv2.invert(DECOMP_SVD))."""
    g = np.load(os.path.join(GOLD, "golden_compat.npz"))
    for tag in ("a", "b"):
        lab, H, med, rem, rng = orc.compatibility_check(g["pts"], g["labels"], np.tile(np.eye(3).ravel(), (5, 1)), g["F"], 2.2,
                                                        int(g[f"{tag}_min_inliers"]), int(g[f"{tag}_seed"]))
        assert np.array_equal(rem, g[f"{tag}_removed"]) and rng == int(g[f"{tag}_rng"])
        gm = g[f"{tag}_medians"]
        ok = ~np.isnan(gm)
        assert np.array_equal(np.isnan(med), np.isnan(gm)) and np.allclose(med[ok], gm[ok], rtol=1e-6, atol=0), (med, gm)
        keep = np.where(~rem)[0]
        assert len(H) == len(keep) and set(np.unique(lab)) <= set(range(-1, len(keep)))
        for new, old in enumerate(keep):   # survivors keep their members, in order; removed clusters' members become outliers
            assert np.array_equal(lab == new, g["labels"] == old)


def test_meanshift_golden(orc):
    """orc_meanshift against tests/golden/golden_meanshift.npz — an independent numpy transliteration of
    MeanShiftClustering<double>::Cluster (MeanShiftClustering.h:22-157) with the MSVC rand() (make_golden.py): same
    trajectories, window iterations, rand() consumption and assignments; centres to 1e-9."""
    g = np.load(os.path.join(GOLD, "golden_meanshift.npz"))
    for tag in ("f10", "f6"):
        cen, asg, rng, stats = orc.meanshift(g[f"{tag}_data"], 2.2, 0, int(g[f"{tag}_seed"]))
        assert tuple(stats) == tuple(g[f"{tag}_stats"]) and rng == int(g[f"{tag}_rng"])
        assert cen.shape == g[f"{tag}_centres"].shape and np.allclose(cen, g[f"{tag}_centres"], rtol=1e-9, atol=1e-9)
        assert np.array_equal(asg, g[f"{tag}_assign"])


# ---- the oracle against the REFERENCE SOURCE compiled in place (oracle/_ref/libmultih_ref.so) ---------------------------------------
@pytest.fixture(scope="module")
def refmh(orc):
    if orc.ref_multih_lib() is None:
        pytest.skip("oracle/_ref/libmultih_ref.so was never built (needs /root/reference once)")
    return orc


def test_oracle_functions_equal_reference_source(refmh, scene):
    """GetHomographyHAF, dataEnergy / smoothnessEnergy, MeanShiftClustering<double>::Cluster, GetHomography3PT and
    GetHomographyHAFNonminimal of the reference's own MultiH.cpp (compiled unmodified against the mini OpenCV shim) against the
    oracle's restatement: identical integer costs, identical mean-shift (centres, assignments, rand() consumption), homographies
    to 1e-7 (two different Jacobi eigen-solvers / SVDs)."""
    orc = refmh
    pts, aff, F = scene.pts[:3000], scene.aff[:3000], scene.F
    Ho, Hr = orc.haf_hypotheses(pts, aff, F), orc.ref_haf_hypotheses(pts, aff, F)
    rel = np.abs(Hr / Hr[:, 8:9] - Ho / Ho[:, 8:9]).max(1) / np.abs(Ho / Ho[:, 8:9]).max(1)
    assert np.percentile(rel, 99) < 1e-7 and rel.max() < 1e-5, np.percentile(rel, [50, 99, 100])
    hyps = np.concatenate([scene.planes, Ho[:50]])
    assert np.array_equal(orc.ref_data_cost_dense(pts, hyps), orc.data_cost_dense(pts, hyps))
    assert np.array_equal(orc.ref_data_cost_dense(pts[:200], hyps, lam=0.3, thr=3.1), orc.data_cost_dense(pts[:200], hyps, 0.3, 3.1))
    assert orc.ref_multih_lib().ref_smooth_cost(0, 1, __import__("ctypes").c_double(0.5)) == orc.smooth_cost(0, 1) == 50
    f10 = orc.features10(Ho, pts, 0.005)
    for data, seed in ((f10, 1), (orc.features6(Ho[:400]), 1234)):
        cr, ar, rr = orc.ref_meanshift(data, 2.2, seed)
        co, ao, ro, _ = orc.meanshift(data, 2.2, 0, seed)
        assert rr == ro and cr.shape == co.shape and np.array_equal(ar, ao)
        assert np.abs(cr - co).max() <= 1e-9
    for label in range(3):
        idx = np.where(scene.gt[:3000] == label)[0][:40]
        H3r, H3o = orc.ref_homography_3pt(pts[idx, :2], pts[idx, 2:], F), orc.homography_3pt(pts[idx, :2], pts[idx, 2:], F)
        assert np.abs(H3r / H3r[2, 2] - H3o / H3o[2, 2]).max() < 1e-8
        lab = np.full(len(pts), -1, dtype=np.int32); lab[idx] = 0
        Hn = orc.ref_haf_nonminimal(pts[idx], aff[idx], F)
        Hno = orc.refit_haf(pts, aff, lab, 1, F)[0][0].reshape(3, 3)
        assert np.abs(Hn / Hn[2, 2] - Hno / Hno[2, 2]).max() < 1e-7
        # RefineHomographyHAF rebinds its local H (Homography_RefineHAFCallback.h:58) and never writes _H: the LM result is dropped
        Hn_lm = orc.ref_haf_nonminimal(pts[idx], aff[idx], F, refine=True)
        assert np.abs(Hn_lm / Hn_lm[2, 2] - Hn / Hn[2, 2]).max() < 1e-9


def test_oracle_pipeline_equals_reference_process(refmh):
    """MultiH::Process() of the reference source on the bundled pair (raw rows that pass the F test; F injected) against the
    oracle pipeline with the refinement filter and the compatibility check: same survivors, clusters, iterations, energy, labels."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from ref_pipeline import oracle_process

    orc = refmh
    g = np.load(os.path.join(GOLD, "golden_prefilter.npz"))
    F = g["barr_F"]
    x1 = np.c_[g["barr_pts"][:, :2], np.ones(len(g["barr_pts"]))]; x2 = np.c_[g["barr_pts"][:, 2:], np.ones(len(x1))]
    l = x1 @ F.T
    inl = np.abs(np.einsum("ij,ij->i", x2, l)) / np.hypot(l[:, 0], l[:, 1]) < 2.6
    pts, aff = g["barr_pts"][inl], g["barr_aff"][inl]
    for lm in (True, False):   # the reference as it is (its LM polish on), and with LM leaving the linear solutions untouched
        lab_r, H_r, info_r = orc.ref_process(pts, aff, F, lm=lm)
        lab_o, H_o, info_o = oracle_process(pts, aff, F, prefilter=True, compatibility_check=True, lm=lm)
        keep = lab_o > -2
        assert keep.sum() == len(lab_r) and not info_r["degenerate"]
        kp, _, _ = orc.prefilter(pts, aff, F)
        assert np.abs(kp - info_r["pts"]).max() < 1e-6                                   # the refined coordinates
        assert len(H_r) == len(H_o) and info_r["iterations"] == info_o["iterations"] and info_r["energy"] == info_o["energy"]
        assert np.array_equal(lab_r, lab_o[keep])
        assert np.abs(H_r / H_r[:, 8:9] - H_o / H_o[:, 8:9]).max() < 1e-6


def test_oracle_lm_equals_reference_lm(refmh, mh):
    """GetHomography3PT with do_numerical_refinement = true: the oracle's restatement of RefineHomography3PT + LMSolverImpl::run
    against the reference source, on clean clusters, contaminated clusters and three-point mode fits."""
    orc = refmh
    sc = mh.scenes.make_scene(3000, 5, seed=3)
    rng = np.random.default_rng(0)
    moved = 0
    for t in range(12):
        idx = np.where(sc.gt == t % 5)[0][:60] if t < 5 else rng.choice(len(sc.pts), 3 if t >= 9 else 30, replace=False)
        Hr = orc.ref_homography_3pt(sc.pts[idx, :2], sc.pts[idx, 2:], sc.F, refine=True)
        Ho = orc.homography_3pt(sc.pts[idx, :2], sc.pts[idx, 2:], sc.F, refine=True)
        Hl = orc.homography_3pt(sc.pts[idx, :2], sc.pts[idx, 2:], sc.F, refine=False)
        assert np.abs(Hr / Hr[2, 2] - Ho / Ho[2, 2]).max() < 1e-8
        moved += np.abs(Ho / Ho[2, 2] - Hl / Hl[2, 2]).max() > 1e-9
    for mo in orc.features6(sc.planes) + rng.normal(0, 0.7, (5, 6)):   # MergingStep's fits: three synthetic points
        Hr = orc.ref_homography_3pt(np.array([[0, 0], [1, 0], [0, 1.0]]), mo.reshape(3, 2), sc.F, refine=True)
        Ho = orc.mode_to_homography(mo, sc.F, refine=True)
        assert np.abs(Hr / Hr[2, 2] - Ho / Ho[2, 2]).max() < 1e-8
        moved += np.abs(Ho / Ho[2, 2] - orc.mode_to_homography(mo, sc.F) / orc.mode_to_homography(mo, sc.F)[2, 2]).max() > 1e-9
    print(f"\n[oracle] LM moved {moved} of 17 fits")
