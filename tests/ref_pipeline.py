"""Oracle-side restatement of MultiH::Process's control flow (MultiH/MultiH/MultiH.cpp:42-98, 224-312, 352-471,
513-602, 604-694) from ComputeLocalHomographies on, built from the oracle's FP64 functions and the REFERENCE's own
alpha-expansion (oracle/_ref).  Test infrastructure only: the parity tests run mh_process (GPU) and this on the same
inputs and report label agreement."""
import numpy as np

from oracle import oracle as orc


def _csr(assign, C):
    order = np.argsort(assign, kind="stable")
    order = order[assign[order] >= 0]
    counts = np.bincount(assign[assign >= 0], minlength=C)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    return offsets, order.astype(np.int32)


def oracle_process(pts, aff, F, thr=2.2, locality=0.005, lam=0.5, straightness=0.005, max_iterations=500,
                   convergence=1e-5, max_gc_cycles=1000, rng_state=1, max_neighbours=31, prefilter=False, use_ref_gco=True, expansion=None, trace=None,
                   compatibility_check=False, min_inliers=20, lm=False):
    if prefilter:                                                            # MultiH.cpp:807-838
        pts, aff, keepmask = orc.prefilter(pts, aff, F)
    N = len(pts)
    e2 = orc.epipole2(F)
    H_pt = orc.haf_hypotheses(pts, aff, F, e2)                              # MultiH.cpp:696-717
    f10 = orc.features10(H_pt, pts, locality)                               # :612-646
    centres, assign, rng_state, _ = orc.meanshift(f10, thr, 0, rng_state)   # :654
    offsets, members = _csr(assign, len(centres))
    Hc, keep = orc.cluster_3pt(pts, offsets, members, F, refine=lm)         # :664-688 (LM polish: 3PTcb.h, when lm)
    hyp = Hc[keep]
    if trace is not None:
        trace["K0"] = len(hyp)
    off, adj = orc.radius_neighbours(pts, 1.0 / locality, max_neighbours)                 # :231-253
    labeling = np.full(N, -1, dtype=np.int32)
    last_energy, not_changed, it, energy_final, k1_exit = float(2 ** 31 - 1), 0, 0, 0.0, False
    potts = orc.smooth_cost(0, 1, lam)
    if expansion is None:
        expansion = (lambda c, o, a, init: orc.gco_ref_expansion(c, potts, o, a, init, max_gc_cycles))
    while it < max_iterations:
        it += 1
        changed = False
        K = len(hyp)
        if K > 0:                                                            # MergingStep :352-471
            f6 = orc.features6(hyp)
            modes, _, rng_state, _ = orc.meanshift(f6, thr, 0, rng_state)
            Hm = np.stack([orc.mode_to_homography(mo, F, refine=lm).ravel() for mo in modes]) if len(modes) else np.zeros((0, 9))
            _, _, _, keep = orc.inlier_stats(pts, Hm, thr, straightness)
            merged = Hm[keep]
            changed = len(merged) != K
            if changed:
                hyp = merged
        not_changed = 0 if changed else not_changed + 1
        K = len(hyp)
        if K == 1:                                                           # :280-285; the stale labels of earlier steps are
            k1_exit = True                                                   # dropped as Process() does right after (:88-94)
            labeling = orc.inliers_of_homography(pts, hyp[0], thr, 0, np.full(N, -1, dtype=np.int32))
            break
        if K == 0:
            labeling = np.full(N, -1, dtype=np.int32)
            break
        cost = orc.data_cost_dense(pts, hyp, lam, thr, threads=8)           # LabelingStep :513-543
        init = None if changed else np.clip(labeling + 1, 0, K)
        energy, gl = expansion(cost, off, adj, init)
        labeling = (gl - 1).astype(np.int32)
        hyp, _, _ = orc.refit_haf(pts, aff, labeling, K, F, e2, H_init=hyp)  # :587-599 (linear solution)
        if trace is not None:
            trace.setdefault("K", []).append(K); trace.setdefault("E", []).append(energy)
        if (not changed and abs(last_energy - energy) < convergence) or not_changed > 10:   # :295
            energy_final = energy
            break
        last_energy = energy
    if compatibility_check and len(hyp) > 1:                                  # MultiH.cpp:78-86, 100-222
        labeling, hyp, _, _, rng_state = orc.compatibility_check(pts, labeling, hyp, F, thr, min_inliers, rng_state)
    if prefilter:
        full = np.full(len(keepmask), -2, dtype=np.int32)
        full[keepmask] = labeling
        labeling = full
    return labeling, hyp, dict(iterations=it - 1, energy=energy_final, k1_exit=k1_exit)  # final_iteration_number, MultiH.cpp:311
