"""ctypes front-end of the CPU FP64 oracle (TEST INFRASTRUCTURE — see multih_oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Nothing here reads /root/reference at run time; `build()` compiles oracle/_ref from it only when it is present
(the prebuilt oracle/_ref/libgco_ref.so and oracle/_ref/libmultih_ref.so travel to the GPU box).

oracle/_ref/libmultih_ref.so is the REFERENCE ITSELF: MultiH.cpp, MeanShiftClustering.h, the refinement callbacks and the
alpha-expansion compiled unmodified, in place, against a mini OpenCV shim (oracle/cvshim/mini_cv.hpp, oracle/ref_multih_wrapper.cpp).
The `ref_*` functions below call it; tests/test_oracle.py pins the restatement on them.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libmultih_oracle.so")
_REF = os.path.join(_HERE, "_ref", "libgco_ref.so")
_REFMH = os.path.join(_HERE, "_ref", "libmultih_ref.so")

_lib = None
_ref = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_lp = C.POINTER(C.c_int64)


def build(force: bool = False) -> None:
    """Compile the restatement (always, if stale) and oracle/_ref (only when the reference tree is mounted)."""
    src = os.path.join(_HERE, "multih_oracle.cpp")
    stale = force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src)
    ref_missing = (not os.path.exists(_REF) or not os.path.exists(_REFMH) or
                   os.path.getmtime(_REFMH) < max(os.path.getmtime(os.path.join(_HERE, "ref_multih_wrapper.cpp")),
                                                  os.path.getmtime(os.path.join(_HERE, "cvshim", "mini_cv.hpp")))) \
        and os.path.isdir("/root/reference")
    if stale or ref_missing:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_data_cost_sweep.restype = C.c_int64
        _lib.orc_radius_neighbours.restype = C.c_int64
    return _lib


def ref_lib():
    """The reference's own alpha-expansion (GCO) compiled in place — None if it was never built."""
    global _ref
    if _ref is None:
        build()
        if not os.path.exists(_REF):
            return None
        _ref = C.CDLL(_REF)
        _ref.gco_ref_expansion.restype = C.c_int64
    return _ref


_refmh = None


def ref_multih_lib():
    """The reference's hot-path sources compiled in place (oracle/_ref/libmultih_ref.so) — None if never built."""
    global _refmh
    if _refmh is None:
        build()
        if not os.path.exists(_REFMH):
            return None
        _refmh = C.CDLL(_REFMH)
    return _refmh


class _StdoutToStderr:
    """The reference source reports its progress with printf; keep the caller's stdout clean (bench.py prints ONE JSON line there)
    by pointing file descriptor 1 at stderr while it runs."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        try:
            C.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self._saved, 1)
        os.close(self._saved)


def ref_process(pts, aff, F, thr_fund=2.6, thr=2.2, locality=0.005, lam=0.5, min_inliers=20, rng_state=1, lm=True):
    """MultiH::Process() of the reference source (MultiH.cpp:32-98) with F injected in place of its RANSAC, MSVC rand(), the exact
    31-nearest neighbourhood for FLANN, and its LM solver on (lm=True) or leaving the linear solutions untouched (lm=False).
    Returns (labels of the kept correspondences, H K x 9, dict(pts, haf, iterations, energy, degenerate))."""
    pts, pp = _d(pts); aff, pa = _d(aff); F, pf = _d(np.asarray(F).reshape(9))
    N = len(pts)
    lab = np.full(N, -9, dtype=np.int32); H = np.zeros((4096, 9)); po = np.zeros((N, 4)); ho = np.zeros((N, 9))
    K, kept, it, dg, en = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_double(0)
    with _StdoutToStderr():
        rc = ref_multih_lib().ref_multih_process(pp, pa, pf, N, C.c_double(thr_fund), C.c_double(thr), C.c_double(locality),
                                                  C.c_double(lam), int(min_inliers), C.c_uint(rng_state), int(bool(lm)),
                                                  lab.ctypes.data_as(c_ip), H.ctypes.data_as(c_dp), 4096, C.byref(K),
                                                  po.ctypes.data_as(c_dp), ho.ctypes.data_as(c_dp), C.byref(kept), C.byref(it),
                                                  C.byref(en), C.byref(dg))
    if rc:
        raise RuntimeError("the reference's Process() refused the input (fewer than 8 correspondences)")
    M = kept.value
    return lab[:M].copy(), H[: K.value].copy(), dict(pts=po[:M].copy(), haf=ho[:M].copy(), iterations=it.value, energy=en.value,
                                                      degenerate=bool(dg.value))


def ref_haf_hypotheses(pts, aff, F):
    pts, pp = _d(pts); aff, pa = _d(aff); F, pf = _d(np.asarray(F).reshape(9))
    H = np.zeros((len(pts), 9))
    ref_multih_lib().ref_haf_hypotheses(pp, pa, pf, len(pts), H.ctypes.data_as(c_dp))
    return H


def ref_homography_3pt(pts1, pts2, F, refine=False):
    p1, a = _d(pts1); p2, b = _d(pts2); F, pf = _d(np.asarray(F).reshape(9))
    H = np.zeros(9)
    ref_multih_lib().ref_homography_3pt(a, b, len(p1), pf, int(refine), H.ctypes.data_as(c_dp))
    return H.reshape(3, 3)


def ref_haf_nonminimal(pts, aff, F, refine=False):
    pts, pp = _d(pts); aff, pa = _d(aff); F, pf = _d(np.asarray(F).reshape(9))
    H = np.zeros(9)
    ref_multih_lib().ref_haf_nonminimal(pp, pa, len(pts), pf, int(refine), H.ctypes.data_as(c_dp))
    return H.reshape(3, 3)


def ref_data_cost_dense(pts, H, lam=0.5, thr=2.2):
    pts, pp = _d(pts); H, ph = _d(np.asarray(H).reshape(-1, 9))
    out = np.zeros((len(pts), len(H) + 1), dtype=np.int32)
    ref_multih_lib().ref_data_cost_dense(pp, len(pts), ph, len(H), C.c_double(lam), C.c_double(thr), out.ctypes.data_as(c_ip))
    return out


def ref_meanshift(data, bw, rng_state=1):
    data, pd = _d(data)
    N, D = data.shape
    centres = np.zeros((N, D)); assign = np.zeros(N, dtype=np.int32); st = C.c_uint(rng_state)
    Cn = ref_multih_lib().ref_meanshift(pd, N, D, C.c_double(bw), C.byref(st), centres.ctypes.data_as(c_dp), assign.ctypes.data_as(c_ip))
    return centres[:Cn].copy(), assign, int(st.value)


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_ip)


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(c_lp)


def hardware_threads() -> int:
    return int(lib().orc_hardware_threads())


def sym_eigen(A):
    A, pA = _d(A)
    n = A.shape[0]
    w = np.empty(n)
    V = np.empty((n, n))
    lib().orc_sym_eigen(n, pA, w.ctypes.data_as(c_dp), V.ctypes.data_as(c_dp))
    return w, V


def epipole2(F):
    F, pF = _d(F)
    e = np.empty(2)
    lib().orc_epipole2(pF, e.ctypes.data_as(c_dp))
    return e


def haf_hypotheses(pts, aff, F, e2=None, threads=1):
    pts, pp = _d(pts)
    aff, pa = _d(aff)
    F, pF = _d(F)
    e2 = epipole2(F) if e2 is None else np.asarray(e2, dtype=np.float64)
    e2, pe = _d(e2)
    N = pts.shape[0]
    H = np.empty((N, 9))
    lib().orc_haf_hypotheses(pp, pa, pF, pe, C.c_int64(N), H.ctypes.data_as(c_dp), int(threads))
    return H


def features10(H, pts, locality):
    H, pH = _d(H)
    pts, pp = _d(pts)
    N = H.shape[0]
    out = np.empty((N, 10))
    lib().orc_features10(pH, pp, C.c_double(locality), C.c_int64(N), out.ctypes.data_as(c_dp))
    return out


def features6(H):
    H, pH = _d(H)
    K = H.shape[0]
    out = np.empty((K, 6))
    lib().orc_features6(pH, C.c_int64(K), out.ctypes.data_as(c_dp))
    return out


def meanshift(data, bw, metric=0, rng_state=1):
    """Returns (centres CxD, assign N, rng_state', (trajectories, window_iterations))."""
    data, pd = _d(data)
    N, D = data.shape
    centres = np.empty((max(N, 1), D))
    assign = np.empty(N, dtype=np.int32)
    st = C.c_uint32(rng_state)
    stats = np.zeros(2, dtype=np.int64)
    Cn = lib().orc_meanshift(pd, N, D, C.c_double(bw), int(metric), C.byref(st), centres.ctypes.data_as(c_dp), N,
                             assign.ctypes.data_as(c_ip), stats.ctypes.data_as(c_lp))
    return centres[:Cn].copy(), assign, int(st.value), (int(stats[0]), int(stats[1]))


def normalize_points(pts):
    pts, pp = _d(pts)
    n = pts.shape[0]
    out = np.empty((n, 2))
    T = np.empty((3, 3))
    lib().orc_normalize_points(pp, n, out.ctypes.data_as(c_dp), T.ctypes.data_as(c_dp))
    return out, T


def homography_3pt(pts1, pts2, F, refine=False):
    """GetHomography3PT (MultiH.cpp:995-1055); refine = do_numerical_refinement (the reference's LM polish, 3PTcb.h)."""
    pts1, p1 = _d(pts1)
    pts2, p2 = _d(pts2)
    F, pF = _d(F)
    H = np.empty((3, 3))
    lib().orc_homography_3pt_ex(p1, p2, pts1.shape[0], pF, int(refine), H.ctypes.data_as(c_dp))
    return H


def cluster_3pt(pts, offsets, members, F, refine=False):
    pts, pp = _d(pts)
    offsets, po = _i32(offsets)
    members, pm = _i32(members)
    F, pF = _d(F)
    Cn = offsets.shape[0] - 1
    H = np.zeros((Cn, 9))
    keep = np.zeros(Cn, dtype=np.int32)
    lib().orc_cluster_3pt_ex(pp, po, pm, Cn, pF, int(refine), H.ctypes.data_as(c_dp), keep.ctypes.data_as(c_ip))
    return H, keep.astype(bool)


def compatibility_check(pts, labels, H, F, thr=2.2, min_inliers=20, rng_state=1):
    """MultiH::HomographyCompatibilityCheck (MultiH.cpp:100-222), serial draw order.  Returns (labels, H, medians, removed,
    rng_state); medians are NaN for clusters too small to be tested."""
    pts, pp = _d(pts)
    lab = np.ascontiguousarray(labels, dtype=np.int32).copy()
    Hc = np.ascontiguousarray(np.asarray(H, dtype=np.float64).reshape(-1, 9)).copy()
    K = Hc.shape[0]
    F, pF = _d(F)
    med = np.full(max(K, 1), np.nan)
    rem = np.zeros(max(K, 1), dtype=np.int32)
    st = C.c_uint32(rng_state)
    f = lib().orc_compatibility_check
    f.restype = C.c_int
    Kn = f(pp, C.c_int64(len(pts)), lab.ctypes.data_as(c_ip), Hc.ctypes.data_as(c_dp), K, pF, C.c_double(thr),
           int(min_inliers), C.byref(st), med.ctypes.data_as(c_dp), rem.ctypes.data_as(c_ip))
    return lab, Hc[:Kn], med[:K], rem[:K].astype(bool), int(st.value)


def mode_to_homography(mode6, F, refine=False):
    m, pm = _d(mode6)
    F, pF = _d(F)
    H = np.empty((3, 3))
    lib().orc_mode_to_homography_ex(pm, pF, int(refine), H.ctypes.data_as(c_dp))
    return H


def residuals(pts, H, threads=1):
    pts, pp = _d(pts)
    H, pH = _d(H)
    N, K = pts.shape[0], H.shape[0]
    out = np.empty((N, K))
    lib().orc_residuals(pp, C.c_int64(N), pH, K, out.ctypes.data_as(c_dp), int(threads))
    return out


def data_cost_dense(pts, H, lam=0.5, thr=2.2, threads=1):
    pts, pp = _d(pts)
    H, pH = _d(H)
    N, K = pts.shape[0], H.shape[0]
    out = np.empty((N, K + 1), dtype=np.int32)
    lib().orc_data_cost_dense(pp, C.c_int64(N), pH, K, C.c_double(lam), C.c_double(thr), out.ctypes.data_as(c_ip),
                              int(threads))
    return out


def data_cost_sweep(pts, H, lam=0.5, thr=2.2, threads=1, want_argmin=True, want_counts=True):
    pts, pp = _d(pts)
    H, pH = _d(H)
    N, K = pts.shape[0], H.shape[0]
    arg = np.empty(N, dtype=np.int32) if want_argmin else None
    cnt = np.zeros(K, dtype=np.int64) if want_counts else None
    tot = lib().orc_data_cost_sweep(pp, C.c_int64(N), pH, K, C.c_double(lam), C.c_double(thr),
                                    arg.ctypes.data_as(c_ip) if want_argmin else None,
                                    cnt.ctypes.data_as(c_lp) if want_counts else None, int(threads))
    return int(tot), arg, cnt


def smooth_cost(l1, l2, lam=0.5):
    return int(lib().orc_smooth_cost(int(l1), int(l2), C.c_double(lam)))


def inlier_stats(pts, H, thr=2.2, straightness=0.005):
    pts, pp = _d(pts)
    H, pH = _d(H)
    N, K = pts.shape[0], H.shape[0]
    count = np.zeros(K, dtype=np.int64)
    scat = np.zeros((K, 6))
    lmin = np.zeros(K)
    keep = np.zeros(K, dtype=np.int32)
    lib().orc_inlier_stats(pp, C.c_int64(N), pH, K, C.c_double(thr), C.c_double(straightness),
                           count.ctypes.data_as(c_lp), scat.ctypes.data_as(c_dp), lmin.ctypes.data_as(c_dp),
                           keep.ctypes.data_as(c_ip))
    return count, scat, lmin, keep.astype(bool)


def inliers_of_homography(pts, h, thr, idx, labels):
    pts, pp = _d(pts)
    h, ph = _d(h)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    lib().orc_inliers_of_homography(pp, C.c_int64(pts.shape[0]), ph, C.c_double(thr), int(idx),
                                    labels.ctypes.data_as(c_ip))
    return labels


def refit_haf(pts, aff, labels, K, F, e2=None, H_init=None):
    pts, pp = _d(pts)
    aff, pa = _d(aff)
    labels, pl = _i32(labels)
    F, pF = _d(F)
    e2 = epipole2(F) if e2 is None else np.asarray(e2, dtype=np.float64)
    e2, pe = _d(e2)
    H = np.zeros((K, 9)) if H_init is None else np.ascontiguousarray(H_init, dtype=np.float64).reshape(K, 9).copy()
    M10 = np.zeros((K, 10))
    cnt = np.zeros(K, dtype=np.int64)
    lib().orc_refit_haf(pp, pa, pl, C.c_int64(pts.shape[0]), int(K), pF, pe, H.ctypes.data_as(c_dp),
                        M10.ctypes.data_as(c_dp), cnt.ctypes.data_as(c_lp))
    return H, M10, cnt


def radius_neighbours(pts, radius, max_neighbours=31):
    pts, pp = _d(pts)
    N = pts.shape[0]
    offsets = np.zeros(N + 1, dtype=np.int64)
    total = lib().orc_radius_neighbours(pp, N, C.c_double(radius), int(max_neighbours), offsets.ctypes.data_as(c_lp), None)
    adj = np.empty(max(total, 1), dtype=np.int32)
    lib().orc_radius_neighbours(pp, N, C.c_double(radius), int(max_neighbours), offsets.ctypes.data_as(c_lp),
                                adj.ctypes.data_as(c_ip))
    return offsets, adj[:total]


def gco_ref_expansion(data_cost, potts_weight, offsets, adj, init_labels=None, max_iter=1000):
    """The reference's alpha-expansion on a dense site-major cost matrix. Returns (energy, labels)."""
    r = ref_lib()
    if r is None:
        raise RuntimeError("oracle/_ref/libgco_ref.so not built (reference tree absent and no prebuilt copy)")
    dc, pdc = _i32(data_cost)
    N, L = dc.shape
    offsets, po = _i64(offsets)
    adj, pa = _i32(adj)
    out = np.empty(N, dtype=np.int32)
    status = C.c_int(0)
    if init_labels is not None:
        init_labels, pi = _i32(init_labels)
    else:
        pi = None
    e = r.gco_ref_expansion(N, L, pdc, int(potts_weight), po, pa, pi, int(max_iter), out.ctypes.data_as(c_ip),
                            C.byref(status))
    if status.value:
        raise RuntimeError("reference GCO raised GCException")
    return int(e), out


def prefilter(pts, aff, F):
    """MH.cpp:786-838 with F given. Returns (pts_kept Mx4, aff_kept Mx4, keep mask N)."""
    pts, pp = _d(pts)
    aff, pa = _d(aff)
    F, pF = _d(F)
    N = pts.shape[0]
    op = np.zeros((N, 4)); oa = np.zeros((N, 4)); keep = np.zeros(N, dtype=np.int32)
    lib().orc_prefilter(pp, pa, pF, C.c_int64(N), op.ctypes.data_as(c_dp), oa.ctypes.data_as(c_dp), keep.ctypes.data_as(c_ip))
    k = keep.astype(bool)
    return op[k], oa[k], k
