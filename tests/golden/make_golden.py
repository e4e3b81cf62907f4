"""Generates the committed golden vectors under tests/golden/.

This is an INDEPENDENT, line-by-line transliteration of the reference's hot-path functions (MultiH/MultiH/MultiH.cpp;
citations per function) written on top of the real OpenCV numerical routines the reference calls — cv2.eigen,
cv2.invert(DECOMP_SVD), cv2.solvePoly, cv2.findFundamentalMat — instead of the oracle's own Jacobi/SVD code.  OpenCV
here is cv2 4.13 (the reference pins 3.1.0; the routines are the same algorithms).  It runs only in the build container
(it reads /root/reference for the bundled barrsmith correspondences); the .npz files it writes are what travels.

    python tests/golden/make_golden.py
"""
import math
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import multih_b200 as m  # noqa: E402  (scenes only; no GPU code is touched)


def c_round(v):  # C round(): half away from zero
    return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)


def epipole2(F):  # MultiH.cpp:789-793
    _, _, evec = cv2.eigen(F @ F.T)
    e = evec[-1]
    return e / e[2]


def haf_rows(x1, y1, x2, y2, a11, a12, a21, a22, F, e):  # MultiH.cpp:859-887
    f = F.ravel()
    return np.array([
        [a11 * x1 + x2 - e[0], a11 * y1, a11, -f[3]],
        [a12 * x1, a12 * y1 + x2 - e[0], a12, -f[4]],
        [a21 * x1 + y2 - e[1], a21 * y1, a21, f[0]],
        [a22 * x1, a22 * y1 + y2 - e[1], a22, f[1]],
        [e[0] * x1 - x2 * x1, e[0] * y1 - x2 * y1, e[0] - x2, x1 * f[3] + y1 * f[4] + f[5]],
        [e[1] * x1 - y2 * x1, e[1] * y1 - y2 * y1, e[1] - y2, -(x1 * f[0] + y1 * f[1] + f[2])],
    ])


def assemble(res, F, e):  # MultiH.cpp:899-909
    f = F.ravel()
    H = np.zeros(9)
    H[6:9] = res[:3]
    lam = res[3]
    H[3] = e[1] * H[6] - lam * f[0]; H[4] = e[1] * H[7] - lam * f[1]; H[5] = e[1] * H[8] - lam * f[2]
    H[0] = e[0] * H[6] + lam * f[3]; H[1] = e[0] * H[7] + lam * f[4]; H[2] = e[0] * H[8] + lam * f[5]
    return H


def get_homography_haf(p, a, F, e):  # MultiH.cpp:850-911
    A = haf_rows(*p, *a, F, e)
    _, _, evec = cv2.eigen(A.T @ A)
    H = assemble(evec[3], F, e)
    return H / H[8]


def get_homography_haf_nonminimal(pts, aff, F, e):  # MultiH.cpp:913-990 (linear part)
    A = np.concatenate([haf_rows(*pts[i], *aff[i], F, e) for i in range(len(pts))])
    _, _, evec = cv2.eigen(A.T @ A)
    return assemble(evec[3], F, e)


def normalize_points(pts):  # Homography_Refine3PTCallback.h:161-195
    mass = pts.mean(0)
    q = pts - mass
    avg = np.sqrt((q ** 2).sum(1)).mean()
    ratio = math.sqrt(2) / avg
    T = np.array([[ratio, 0, -mass[0] * ratio], [0, ratio, -mass[1] * ratio], [0, 0, 1.0]])
    return q * ratio, T


def get_homography_3pt(p1, p2, F):  # MultiH.cpp:995-1055 (linear part)
    n1, T1 = normalize_points(p1)
    n2, T2 = normalize_points(p2)
    Fn = np.linalg.inv(T2).T @ F @ np.linalg.inv(T1)
    _, _, evec = cv2.eigen(Fn @ Fn.T)
    e = evec[-1] / evec[-1][2]
    f = Fn.ravel()
    A = np.zeros((2 * len(p1), 3)); b = np.zeros((2 * len(p1), 1))
    for i in range(len(p1)):
        x1, y1 = n1[i]; x2, y2 = n2[i]
        A[2 * i] = [e[0] * x1 - x2 * x1, e[0] * y1 - x2 * y1, e[0] - x2]
        A[2 * i + 1] = [e[1] * x1 - y2 * x1, e[1] * y1 - y2 * y1, e[1] - y2]
        b[2 * i] = -(x1 * f[3] + y1 * f[4] + f[5])
        b[2 * i + 1] = (x1 * f[0] + y1 * f[1] + f[2])
    _, Ainv = cv2.invert(A, flags=cv2.DECOMP_SVD)
    res = (Ainv @ b).ravel()
    Hn = assemble(np.array([res[0], res[1], res[2], 1.0]), Fn, e).reshape(3, 3)
    return (np.linalg.inv(T2) @ Hn @ T1).ravel()


def features(H):
    h = H
    s1 = h[8]; x1 = h[2] / s1; y1 = h[5] / s1
    s2 = h[6] + h[8]; x2 = (h[0] + h[2]) / s2; y2 = (h[3] + h[5]) / s2
    s3 = h[7] + h[8]; x3 = (h[1] + h[2]) / s3; y3 = (h[4] + h[5]) / s3
    return x1, y1, x2, y2, x3, y3


def data_energy(p, h, lam_w=0.5, thr=2.2):  # MultiH.cpp:473-504, MultiH.h:41-44
    lam = 100.0 / lam_w
    T = thr * thr * 81.0 / 16.0
    if h is None:
        return c_round(lam * T)
    s1 = h[6] * p[0] + h[7] * p[1] + h[8]
    x1 = (h[0] * p[0] + h[1] * p[1] + h[2]) / s1
    y1 = (h[3] * p[0] + h[4] * p[1] + h[5]) / s1
    d = (x1 - p[2]) ** 2 + (y1 - p[3]) ** 2
    if d < T:
        return c_round(lam * (1.0 - d / T))
    return 2 * c_round(lam * T)


def inlier_stats(pts, h, thr=2.2, straight=0.005):  # MultiH.cpp:430-463
    s = h[6] * pts[:, 0] + h[7] * pts[:, 1] + h[8]
    x = (h[0] * pts[:, 0] + h[1] * pts[:, 1] + h[2]) / s
    y = (h[3] * pts[:, 0] + h[4] * pts[:, 1] + h[5]) / s
    inl = (pts[:, 2] - x) ** 2 + (pts[:, 3] - y) ** 2 < thr * thr
    L = np.c_[pts[inl, 0], pts[inl, 1], np.ones(inl.sum())]
    _, ev, _ = cv2.eigen(L.T @ L)
    lmin = float(ev[2, 0])
    return int(inl.sum()), lmin, not (lmin < straight or inl.sum() < 3)


def small_vectors():
    sc = m.scenes.make_scene(300, 4, seed=0xB200 + 77)
    pts, aff, F = sc.pts, sc.aff, sc.F
    e = epipole2(F)
    haf = np.stack([get_homography_haf(pts[i], aff[i], F, e) for i in range(len(pts))])
    f10 = np.stack([np.array([*(np.array(features(h))[[0, 2, 4, 1, 3, 5]]), *(0.005 * pts[i])]) for i, h in enumerate(haf)])
    f6 = np.stack([np.array(features(h)) for h in haf[:16]])
    cost_H = np.concatenate([sc.planes, haf[:12]])
    cost = np.array([[data_energy(p, None)] + [data_energy(p, h) for h in cost_H] for p in pts], dtype=np.int32)
    rng = np.random.default_rng(5)
    pt3_idx, pt3_H = [], []
    for pl in range(4):
        members = np.where(sc.gt == pl)[0]
        for n in (3, 7, 25):
            idx = rng.choice(members, size=min(n, len(members)), replace=False)
            pt3_idx.append(np.pad(idx, (0, 25 - len(idx)), constant_values=idx[0]) if False else idx)
            pt3_H.append(get_homography_3pt(pts[idx, :2], pts[idx, 2:], F))
    labels = sc.gt.copy()
    refit = np.zeros((4, 9))
    for pl in range(4):
        sel = labels == pl
        refit[pl] = get_homography_haf_nonminimal(pts[sel], aff[sel], F, e)
    modes = np.stack([np.array(features(h)) for h in sc.planes]) + rng.normal(0, 0.3, size=(4, 6))
    p1 = np.array([[0.0, 0], [1, 0], [0, 1]])
    modes_H = np.stack([get_homography_3pt(p1, mo.reshape(3, 2), F) for mo in modes])
    st = [inlier_stats(pts, h) for h in cost_H]
    # ragged index lists: store as object-free padded array + lengths
    maxn = max(len(i) for i in pt3_idx)
    pad = np.full((len(pt3_idx), maxn), -1, dtype=np.int64)
    for k, i in enumerate(pt3_idx):
        pad[k, :len(i)] = i
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), pts=pts, aff=aff, F=F, e2=e[:2], haf_H=haf, feat10=f10,
                        feat6=f6, cost_H=cost_H, cost=cost, pt3_idx_padded=pad, pt3_H=np.stack(pt3_H), labels=labels,
                        refit_H=refit, modes=modes, modes_H=modes_H, inl_count=np.array([s[0] for s in st]),
                        inl_lmin=np.array([s[1] for s in st]), inl_keep=np.array([s[2] for s in st]), gt=sc.gt,
                        planes=sc.planes)
    print("golden_small.npz written")


# ---- pre-path transliteration, only to build the barrsmith hot-path input fixture ------------
def optimal_triangulation(pt1, pt2, F, e1, e2, R1, R2):  # MultiH.cpp:1116-1188
    T1 = np.array([[1, 0, -pt1[0]], [0, 1, -pt1[1]], [0, 0, 1.0]])
    T2 = np.array([[1, 0, -pt2[0]], [0, 1, -pt2[1]], [0, 0, 1.0]])
    F2 = np.linalg.inv(T2.T) @ F @ np.linalg.inv(T1)
    F3 = R2 @ F2 @ R1.T
    f1, f2 = e1[2], e2[2]
    a, b, c, d = F3[1, 1], F3[1, 2], F3[2, 1], F3[2, 2]
    t6 = -a * c * f1 ** 4 * (a * d - b * c)
    t5 = (a * a + f2 * f2 * c * c) ** 2 - (a * d + b * c) * f1 ** 4 * (a * d - b * c)
    t4 = 2 * (a * a + f2 * f2 * c * c) * (2 * a * b + 2 * c * d * f2 * f2) - d * b * f1 ** 4 * (a * d - b * c) - 2 * a * c * f1 * f1 * (a * d - b * c)
    t3 = (2 * a * b + 2 * c * d * f2 * f2) ** 2 + 2 * (a * a + f2 * f2 * c * c) * (b * b + f2 * f2 * d * d) - 2 * f1 * f1 * (a * d - b * c) * (a * d + b * c)
    t2 = 2 * (2 * a * b + 2 * c * d * f2 * f2) * (b * b + f2 * f2 * d * d) - 2 * (f1 * f1 * a * d - f1 * f1 * b * c) * b * d - a * c * (a * d - b * c)
    t1 = (b * b + f2 * f2 * d * d) ** 2 - (a * d + b * c) * (a * d - b * c)
    t0 = -(a * d - b * c) * b * d
    _, roots = cv2.solvePoly(np.array([[t0, t1, t2, t3, t4, t5, t6]]), maxIters=300)
    bestS, bestT = float(2 ** 31 - 1), 0.0
    for r in roots.reshape(-1, 2):
        if abs(r[1]) <= 1e-10:
            t = r[0]
            val = t * t / (1 + f1 * f1 * t * t) + (c * t + d) ** 2 / ((a * t + b) ** 2 + f2 * f2 * (c * t + d) ** 2)
            if val < bestS:
                bestS, bestT = val, t
    valInf = 1 / (f1 * f1) + c * c / (a * a + f2 * f2 * c * c)
    if valInf < bestS:
        return None
    point1 = np.array([0, bestT, 1.0])
    l2 = F3 @ point1
    point2 = np.array([-l2[0] * l2[2], -l2[1] * l2[2], l2[0] ** 2 + l2[1] ** 2])
    point2 /= point2[2]
    return np.linalg.inv(R1 @ T1) @ point1, np.linalg.inv(R2 @ T2) @ point2


def beta_scale(F, A, pt1, pt2):  # MultiH.cpp:1092-1114
    l1 = F.T @ pt2; l2 = F @ pt1
    xn1 = pt1[0] + 1.0; yn1 = -(l1[0] * xn1 + l1[2]) / l1[1]
    dx1 = np.array([xn1, yn1, 1.0]) - pt1
    dx1 /= np.linalg.norm(dx1)
    f = F.ravel()
    return abs(math.sqrt(l2[0] ** 2 + l2[1] ** 2) / ((-f[0] * dx1[1] + f[1] * dx1[0]) * pt2[0] + (-f[3] * dx1[1] + f[4] * dx1[0]) * pt2[1] - f[6] * dx1[1] + f[7] * dx1[0]))


def affine_consistency_distance(F, A, pt1, pt2):  # MultiH.cpp:1057-1090 (distanceError only)
    l1 = F.T @ pt2; l2 = F @ pt1
    l1 = l1 / l1[2]; l2 = l2 / l2[2]
    n1 = np.array([l1[0], l1[1]]); n2 = np.array([l2[0], l2[1]])
    n1 /= np.linalg.norm(n1); n2 /= np.linalg.norm(n2)
    beta = beta_scale(F, A, pt1, pt2)
    r1 = np.linalg.inv(A).T @ n1
    return np.linalg.norm(r1 - beta * n2)


def optimal_affine(A, F, pt1, pt2):  # MultiH.cpp:1190-1223
    l1 = F.T @ pt2; l2 = F @ pt1
    l1 = l1 / l1[2]; l2 = l2 / l2[2]
    n1 = np.array([l1[0], l1[1]]); n2 = np.array([l2[0], l2[1]])
    n1 /= np.linalg.norm(n1); n2 /= np.linalg.norm(n2)
    beta = beta_scale(F, A, pt1, pt2)
    if n1 @ n2 < 0:
        n2 = -n2
    C = np.array([[1, 0, 0, 0, -beta * n2[0], 0], [0, 1, 0, 0, 0, -beta * n2[0]], [0, 0, 1, 0, -beta * n2[1], 0],
                  [0, 0, 0, 1, 0, -beta * n2[1]], [-beta * n2[0], 0, -beta * n2[1], 0, 0, 0],
                  [0, -beta * n2[0], 0, -beta * n2[1], 0, 0]])
    b = np.array([A[0, 0], A[0, 1], A[1, 0], A[1, 1], -n1[0], -n1[1]])
    x = np.linalg.inv(C) @ b
    return x[:4]


def barrsmith_fixture():
    src = "/root/reference/Executable/results/barrsmith/barrsmith_points_with_no_annotation.txt"
    if not os.path.exists(src):
        print("reference tree absent: barrsmith fixture not regenerated")
        return
    pts, aff, _ = m.scenes.load_points(src)
    cv2.setRNGSeed(12345)
    _, mask = cv2.findFundamentalMat(pts[:, :2], pts[:, 2:], cv2.FM_RANSAC, 2.0, 0.99)  # main.cpp:399-409
    keep = mask.ravel().astype(bool)
    pts, aff = pts[keep], aff[keep]
    F, mask = cv2.findFundamentalMat(pts[:, :2], pts[:, 2:], cv2.FM_RANSAC, 2.6, 0.99)  # MultiH.cpp:775
    mask = mask.ravel().astype(bool)
    e2 = epipole2(F)
    _, _, ev = cv2.eigen(F.T @ F)
    e1 = ev[-1] / ev[-1][2]
    R1 = np.array([[e1[0], e1[1], 0], [-e1[1], e1[0], 0], [0, 0, 1.0]])      # MultiH.cpp:801
    R2 = np.array([[-e2[0], -e2[1], 0], [e2[1], -e2[0], 0], [0, 0, 1.0]])   # MultiH.cpp:802
    out_p, out_a = [], []
    for i in np.where(mask)[0]:  # MultiH.cpp:807-838
        r = optimal_triangulation(np.array([*pts[i, :2], 1.0]), np.array([*pts[i, 2:], 1.0]), F, e1, e2, R1, R2)
        if r is None:
            continue
        c, d = r
        A = aff[i].reshape(2, 2)
        if affine_consistency_distance(F, A, c, d) > 1.0:
            continue
        out_a.append(optimal_affine(A, F, c, d))
        out_p.append([c[0], c[1], d[0], d[1]])
    out_p, out_a = np.array(out_p), np.array(out_a)
    np.savez_compressed(os.path.join(HERE, "barrsmith_hotpath_input.npz"), pts=out_p, aff=out_a, F=F,
                        n_file=2903, n_after_load_ransac=int(keep.sum()), n_after_f_ransac=int(mask.sum()))
    print("barrsmith fixture:", 2903, "->", int(keep.sum()), "->", int(mask.sum()), "->", len(out_p), "kept")


def prefilter_vectors():
    """Golden vectors of the pre-filter (MultiH.cpp:786-838) on ALL bundled barrsmith rows + a noisy synthetic scene,
    through the cv2.solvePoly-based transliteration above."""
    src = "/root/reference/Executable/results/barrsmith/barrsmith_points_with_no_annotation.txt"
    if not os.path.exists(src):
        print("reference tree absent: prefilter golden not regenerated")
        return
    pts, aff, _ = m.scenes.load_points(src)
    F = np.load(os.path.join(HERE, "barrsmith_hotpath_input.npz"))["F"]
    sc = m.scenes.make_scene(1500, 5, seed=0xB200 + 55, noise_px=1.0, noise_aff=0.05)
    out = {}
    for name, (P_, A_, F_) in {"barr": (pts, aff, F), "syn": (sc.pts, sc.aff, sc.F)}.items():
        e2 = epipole2(F_)
        _, _, ev = cv2.eigen(F_.T @ F_)
        e1 = ev[-1] / ev[-1][2]
        R1 = np.array([[e1[0], e1[1], 0], [-e1[1], e1[0], 0], [0, 0, 1.0]])
        R2 = np.array([[-e2[0], -e2[1], 0], [e2[1], -e2[0], 0], [0, 0, 1.0]])
        keep = np.zeros(len(P_), dtype=bool); op, oa = [], []
        for i in range(len(P_)):
            r = optimal_triangulation(np.array([*P_[i, :2], 1.0]), np.array([*P_[i, 2:], 1.0]), F_, e1, e2, R1, R2)
            if r is None:
                continue
            c, d = r
            A = A_[i].reshape(2, 2)
            if affine_consistency_distance(F_, A, c, d) > 1.0:
                continue
            oa.append(optimal_affine(A, F_, c, d)); op.append([c[0], c[1], d[0], d[1]]); keep[i] = True
        out.update({f"{name}_pts": P_, f"{name}_aff": A_, f"{name}_F": F_, f"{name}_keep": keep,
                    f"{name}_out_pts": np.array(op), f"{name}_out_aff": np.array(oa)})
        print(name, len(P_), "->", int(keep.sum()))
    np.savez_compressed(os.path.join(HERE, "golden_prefilter.npz"), **out)


def compatibility_check(pts, labels, K, F, thr=2.2, min_inliers=20, seed=1):  # MultiH.cpp:100-222, serial draw order
    """Transliteration with Python lists standing in for the std::vectors (erase / resize / write-back at the end) and the MSVC
    rand() the reference is built against; the 3-point fits go through the cv2-based get_homography_3pt above."""
    hold = [seed]

    def rand():
        hold[0] = (hold[0] * 214013 + 2531011) & 0xFFFFFFFF
        return (hold[0] >> 16) & 0x7FFF

    medians, removed = np.full(K, np.nan), np.zeros(K, dtype=bool)
    for c in range(K):
        v = [tuple(p) for p in pts[labels == c]]                          # :106-114
        N = len(v)
        trials = max(501, min(501, N * (N - 1) * (N - 2) // 6))           # :130
        if N >= max(min_inliers, 4):                                      # :138
            dist = [0.0] * N                                              # :140 (size N, only N - 3 entries rewritten per trial)
            distances = []
            for _ in range(trials):
                mss = []
                for _j in range(3):
                    idx = int((len(v) - 1) * (rand() / 32767.0))          # :145
                    mss.append(v.pop(idx))                                # :147-153
                m = np.array(mss)
                h = get_homography_3pt(m[:, :2], m[:, 2:], F)             # :157 (no refinement)
                for j, (ox1, oy1, ox2, oy2) in enumerate(v):              # :162-175
                    sden = h[6] * ox1 + h[7] * oy1 + h[8]
                    x1 = (h[0] * ox1 + h[1] * oy1 + h[2]) / sden
                    y1 = (h[3] * ox1 + h[4] * oy1 + h[5]) / sden
                    dist[j] = (ox2 - x1) ** 2 + (oy2 - y1) ** 2
                dist.sort()                                               # :177 — all N entries
                n = len(v)
                distances.append(dist[n // 2] if n % 2 else 0.5 * (dist[n // 2] + dist[n // 2 + 1]))   # :178
                v = v + [None] * 3                                        # :180-181 resize(N)
                for j in range(3):
                    v[N - j - 1] = mss[j]                                 # :183-194
            distances.sort()
            medians[c] = distances[trials // 2] if trials % 2 else 0.5 * (distances[trials // 2] + distances[trials // 2 + 1])
            removed[c] = medians[c] > thr * thr * 81.0 / 16.0             # :202
        elif N < min_inliers:
            removed[c] = True                                             # :207-208
    return medians, removed, hold[0]


def meanshift_cluster(data, bw, seed=1, max_window_iters=200):  # MeanShiftClustering.h:22-157 (T = double)
    """Transliteration in numpy with the MSVC rand(); `max_window_iters` is the one deliberate deviation shared by oracle and
    kernel (the reference's while(1) at :62 does not return when the mean cycles)."""
    hold = [seed]

    def rand():
        hold[0] = (hold[0] * 214013 + 2531011) & 0xFFFFFFFF
        return (hold[0] >> 16) & 0x7FFF

    num = len(data)
    band_sq, stop = bw * bw, 1e-3 * bw                                   # :31, :48
    init = list(range(num))
    visited = np.zeros(num, dtype=bool)
    cents, votes = [], []
    traj = iters = 0
    while init:                                                          # :52
        rnd = rand() / 32767.0
        st = init[int(math.floor(rnd * (len(init) - 1) + 0.5))]         # :54-56 (round)
        mean = data[st].copy()
        this_votes = np.zeros(num, dtype=np.int64)
        traj += 1
        w = 0
        while True:                                                      # :62
            iters += 1; w += 1
            old = mean.copy()
            sq = np.sqrt((old[None, :] - data) ** 2).sum(1)              # :76-83: SUM_j sqrt(d_j^2) = L1
            inside = sq < band_sq                                        # :85
            this_votes[inside] += 1
            visited[inside] = True
            acc = np.zeros(data.shape[1])
            for i in np.where(inside)[0]:                                # :89 (same summation order as the reference loop)
                acc = acc + data[i]
            mean = acc / inside.sum()                                    # :96
            if np.linalg.norm(mean - old) < stop or w >= max_window_iters:   # :98
                merge = -1
                for cn, c in enumerate(cents):                           # :101-109
                    if np.linalg.norm(mean - c) < bw / 2:
                        merge = cn
                        break
                if merge > -1:
                    cents[merge] = 0.5 * (cents[merge] + mean)           # :113
                    votes[merge] = votes[merge] + this_votes
                else:
                    cents.append(mean)
                    votes.append(this_votes)
                break
        init = [i for i in range(num) if not visited[i]]                 # :124-130
    best, idx = np.zeros(num, dtype=np.int64), np.full(num, -1)
    for r, v in enumerate(votes):                                        # :136-146: most votes, first wins ties
        upd = best < v
        best[upd] = v[upd]
        idx[upd] = r
    return np.array(cents), idx, hold[0], (traj, iters)


def meanshift_vectors():
    """golden_meanshift.npz: the 10-D features of a small scene (EstablishStablePointSets) and a 6-D set (MergingStep)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import multih_b200 as m

    sc = m.scenes.make_scene(260, 3, seed=23)
    F = sc.F
    e = epipole2(F)
    H = np.array([get_homography_haf(sc.pts[i], sc.aff[i], F, e) for i in range(len(sc.pts))])
    f10 = []
    for h, p in zip(H, sc.pts):                                          # MultiH.cpp:612-646
        x1, y1 = h[2] / h[8], h[5] / h[8]
        x2, y2 = (h[0] + h[2]) / (h[6] + h[8]), (h[3] + h[5]) / (h[6] + h[8])
        x3, y3 = (h[1] + h[2]) / (h[7] + h[8]), (h[4] + h[5]) / (h[7] + h[8])
        f10.append([x1, x2, x3, y1, y2, y3] + list(0.005 * p))
    f10 = np.array(f10)
    f10 = f10[np.isfinite(f10).all(1)]
    rng = np.random.default_rng(5)
    f6 = np.concatenate([rng.normal(0, 0.4, (25, 6)) + k * 3.0 for k in range(4)])
    out = {}
    for tag, X, seed in (("f10", f10, 1), ("f6", f6, 777)):
        c, idx, st, stats = meanshift_cluster(X, 2.2, seed)
        out.update({f"{tag}_data": X, f"{tag}_seed": seed, f"{tag}_centres": c, f"{tag}_assign": idx, f"{tag}_rng": st,
                    f"{tag}_stats": np.array(stats)})
        print("meanshift", tag, X.shape, "->", len(c), "centres", stats)
    np.savez_compressed(os.path.join(HERE, "golden_meanshift.npz"), **out)


def compatibility_vectors():
    """golden_compat.npz: a small labelled scene (three planes, one cluster of outliers, one 6-member cluster)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import multih_b200 as m

    sc = m.scenes.make_scene(420, 3, seed=17)
    lab = sc.gt.astype(np.int32).copy()
    out_idx = np.where(lab < 0)[0]
    lab[out_idx[:25]] = 3
    lab[out_idx[25:31]] = 4
    out = dict(pts=sc.pts, labels=lab, F=sc.F)
    for tag, min_inl, seed in (("a", 20, 1), ("b", 4, 4242)):
        med, rem, rng = compatibility_check(sc.pts, lab, 5, sc.F, 2.2, min_inl, seed)
        out.update({f"{tag}_min_inliers": min_inl, f"{tag}_seed": seed, f"{tag}_medians": med, f"{tag}_removed": rem, f"{tag}_rng": rng})
        print("compat", tag, med, rem)
    np.savez_compressed(os.path.join(HERE, "golden_compat.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "compat":
        compatibility_vectors()
    elif len(sys.argv) > 1 and sys.argv[1] == "meanshift":
        meanshift_vectors()
    else:
        small_vectors()
        barrsmith_fixture()
        prefilter_vectors()
        compatibility_vectors()
        meanshift_vectors()
