"""Synthetic multi-plane stereo scenes and the reference's correspondence text format.

The generator is the shared workload definition of SURVEY.md §8(d) / BASELINE.json `configs[2..4]`:
two pinhole views, P planes, affine correspondences = (x1, y1, x2, y2, a11, a12, a21, a22) where the affine is the
Jacobian of the plane homography at the point, plus epipolar-consistent outliers. Counter-based RNG (Philox),
``seed = 0xB200 + cfg``.

Text I/O follows the reference's `LoadPointsFromFile` / `SavePointsToFile`
(MultiH/MultiH/main.cpp:380-446): one correspondence per line, `x1 y1 x2 y2 a11 a12 a21 a22 [label]`.
"""
from __future__ import annotations

import dataclasses

import numpy as np

IMG_W, IMG_H = 1000.0, 700.0
K_CAM = np.array([[800.0, 0.0, 500.0], [0.0, 800.0, 350.0], [0.0, 0.0, 1.0]])


@dataclasses.dataclass
class Scene:
    pts: np.ndarray  # N x 4 float64: x1 y1 x2 y2
    aff: np.ndarray  # N x 4 float64: a11 a12 a21 a22 (row-major 2x2, dx2/dx1)
    F: np.ndarray  # 3 x 3, x2^T F x1 = 0, F[2,2] == 1
    planes: np.ndarray  # P x 9 generating homographies (h33 == 1)
    gt: np.ndarray  # N int32: plane index, -1 = outlier


def _skew(t):
    return np.array([[0.0, -t[2], t[1]], [t[2], 0.0, -t[0]], [-t[1], t[0], 0.0]])


def camera_pair():
    a = 0.15
    R = np.array([[np.cos(a), 0.0, np.sin(a)], [0.0, 1.0, 0.0], [-np.sin(a), 0.0, np.cos(a)]])
    t = np.array([1.0, 0.1, 0.2])
    Ki = np.linalg.inv(K_CAM)
    F = Ki.T @ _skew(t) @ R @ Ki
    F = F / F[2, 2]
    return R, t, F


def _plane_h(R, t, n, d):
    """H = K (R + t n^T / d) K^-1 for planes n^T X = d (camera-1 frame); batched over leading axis."""
    Ki = np.linalg.inv(K_CAM)
    M = R[None] + t[None, :, None] * n[:, None, :] / d[:, None, None]
    H = K_CAM[None] @ M @ Ki[None]
    return H / H[:, 2:3, 2:3]


def _apply_h(H, x, y):
    """H: (..., 9) per-point or broadcast; returns x2, y2 and the analytic 2x2 Jacobian."""
    s = H[..., 6] * x + H[..., 7] * y + H[..., 8]
    x2 = (H[..., 0] * x + H[..., 1] * y + H[..., 2]) / s
    y2 = (H[..., 3] * x + H[..., 4] * y + H[..., 5]) / s
    a11 = (H[..., 0] - H[..., 6] * x2) / s
    a12 = (H[..., 1] - H[..., 7] * x2) / s
    a21 = (H[..., 3] - H[..., 6] * y2) / s
    a22 = (H[..., 4] - H[..., 7] * y2) / s
    return x2, y2, np.stack([a11, a12, a21, a22], axis=-1)


def make_scene(n_points: int, n_planes: int, *, outlier_ratio: float = 0.5, noise_px: float = 0.5,
               noise_aff: float = 0.02, seed: int = 0xB200, coherent: bool = True) -> Scene:
    rng = np.random.Generator(np.random.Philox(seed))
    R, t, F = camera_pair()

    n = rng.normal(0.0, 0.3, size=(n_planes, 3)) + np.array([0.0, 0.0, 1.0])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    d = rng.uniform(4.0, 12.0, size=n_planes)
    Hp = _plane_h(R, t, n, d).reshape(n_planes, 9)

    n_out = int(round(n_points * outlier_ratio))
    n_in = n_points - n_out

    # inliers: uniform in image 1, plane = Voronoi cell of P random seeds (spatially coherent) or uniform random
    x1 = rng.uniform(0.0, IMG_W, size=n_in)
    y1 = rng.uniform(0.0, IMG_H, size=n_in)
    if coherent and n_planes > 1:
        seeds = np.stack([rng.uniform(0.0, IMG_W, n_planes), rng.uniform(0.0, IMG_H, n_planes)], axis=1)
        from scipy.spatial import cKDTree

        _, pid = cKDTree(seeds).query(np.stack([x1, y1], axis=1), workers=-1)
        pid = pid.astype(np.int32)
    else:
        pid = rng.integers(0, n_planes, size=n_in).astype(np.int32)
    x2, y2, A = _apply_h(Hp[pid], x1, y1)
    pts_in = np.stack([x1, y1, x2, y2], axis=1)

    # outliers: random 3-D points (depth U[2,20]) seen in both views, affine from a random local plane
    xo = rng.uniform(0.0, IMG_W, size=n_out)
    yo = rng.uniform(0.0, IMG_H, size=n_out)
    depth = rng.uniform(2.0, 20.0, size=n_out)
    Ki = np.linalg.inv(K_CAM)
    ray = (Ki @ np.stack([xo, yo, np.ones(n_out)], axis=0)).T  # z == 1
    X = ray * depth[:, None]
    no = rng.normal(0.0, 0.5, size=(n_out, 3)) + np.array([0.0, 0.0, 1.0])
    no /= np.linalg.norm(no, axis=1, keepdims=True)
    do = np.einsum("ij,ij->i", no, X)
    bad = np.abs(do) < 1e-3
    do[bad] = 1e-3
    Ho = _plane_h(R, t, no, do).reshape(n_out, 9)
    xo2, yo2, Ao = _apply_h(Ho, xo, yo)
    pts_out = np.stack([xo, yo, xo2, yo2], axis=1)

    pts = np.concatenate([pts_in, pts_out], axis=0)
    aff = np.concatenate([A, Ao], axis=0)
    gt = np.concatenate([pid, -np.ones(n_out, dtype=np.int32)])
    pts = pts + rng.normal(0.0, noise_px, size=pts.shape) if noise_px > 0 else pts
    aff = aff + rng.normal(0.0, noise_aff, size=aff.shape) if noise_aff > 0 else aff

    perm = rng.permutation(n_points)
    return Scene(np.ascontiguousarray(pts[perm]), np.ascontiguousarray(aff[perm]), F, Hp, gt[perm].astype(np.int32))


def load_points(path: str):
    """`x1 y1 x2 y2 a11 a12 a21 a22 [label]` per line (main.cpp:390, 420-421, 438-440)."""
    a = np.loadtxt(path, dtype=np.float64, ndmin=2)
    pts, aff = np.ascontiguousarray(a[:, 0:4]), np.ascontiguousarray(a[:, 4:8])
    labels = a[:, 8].astype(np.int32) if a.shape[1] > 8 else None
    return pts, aff, labels


def save_points(path: str, pts, aff, labels=None) -> None:
    with open(path, "w") as f:
        for i in range(len(pts)):
            row = [f"{v:g}" for v in (*pts[i], *aff[i])]
            if labels is not None:
                row.append(str(int(labels[i])))
            f.write(" ".join(row) + "\n")
