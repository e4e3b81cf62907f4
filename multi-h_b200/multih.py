"""Python mirror of the reference's `class MultiH` (MultiH/MultiH/MultiH.h:20-149): same constructor arguments, same
method names and return conventions, forwarding to the C ABI (mh_process).  Differences, all deliberate:
  * F is an input (`Process(src, dst, affines, F)`): GetFundamentalMatrixAndRefineData (MultiH.cpp:770) is upstream of
    the hot path;
  * GetDestinationPoints returns the destination points (the reference returns src, MultiH.h:64 — a bug);
  * GetHomography stays 1-based like the reference (MultiH.h:69);
  * like the reference's Process(), the per-correspondence refinement filter (MultiH.cpp:807-838) and the final
    HomographyCompatibilityCheck (:76-86) run; the getters only know the correspondences that survive the filter.
"""
from __future__ import annotations

import numpy as np

from . import capi


class MultiH:
    def __init__(self, thr_fund_mat=3.0, thr_hom=2.5, locality=0.002, lambda_=0.5, minimum_inlier_number=0, device=0):
        # defaults: MultiH.h:7-10, 49-53
        self._params = capi.default_params(thr_fundamental=thr_fund_mat, thr_homography=thr_hom, locality=locality,
                                           lambda_=lambda_, min_inliers=minimum_inlier_number, prefilter=1, compatibility_check=1)
        self._device = device
        self._ctx = None
        self._labels = np.zeros(0, dtype=np.int32)
        self._H = np.zeros((0, 9))
        self._src = self._dst = self._aff = None

    def Process(self, src_points, dst_points, affines, F) -> bool:
        src = np.asarray(src_points, dtype=np.float64).reshape(-1, 2)
        dst = np.asarray(dst_points, dtype=np.float64).reshape(-1, 2)
        aff = np.asarray(affines, dtype=np.float64).reshape(-1, 4)
        if len(src) < 8 or len(dst) != len(src) or len(aff) != len(src):  # MultiH.cpp:44-50
            print("Error: Features are not set!")
            return False
        if self._ctx is None:
            self._ctx = capi.Context(self._params, self._device)
        labels, self._H, _ = self._ctx.process(np.concatenate([src, dst], axis=1), aff, F)
        kept = labels > -2   # -2 = dropped by the refinement filter: the reference forgets those correspondences
        self._src, self._dst, self._aff, self._labels = src[kept], dst[kept], aff[kept], labels[kept]
        return True

    def GetLabel(self, idx): return int(self._labels[idx])
    def GetLabels(self): return self._labels.copy()
    def GetSourcePoints(self): return self._src
    def GetDestinationPoints(self): return self._dst
    def GetAffinities(self): return self._aff
    def GetPointNumber(self): return int(len(self._labels))
    def GetClusterNumber(self): return int(len(self._H))
    def GetIterationNumber(self): return self._ctx.iterations if self._ctx else 0
    def GetHomography(self, idx): return self._H[idx - 1].reshape(3, 3)  # 1-based, MultiH.h:69
    def GetEnergy(self): return self._ctx.energy if self._ctx else 0.0
    def GetHomographyThreshold(self): return float(self._params.thr_homography)
