"""ctypes binding of libmultih_b200.so (include/multih_b200.h).

Device buffers are torch CUDA tensors (torch is plumbing: memory + streams); every compute call goes through the C ABI
and launches our CUDA kernels.  There is no CPU fallback: constructing a Context without the built library or without a
CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmultih_b200.so")
_lib = None

MH_OK, MH_EINVAL, MH_ECUDA, MH_ENCCL, MH_EDEGENERATE, MH_ENOMEM = range(6)
_STATUS = {0: "MH_OK", 1: "MH_EINVAL", 2: "MH_ECUDA", 3: "MH_ENCCL", 4: "MH_EDEGENERATE", 5: "MH_ENOMEM"}

# every symbol include/multih_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "mh_default_params", "mh_create", "mh_destroy", "mh_last_error", "mh_version", "mh_set_stream", "mh_sync",
    "mh_alloc", "mh_free", "mh_host_alloc", "mh_host_free", "mh_memcpy_d2h", "mh_memcpy_h2d", "mh_kernel_launches",
    "mh_set_geometry", "mh_get_geometry", "mh_upload_correspondences", "mh_hypotheses_from_host",
    "mh_hypotheses_to_host", "mh_prefilter", "mh_prefilter_device", "mh_haf_hypotheses", "mh_data_cost_dense", "mh_residuals", "mh_data_cost_fused",
    "mh_inlier_stats", "mh_inliers_of_homography", "mh_features10", "mh_features6", "mh_set_rng_state", "mh_get_rng_state", "mh_meanshift", "mh_refit_haf",
    "mh_refit_haf_accumulate", "mh_refit_haf_solve", "mh_labels_from_best", "mh_pack_inlier_counts", "mh_refit_3pt", "mh_modes_to_hypotheses", "mh_neighbourhood", "mh_alpha_expansion", "mh_compatibility_check", "mh_compat_plan", "mh_compat_decide", "mh_process", "mh_get_energy",
    "mh_get_iterations", "mh_get_stage_ms", "mh_diag_fp32_peak", "mh_diag_mma_tf32_peak", "mh_diag_set_fused_variant", "mh_diag_set_fast_config", "mh_diag_get_fast_config", "mh_diag_set_dense_variant", "mh_diag_get_alternating_ms", "mh_diag_set_neighbourhood_backend",
    "mh_alpha_expansion_sparse", "mh_comm_unique_id", "mh_comm_init", "mh_comm_destroy", "mh_comm_rank", "mh_comm_world", "mh_comm_broadcast",
    "mh_comm_allreduce_sum_f64", "mh_step_sharded", "mh_step_sharded_finish",
]


class MHError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"{_STATUS.get(status, status)}: {msg}")
        self.status = status


class Params(C.Structure):
    """mh_params — mirrors the MultiH constructor (MultiH.h:49-53) + compile-time constants (MultiH.h:7-18)."""
    _fields_ = [
        ("thr_fundamental", C.c_double), ("thr_homography", C.c_double), ("locality", C.c_double),
        ("lambda_", C.c_double), ("min_inliers", C.c_int32), ("straightness", C.c_double),
        ("max_iterations", C.c_int32), ("convergence", C.c_double), ("meanshift_metric", C.c_int32),
        ("rng_seed", C.c_uint32), ("max_gc_cycles", C.c_int32), ("max_neighbours", C.c_int32), ("precise_pipeline", C.c_int32), ("prefilter", C.c_int32),
        ("compatibility_check", C.c_int32), ("lm_refine", C.c_int32),
    ]


def library_path() -> str:
    return _LIB_PATH


def build_library(force: bool = False) -> str:
    """Compile libmultih_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src = os.path.join(_HERE, "csrc")
    if force:
        subprocess.run(["make", "-C", src, "-s", "clean"], check=True)
    subprocess.run(["make", "-C", src, "-s", "-j8"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise MHError(MH_ECUDA, f"{_LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                                    "there is no CPU fallback")
        L = C.CDLL(_LIB_PATH)
        L.mh_last_error.restype = C.c_char_p
        L.mh_version.restype = C.c_char_p
        L.mh_kernel_launches.restype = C.c_int64
        L.mh_get_energy.restype = C.c_double
        L.mh_get_iterations.restype = C.c_int32
        L.mh_get_rng_state.restype = C.c_uint32
        for name in SYMBOLS:
            fn = getattr(L, name)
            if name not in ("mh_last_error", "mh_version", "mh_kernel_launches", "mh_get_energy", "mh_get_iterations",
                            "mh_get_rng_state", "mh_comm_rank", "mh_comm_world",
                            "mh_default_params", "mh_destroy"):
                fn.restype = C.c_int
        _lib = L
    return _lib


def default_params(**kw) -> Params:
    p = Params()
    lib().mh_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, "lambda_" if k == "lambda" else k, v)
    return p


def _vp(t):
    """device pointer of a torch tensor (or None)"""
    return None if t is None else C.c_void_p(t.data_ptr())


def _np(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


# ---- host-only entry points (no GPU needed) --------------------------------------------------------------------------
def neighbourhood(pts, radius, max_neighbours=31, ctx=None, backend=0):
    """mh_neighbourhood: 4-D neighbourhood replacing FlannBasedMatcher::radiusMatch (MultiH.cpp:231-253).  ctx = None: host
    search; with a Context: backend 0 = auto, 1 = host, 2 = device (K5)."""
    pts = _np(pts, np.float64)
    N = pts.shape[0]
    h = None if ctx is None else ctx._h
    lib().mh_diag_set_neighbourhood_backend(None, int(backend))
    try:
        offsets = np.zeros(N + 1, dtype=np.int64)
        total = C.c_int64(0)
        st = lib().mh_neighbourhood(h, _p(pts, C.c_double), N, C.c_double(radius), int(max_neighbours),
                                    _p(offsets, C.c_int64), None, C.byref(total))
        if st:
            raise MHError(st, "mh_neighbourhood")
        adj = np.zeros(max(total.value, 1), dtype=np.int32)
        st = lib().mh_neighbourhood(h, _p(pts, C.c_double), N, C.c_double(radius), int(max_neighbours),
                                    _p(offsets, C.c_int64), _p(adj, C.c_int32), C.byref(total))
        if st:
            raise MHError(st, "mh_neighbourhood")
    finally:
        lib().mh_diag_set_neighbourhood_backend(None, 0)
    return offsets, adj[: total.value]


def alpha_expansion(cost, potts, offsets, adj, init=None, max_cycles=1000):
    """mh_alpha_expansion: the role of GCoptimizationGeneralGraph::expansion in LabelingStep (MultiH.cpp:520-543)."""
    cost = _np(cost, np.int32)
    N, L = cost.shape
    offsets = _np(offsets, np.int64)
    adj = _np(adj, np.int32)
    if adj.size == 0:
        adj = np.zeros(1, dtype=np.int32)
    init_a = None if init is None else _np(init, np.int32)
    labels = np.zeros(N, dtype=np.int32)
    energy = C.c_int64(0)
    st = lib().mh_alpha_expansion(None, _p(cost, C.c_int32), N, L, int(potts), _p(offsets, C.c_int64),
                                  _p(adj, C.c_int32), _p(init_a, C.c_int32), int(max_cycles), _p(labels, C.c_int32),
                                  C.byref(energy))
    if st:
        raise MHError(st, "mh_alpha_expansion")
    return labels, int(energy.value)


def alpha_expansion_sparse(lists, counts, L, cost_label0, cost_default, potts, offsets, adj, init=None, max_cycles=1000):
    """mh_alpha_expansion_sparse: per-site (label << 16 | cost) lists + the two default costs instead of the dense matrix."""
    lists = _np(lists, np.uint32)
    N, kmax = lists.shape
    counts = _np(counts, np.int32)
    offsets = _np(offsets, np.int64)
    adj = _np(adj, np.int32)
    if adj.size == 0:
        adj = np.zeros(1, dtype=np.int32)
    init_a = None if init is None else _np(init, np.int32)
    labels = np.zeros(N, dtype=np.int32)
    energy = C.c_int64(0)
    st = lib().mh_alpha_expansion_sparse(None, _p(lists, C.c_uint32), _p(counts, C.c_int32), int(kmax), N, int(L), int(cost_label0),
                                         int(cost_default), int(potts), _p(offsets, C.c_int64), _p(adj, C.c_int32),
                                         _p(init_a, C.c_int32), int(max_cycles), _p(labels, C.c_int32), C.byref(energy))
    if st:
        raise MHError(st, "mh_alpha_expansion_sparse")
    return labels, int(energy.value)


# ---- GPU context -----------------------------------------------------------------------------------------------------
class Context:
    """One mh_ctx = one CUDA device + one stream (torch's current stream by default)."""

    def __init__(self, params: Params | None = None, device: int = 0, use_torch_stream: bool = True, **kw):
        import torch  # plumbing only

        self.torch = torch
        self._h = C.c_void_p(None)
        self.params = params if params is not None else default_params(**kw)
        st = lib().mh_create(C.byref(self.params), int(device), C.byref(self._h))
        if st:
            raise MHError(st, "mh_create failed: no usable CUDA device (this library has no CPU fallback)")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(device)
        if use_torch_stream:
            self._check(lib().mh_set_stream(self._h, C.c_void_p(torch.cuda.current_stream(device).cuda_stream)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().mh_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st:
            raise MHError(st, (lib().mh_last_error(self._h) or b"").decode())

    def _empty(self, shape, dtype):
        return self.torch.empty(shape, dtype=dtype, device=self.device)

    @property
    def launches(self) -> int:
        return int(lib().mh_kernel_launches(self._h))

    def sync(self):
        self._check(lib().mh_sync(self._h))

    # -- geometry / conversion
    def set_geometry(self, F, pts=None, norm1=None, norm2=None):
        F = _np(F, np.float64).reshape(9)
        n1 = None if norm1 is None else _np(norm1, np.float64)
        n2 = None if norm2 is None else _np(norm2, np.float64)
        p = None if pts is None else (pts if isinstance(pts, np.ndarray) and pts.dtype == np.float64 and pts.flags.c_contiguous else _np(pts, np.float64))
        N = 0 if p is None else p.shape[0]
        self._check(lib().mh_set_geometry(self._h, _p(F, C.c_double), _p(n1, C.c_double), _p(n2, C.c_double),
                                          _p(p, C.c_double), C.c_int64(N)))

    def get_geometry(self):
        F = np.zeros(9); e2 = np.zeros(2); n1 = np.zeros(3); n2 = np.zeros(3)
        self._check(lib().mh_get_geometry(self._h, _p(F, C.c_double), _p(e2, C.c_double), _p(n1, C.c_double),
                                          _p(n2, C.c_double)))
        return F.reshape(3, 3), e2, n1, n2

    def upload(self, pts, aff=None, out=None):
        """host FP64 correspondences (numpy or pinned torch CPU tensors) -> normalised float4 device arrays; `pts` or
        `aff` may be None (upload only the other half)"""
        t = self.torch

        def hp(a):
            if a is None:
                return None
            return C.c_void_p(a.data_ptr()) if isinstance(a, t.Tensor) else _p(_np(a, np.float64), C.c_double)

        N = (pts if pts is not None else aff).shape[0]
        if out is None:
            d_pts = self._empty((N, 4), t.float32) if pts is not None else None
            d_aff = self._empty((N, 4), t.float32) if aff is not None else None
        else:
            d_pts, d_aff = out
        self._check(lib().mh_upload_correspondences(self._h, hp(pts), hp(aff), C.c_int64(N),
                                                    _vp(d_pts if pts is not None else None),
                                                    _vp(d_aff if aff is not None else None)))
        return d_pts, d_aff

    def use_stream(self, torch_stream):
        """bind the context to a torch stream (mh_set_stream)"""
        self._check(lib().mh_set_stream(self._h, C.c_void_p(torch_stream.cuda_stream)))

    def hypotheses_from_host(self, H):
        H = _np(H, np.float64).reshape(-1, 9)
        K = H.shape[0]
        d = self._empty((K, 12), self.torch.float32)
        self._check(lib().mh_hypotheses_from_host(self._h, _p(H, C.c_double), K, _vp(d)))
        return d

    def hypotheses_to_host(self, d_hyp, divide_by_h33=False):
        K = d_hyp.shape[0]
        H = np.zeros((K, 9))
        self._check(lib().mh_hypotheses_to_host(self._h, _vp(d_hyp), K, _p(H, C.c_double), int(divide_by_h33)))
        return H

    # -- K0
    def prefilter(self, pts, aff, F):
        """mh_prefilter: MultiH.cpp:786-838 with F given. Returns (pts_kept, aff_kept, keep mask)."""
        pts = _np(pts, np.float64); aff = _np(aff, np.float64); F = _np(F, np.float64).reshape(9)
        N = pts.shape[0]
        po = np.zeros((N, 4)); ao = np.zeros((N, 4)); keep = np.zeros(N, dtype=np.uint8); M = C.c_int64(0)
        self._check(lib().mh_prefilter(self._h, _p(pts, C.c_double), _p(aff, C.c_double), _p(F, C.c_double), C.c_int64(N),
                                       _p(po, C.c_double), _p(ao, C.c_double), _p(keep, C.c_uint8), C.byref(M)))
        return po[: M.value].copy(), ao[: M.value].copy(), keep.astype(bool)

    # -- K1
    def haf_hypotheses(self, d_pts, d_aff, precision=0, out=None):
        N = d_pts.shape[0]
        d = self._empty((N, 12), self.torch.float32) if out is None else out
        self._check(lib().mh_haf_hypotheses(self._h, _vp(d_pts), _vp(d_aff), C.c_int64(N), _vp(d), int(precision)))
        return d

    # -- K2
    def data_cost_dense(self, d_pts, d_hyp, elem_bytes=4, out=None):
        t = self.torch
        N, K = d_pts.shape[0], (0 if d_hyp is None else d_hyp.shape[0])
        d = out if out is not None else self._empty((N, K + 1), t.int32 if elem_bytes == 4 else t.int16)
        self._check(lib().mh_data_cost_dense(self._h, _vp(d_pts), C.c_int64(N), _vp(d_hyp), K, _vp(d), int(elem_bytes)))
        return d

    def residuals(self, d_pts, d_hyp):
        N, K = d_pts.shape[0], d_hyp.shape[0]
        d = self._empty((N, K), self.torch.float32)
        self._check(lib().mh_residuals(self._h, _vp(d_pts), C.c_int64(N), _vp(d_hyp), K, _vp(d)))
        return d

    def data_cost_fused(self, d_pts, d_hyp, kmax=16, want_list=True, want_best=True, want_inliers=True, out=None):
        """returns dict(list=(N,kmax) int32 packed label<<8|cost, count=(N,), best=(N,) int64 packed cost<<32|label,
        inliers=(K,))"""
        t = self.torch
        N, K = d_pts.shape[0], d_hyp.shape[0]
        o = out or {}
        if "count" not in o and want_list:
            o["count"] = self._empty((N,), t.int32)
        if want_list and "list" not in o:
            o["list"] = self._empty((N, max(kmax, 1)), t.int32)
        if want_best and "best" not in o:
            o["best"] = self._empty((N,), t.int64)
        if want_inliers and "inliers" not in o:
            o["inliers"] = self._empty((K,), t.int32)
        self._check(lib().mh_data_cost_fused(self._h, _vp(d_pts), C.c_int64(N), _vp(d_hyp), K, int(kmax),
                                             _vp(o.get("list") if want_list else None),
                                             _vp(o.get("count") if want_list else None),
                                             _vp(o.get("best") if want_best else None),
                                             _vp(o.get("inliers") if want_inliers else None)))
        return o

    def inlier_stats(self, d_pts, d_hyp):
        K = d_hyp.shape[0]
        sc = np.zeros((K, 6)); lm = np.zeros(K); keep = np.zeros(K, dtype=np.int32)
        self._check(lib().mh_inlier_stats(self._h, _vp(d_pts), C.c_int64(d_pts.shape[0]), _vp(d_hyp), K,
                                          _p(sc, C.c_double), _p(lm, C.c_double), _p(keep, C.c_int32)))
        return sc, lm, keep.astype(bool)

    def inliers_of_homography(self, d_pts, d_hyp_one, idx, d_labels):
        self._check(lib().mh_inliers_of_homography(self._h, _vp(d_pts), C.c_int64(d_pts.shape[0]), _vp(d_hyp_one),
                                                   int(idx), _vp(d_labels)))
        return d_labels

    # -- K3
    @property
    def rng_state(self) -> int:
        return int(lib().mh_get_rng_state(self._h))

    @rng_state.setter
    def rng_state(self, v: int):
        self._check(lib().mh_set_rng_state(self._h, C.c_uint32(v)))

    def features10(self, d_hyp, d_pts):
        N = d_hyp.shape[0]
        d = self._empty((N, 10), self.torch.float64)
        self._check(lib().mh_features10(self._h, _vp(d_hyp), _vp(d_pts), C.c_int64(N), _vp(d)))
        return d

    def features6(self, d_hyp):
        K = d_hyp.shape[0]
        d = self._empty((K, 6), self.torch.float64)
        self._check(lib().mh_features6(self._h, _vp(d_hyp), K, _vp(d)))
        return d

    def meanshift(self, d_feat, bandwidth, max_c=None):
        t = self.torch
        N, D = d_feat.shape
        max_c = N if max_c is None else max_c
        centres = self._empty((max(max_c, 1), D), t.float64)
        assign = self._empty((N,), t.int32)
        Cn = C.c_int32(0)
        stats = np.zeros(2, dtype=np.int64)
        self._check(lib().mh_meanshift(self._h, _vp(d_feat), N, D, C.c_double(bandwidth), _vp(centres), int(max_c),
                                       _vp(assign), C.byref(Cn), _p(stats, C.c_int64)))
        return centres[: Cn.value], assign, (int(stats[0]), int(stats[1]))

    # -- K4
    def refit_haf(self, d_pts, d_aff, d_labels, K, d_hyp=None):
        t = self.torch
        if d_hyp is None:
            d_hyp = t.zeros((K, 12), dtype=t.float32, device=self.device)
        cnt = self._empty((max(K, 1),), t.int32)
        self._check(lib().mh_refit_haf(self._h, _vp(d_pts), _vp(d_aff), _vp(d_labels), C.c_int64(d_pts.shape[0]), int(K),
                                       _vp(d_hyp), _vp(cnt)))
        return d_hyp, cnt[:K]

    def refit_haf_accumulate(self, d_pts, d_aff, d_labels, K, out=None):
        acc = out if out is not None else self._empty((K, 12), self.torch.float64)
        self._check(lib().mh_refit_haf_accumulate(self._h, _vp(d_pts), _vp(d_aff), _vp(d_labels),
                                                  C.c_int64(d_pts.shape[0]), int(K), _vp(acc)))
        return acc

    def labels_from_best(self, d_best, d_labels):
        self._check(lib().mh_labels_from_best(self._h, _vp(d_best), C.c_int64(d_best.shape[0]), _vp(d_labels)))
        return d_labels

    def pack_inlier_counts(self, d_cnt, d_acc, unpack=False):
        self._check(lib().mh_pack_inlier_counts(self._h, _vp(d_cnt), int(d_cnt.shape[0]), _vp(d_acc), int(unpack)))

    def refit_haf_solve(self, d_acc, d_hyp, d_count=None):
        self._check(lib().mh_refit_haf_solve(self._h, _vp(d_acc), int(d_acc.shape[0]), _vp(d_hyp), _vp(d_count)))
        return d_hyp

    # -- multi-GPU (csrc/comm.cu): the library's own NCCL communicator over the ranks that shard the correspondences
    def comm_init(self, rank: int, world: int, exchange=None):
        """Collective.  `exchange(id_bytes_or_None) -> id_bytes` carries rank 0's 128-byte NCCL id to every rank; the default
        uses torch.distributed's object broadcast (any initialised backend — it moves 128 bytes, once)."""
        ident = None
        if rank == 0:
            buf = (C.c_char * 128)()
            st = lib().mh_comm_unique_id(buf)
            if st != 0:
                raise MHError(st, "mh_comm_unique_id: NCCL not loadable (set MH_NCCL_LIB)")
            ident = bytes(buf.raw)
        if exchange is None:
            import torch.distributed as dist

            def exchange(b):
                box = [b]
                dist.broadcast_object_list(box, src=0)
                return box[0]
        ident = exchange(ident)
        self._check(lib().mh_comm_init(self._h, C.c_char_p(ident), int(rank), int(world)))

    def comm_destroy(self):
        self._check(lib().mh_comm_destroy(self._h))

    @property
    def comm_world(self) -> int:
        return int(lib().mh_comm_world(self._h))

    @property
    def comm_rank(self) -> int:
        return int(lib().mh_comm_rank(self._h))

    def comm_broadcast(self, t, root=0):
        self._check(lib().mh_comm_broadcast(self._h, _vp(t), C.c_uint64(t.numel() * t.element_size()), int(root)))
        return t

    def comm_allreduce_sum_f64(self, t):
        assert t.dtype == self.torch.float64
        self._check(lib().mh_comm_allreduce_sum_f64(self._h, _vp(t), C.c_uint64(t.numel())))
        return t

    def step_sharded(self, d_pts, d_aff, d_hyp, d_hyp_pt, d_best, d_labels, d_inliers, d_ref, events=None):
        """one sharded hot pass (include/multih_b200.h: mh_step_sharded); `events` = two torch.cuda.Event recorded around K2"""
        e0 = e1 = None
        if events is not None:
            e0, e1 = (C.c_void_p(e.cuda_event) for e in events)
        self._check(lib().mh_step_sharded(self._h, _vp(d_pts), _vp(d_aff), C.c_int64(d_pts.shape[0]), _vp(d_hyp),
                                          int(d_hyp.shape[0]), _vp(d_hyp_pt), _vp(d_best), _vp(d_labels), _vp(d_inliers),
                                          _vp(d_ref), e0, e1))

    def step_sharded_finish(self):
        self._check(lib().mh_step_sharded_finish(self._h))

    def refit_3pt(self, d_pts, d_assign, Cn):
        t = self.torch
        d_hyp = t.zeros((max(Cn, 1), 12), dtype=t.float32, device=self.device)
        keep = t.zeros((max(Cn, 1),), dtype=t.int32, device=self.device)
        self._check(lib().mh_refit_3pt(self._h, _vp(d_pts), _vp(d_assign), C.c_int64(d_pts.shape[0]), int(Cn),
                                       _vp(d_hyp), _vp(keep)))
        return d_hyp[:Cn], keep[:Cn]

    def modes_to_hypotheses(self, d_modes):
        Cn = d_modes.shape[0]
        d = self._empty((max(Cn, 1), 12), self.torch.float32)
        self._check(lib().mh_modes_to_hypotheses(self._h, _vp(d_modes), Cn, _vp(d)))
        return d[:Cn]

    # -- whole path
    def process(self, pts, aff, F, kmax=4096):
        pts = _np(pts, np.float64); aff = _np(aff, np.float64); F = _np(F, np.float64).reshape(9)
        N = pts.shape[0]
        labels = np.full(N, -1, dtype=np.int32)
        H = np.zeros((kmax, 9))
        K = C.c_int32(0)
        self._check(lib().mh_process(self._h, _p(pts, C.c_double), _p(aff, C.c_double), _p(F, C.c_double), N,
                                     _p(labels, C.c_int32), _p(H, C.c_double), int(kmax), C.byref(K)))
        return labels, H[: K.value].copy(), int(K.value)

    def compatibility_check(self, pts, labels, H):
        """mh_compatibility_check (MultiH.cpp:100-222): returns (labels, H, medians) after the cross-validation filter."""
        pts = _np(pts, np.float64)
        lab = np.ascontiguousarray(labels, dtype=np.int32).copy()
        Hc = np.ascontiguousarray(np.asarray(H, dtype=np.float64).reshape(-1, 9)).copy()
        K = C.c_int32(Hc.shape[0])
        med = np.full(max(Hc.shape[0], 1), np.nan)
        self._check(lib().mh_compatibility_check(self._h, _p(pts, C.c_double), pts.shape[0], _p(lab, C.c_int32),
                                                 _p(Hc, C.c_double), C.byref(K), _p(med, C.c_double)))
        return lab, Hc[: K.value], med[: Hc.shape[0]]

    def fp32_peak(self, variant=1, iters=20000):
        tf = C.c_double(0); ms = C.c_double(0)
        self._check(lib().mh_diag_fp32_peak(self._h, int(variant), int(iters), C.byref(tf), C.byref(ms)))
        return tf.value

    def mma_tf32_peak(self, iters=20000):
        tf = C.c_double(0); r = C.c_double(0)
        self._check(lib().mh_diag_mma_tf32_peak(self._h, int(iters), C.byref(tf), C.byref(r)))
        return tf.value, r.value

    def set_fast_config(self, v):
        self._check(lib().mh_diag_set_fast_config(self._h, int(v)))

    def set_dense_variant(self, v):
        self._check(lib().mh_diag_set_dense_variant(self._h, int(v)))

    def get_fast_config(self):
        return int(lib().mh_diag_get_fast_config(self._h))

    def set_fused_variant(self, v):
        self._check(lib().mh_diag_set_fused_variant(self._h, int(v)))

    @property
    def energy(self):
        return float(lib().mh_get_energy(self._h))

    @property
    def iterations(self):
        return int(lib().mh_get_iterations(self._h))

    def alternating_ms(self):
        ms = np.zeros(5)
        self._check(lib().mh_diag_get_alternating_ms(self._h, _p(ms, C.c_double)))
        return dict(zip(["meanshift", "mode_fit_inlier_scan", "data_cost", "graph_cut", "refit"], ms.tolist()))

    def stage_ms(self):
        ms = np.zeros(5)
        self._check(lib().mh_get_stage_ms(self._h, _p(ms, C.c_double)))
        return dict(zip(["pointwise_h", "stable_clusters", "adjacency", "alternating", "total"], ms.tolist()))
