#!/usr/bin/env python
"""bench.py — the Multi-H hot path on B200 (BASELINE.json metric: correspondence x hypothesis residuals/s).

Workload (config.workload = "cfg4"): BASELINE.json configs[3], the configuration the metric is quoted on — synthetic
200-plane scene, 4 194 304 affine correspondences x 8192 hypotheses (200 generating homographies + 7992 HAF hypotheses
of randomly chosen correspondences), correspondence-sharded over the ranks (strong scaling: the scene is fixed).

One "step" = one pass of the hot path over the scene:
    K1  HAF hypothesis per correspondence                    (MultiH.cpp:696-717, 850-911)
    [N>1] NCCL broadcast of the 8192 x 12 hypothesis block from rank 0
    K2  fused N x K residual / data cost: per-site data-term argmin label + cost and per-hypothesis inlier
        count — nothing N x K touches HBM (MultiH.cpp:473-504, 430-443, 743-768)
    K4  per-label refit statistics from the argmin labels (segmented reduction of SUM A^T A)
    [N>1] ONE NCCL all-reduce of the K x 12 FP64 statistics (inlier counts ride in the pad column)
    K4  batched 4x4 eigen-solves -> refined homographies      (MultiH.cpp:545-599, 913-990)
`value` times that with the inputs resident in HBM; `e2e` times the same pass through the C ABI from PINNED HOST
buffers (H2D of the FP64 correspondences + affines, normalisation, the pass, D2H of labels + refined homographies).

--impl reference times the reference's own CPU implementation of the same path: its dataEnergy (MultiH.cpp:473-504),
compiled unmodified and in place against a mini OpenCV shim (oracle/_ref/libmultih_ref.so, see oracle/ref_multih_wrapper.cpp),
evaluated densely over a bounded sample of the same scene — 65 536 of the 4 194 304 correspondences x all 8192 hypotheses
per step, rows split over all host threads (the reference itself evaluates the term from its single-threaded alpha-expansion);
the rate per residual is what is reported, i.e. the sample is extrapolated to the scene, as SURVEY.md §8(d) prescribes.  Its
pair_e2e is MultiH::Process() of the same compiled source on the bundled pair.  If that library was never built, the FP64
oracle port takes its place (kind "port").
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "correspondence x hypothesis residual evaluations per second (whole job)"
WORKLOAD = ("cfg4: synthetic 200-plane scene, 4M affine correspondences x 8192 hypotheses "
            "(200 planes + 7992 HAF hypotheses), correspondence-sharded")
N_TOTAL = 4 * (1 << 20)
N_PLANES = 200
K_HYP = 8192
SEED = 0xB200 + 3
FLOP_PER_RESIDUAL = 20


def make_workload(n_total=N_TOTAL):
    import multih_b200 as m

    sc = m.scenes.make_scene(n_total, N_PLANES, seed=SEED)
    rng = np.random.Generator(np.random.Philox(SEED + 1))
    pick = rng.integers(0, n_total, size=K_HYP - N_PLANES)
    return sc, pick


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle sampling during the timed region (via NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def summary(self):
        return {"sm_mhz": (float(np.median(self.sm)) if self.sm else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


REF_SAMPLE = 65536   # cfg4 on the CPU side is sub-sampled to 64k x 8192 and extrapolated


def reference_cost_rate(pts, hyp, min_seconds=0.0, passes=1):
    """residual evaluations per second of the reference's dataEnergy over pts x hyp on the host cores.  Returns
    (rate, seconds, passes done, cores, kind)."""
    from concurrent.futures import ThreadPoolExecutor
    import ctypes as C

    from oracle import oracle as orc

    cores = orc.hardware_threads()
    lib = orc.ref_multih_lib()
    K = len(hyp)
    if lib is None:   # the reference was never compiled here: the oracle port
        t0, done = time.perf_counter(), 0
        while done < passes or time.perf_counter() - t0 < min_seconds:
            orc.data_cost_sweep(pts, hyp, threads=cores)
            done += 1
        dt = time.perf_counter() - t0
        return len(pts) * K * done / dt, dt, done, cores, "port"
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    hyp = np.ascontiguousarray(hyp, dtype=np.float64)
    chunk = 512
    bufs = [np.empty((chunk, K + 1), dtype=np.int32) for _ in range(cores)]
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)

    def work(t):
        for lo in range(t * chunk, len(pts), cores * chunk):
            n = min(chunk, len(pts) - lo)
            lib.ref_data_cost_dense(pts[lo:lo + n].ctypes.data_as(dp), n, hyp.ctypes.data_as(dp), K, C.c_double(0.5),
                                    C.c_double(2.2), bufs[t].ctypes.data_as(ip))

    t0, done = time.perf_counter(), 0
    with ThreadPoolExecutor(cores) as ex:
        while done < passes or time.perf_counter() - t0 < min_seconds:
            list(ex.map(work, range(cores)))
            done += 1
    dt = time.perf_counter() - t0
    return len(pts) * K * done / dt, dt, done, cores, "reference"


def reference_pair(root):
    """The bundled pair through MultiH::Process() of the compiled reference source (raw rows that pass the F test)."""
    from oracle import oracle as orc

    g = np.load(os.path.join(root, "tests", "golden", "golden_prefilter.npz"))
    F = g["barr_F"]
    x1 = np.c_[g["barr_pts"][:, :2], np.ones(len(g["barr_pts"]))]; x2 = np.c_[g["barr_pts"][:, 2:], np.ones(len(x1))]
    l = x1 @ F.T
    inl = np.abs(np.einsum("ij,ij->i", x2, l)) / np.hypot(l[:, 0], l[:, 1]) < 2.6   # stands in for the mask of MultiH.cpp:775
    pts, aff = g["barr_pts"][inl], g["barr_aff"][inl]
    out = {"pts": pts, "aff": aff, "F": F}
    if orc.ref_multih_lib() is not None:
        tr = time.perf_counter()
        lab, H, info = orc.ref_process(pts, aff, F, lm=True)
        out.update(ms=(time.perf_counter() - tr) * 1e3, labels=lab, planes=int(len(H)), iterations=int(info["iterations"]),
                   kind="MultiH::Process() of the reference source compiled in place (oracle/_ref/libmultih_ref.so: F injected for "
                        "its RANSAC, MSVC rand(), exact 31-nearest neighbourhood for FLANN), 1 run, 1 thread")
    return out


def run_reference(args):
    """CPU arm: the reference's own dataEnergy swept densely over a bounded sample of cfg4."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc

    sc, pick = make_workload(args.n_total)
    cores = orc.hardware_threads()
    hyp = np.concatenate([sc.planes, orc.haf_hypotheses(sc.pts[pick], sc.aff[pick], sc.F, threads=cores)])
    rows = np.random.Generator(np.random.Philox(SEED + 2)).choice(len(sc.pts), min(REF_SAMPLE, len(sc.pts)), replace=False)
    pts = sc.pts[np.sort(rows)]
    if args.warmup:
        reference_cost_rate(pts[:4096], hyp, passes=min(args.warmup, 2))
    value, dt, done, cores, kind = reference_cost_rate(pts, hyp, passes=args.steps)
    sample = (f"{len(pts)} of the scene's {len(sc.pts)} correspondences (random rows) x {K_HYP} hypotheses per step, "
              f"{'dataEnergy of the reference source (MultiH.cpp:473-504) compiled in place' if kind == 'reference' else 'FP64 oracle port of dataEnergy'}"
              f", rows split over {cores} host threads; rate per residual, i.e. extrapolated to the scene")
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "residuals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / done * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "correspondences": args.n_total, "hypotheses": K_HYP,
                   "step": f"bounded sample: {len(pts)} correspondences x {K_HYP} hypotheses per step (extrapolated per residual)"},
        "cpu_baseline": {"value": value, "unit": "residuals/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "residuals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    try:
        rp = reference_pair(ROOT)
        if "ms" in rp:
            line["pair_e2e"] = {"workload": f"bundled barrsmith pair, {len(rp['pts'])} raw correspondences that pass the F test",
                                "ms_per_pair": rp["ms"], "planes": rp["planes"], "iterations": rp["iterations"],
                                "outlier_fraction": float((rp["labels"] < 0).mean()), "kind": rp["kind"]}
    except Exception as e:
        line["pair_e2e"] = {"unavailable": repr(e)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-total", type=int, default=N_TOTAL)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", action="store_true", help="skip the cfg5 batched-pairs leg")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the cfg3 (100k correspondences through mh_process) leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    import multih_b200 as m

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # ---- workload ----------------------------------------------------------------------------------------------------
    n_total = args.n_total
    sc, pick = make_workload(n_total)
    lo, hi = m.dist.shard_range(n_total, rank, world)
    n_loc = hi - lo
    ctx = m.Context(device=local)
    if os.environ.get("MH_FAST_CONFIG"):  # K2 variant override for tuning runs (default: the library's)
        ctx.set_fast_config(int(os.environ["MH_FAST_CONFIG"]))
    ctx.set_geometry(sc.F, sc.pts)  # same strided sample on every rank -> identical normalisation
    h_pts = torch.from_numpy(sc.pts[lo:hi]).pin_memory()
    h_aff = torch.from_numpy(sc.aff[lo:hi]).pin_memory()
    d_pts, d_aff = ctx.upload(h_pts, h_aff)
    # hypothesis block: rank 0 builds it (K1 on the picked correspondences), then it is broadcast every step
    d_hyp = torch.zeros((K_HYP, 12), dtype=torch.float32, device=dev)
    if rank == 0:
        p_pts, p_aff = ctx.upload(sc.pts[pick], sc.aff[pick])
        d_hyp[:N_PLANES] = ctx.hypotheses_from_host(sc.planes)
        d_hyp[N_PLANES:] = ctx.haf_hypotheses(p_pts, p_aff)
        del p_pts, p_aff
    d_hyp_pt = torch.empty((n_loc, 12), dtype=torch.float32, device=dev)
    fused = {"best": torch.empty(n_loc, dtype=torch.int64, device=dev),
             "inliers": torch.empty(K_HYP, dtype=torch.int32, device=dev)}
    labels = torch.empty(n_loc, dtype=torch.int32, device=dev)
    acc = torch.empty((K_HYP, 12), dtype=torch.float64, device=dev)
    d_ref = torch.empty((K_HYP, 12), dtype=torch.float32, device=dev)
    h_labels = torch.empty(n_loc, dtype=torch.int32).pin_memory()
    h_ref = torch.empty((K_HYP, 12), dtype=torch.float32).pin_memory()
    # L2 policy (timing rule: flush L2 between timed iterations OR use inputs larger than L2): no flush — consecutive steps read
    # DIFFERENT resident copies of the rank's inputs and write their own outputs; the copies of one rotation total >= 2 x L2, so
    # by the time a copy comes round again the chip has streamed more than two L2s of other inputs (plus all outputs) through it.
    l2_bytes = int(getattr(torch.cuda.get_device_properties(dev), "L2_cache_size", 126 << 20))
    in_bytes = (d_pts.numel() + d_aff.numel()) * 4
    n_sets = max(2, -(-2 * l2_bytes // max(in_bytes, 1)))
    sets = [{"pts": d_pts, "aff": d_aff, "hyp_pt": d_hyp_pt, "best": fused["best"], "labels": labels}]
    for _ in range(n_sets - 1):
        sets.append({"pts": d_pts.clone(), "aff": d_aff.clone(), "hyp_pt": torch.empty_like(d_hyp_pt),
                     "best": torch.empty_like(fused["best"]), "labels": torch.empty_like(labels)})
    l2_note = (f"no flush: consecutive steps read different resident copies of the rank's inputs, {n_sets} x {in_bytes / 2**20:.0f} MiB "
               f"per rotation >= 2 x {l2_bytes / 2**20:.0f} MiB L2, and write their own outputs")
    step_no = [0]
    k2_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
             for _ in range(args.steps + args.warmup + 1)]

    # Multi-GPU: the library owns the communicator (mh_comm_init) and the whole sharded pass (mh_step_sharded): hypothesis
    # broadcast on the communicator's stream under K1, K2, labels, K4 statistics, ONE all-reduce (statistics + inlier counts)
    # that overlaps the next pass (double-buffered), K4 solves.  torch.distributed only carries the 128-byte id and the
    # bench's own bookkeeping (barrier, max over ranks).  Everything of every step completes inside the timed region
    # (finish_pending() before the closing event).
    if world > 1:
        ctx.comm_init(rank, world)
    for pair in k2_ev:          # torch creates the cudaEvent_t lazily: record once so that the handles exist
        for e in pair:
            e.record()

    def finish_pending():
        ctx.step_sharded_finish()

    def hot_pass(ev=None):
        b = sets[step_no[0] % n_sets]
        step_no[0] += 1
        ctx.step_sharded(b["pts"], b["aff"], d_hyp, b["hyp_pt"], b["best"], b["labels"], fused["inliers"], d_ref, events=ev)

    # end-to-end: the host->device uploads run on their own stream (second context = second stream, same geometry), the
    # device->host downloads on a third.
    s_up = torch.cuda.Stream(device=dev, priority=-1)   # high priority: the small normalisation kernels must not queue behind K2
    ctx_up = m.Context(device=local, use_torch_stream=False)
    ctx_up.use_stream(s_up)
    ctx_up.set_geometry(sc.F, sc.pts)
    # Consecutive e2e steps are pipelined, not serialised: step i+1's uploads (own stream) run under step i's kernels, every step
    # still copies ITS inputs host->device and ITS results device->host inside the timed region; the buffer sets of the rotation
    # double-buffer the device side, events hand them back to the upload stream, the host only synchronises at the end.
    ev_done = [torch.cuda.Event() for _ in range(n_sets)]
    h_labels2 = [h_labels, torch.empty_like(h_labels).pin_memory()]
    h_ref2 = [h_ref, torch.empty_like(h_ref).pin_memory()]
    e2e_no = [0]

    s_down = torch.cuda.Stream(device=dev)
    _dbg = os.environ.get("MH_E2E_DEBUG", "")   # measurement aid: "noup" / "nodown" drop the copies (the line is then not an e2e number)
    ev_in = [torch.cuda.Event() for _ in range(n_sets)]
    ev_res = torch.cuda.Event()
    for b_ in sets:
        b_["ref"] = torch.empty_like(d_ref)
    last_done = [None]

    def e2e_pass():
        """uploads (own stream) -> the same mh_step_sharded call the device-resident pass makes -> downloads (own stream)"""
        i = e2e_no[0]; e2e_no[0] += 1
        j = i % n_sets
        b = sets[j]
        main = torch.cuda.current_stream()
        s_up.wait_event(ev_done[j])                                                          # the set's previous pass is done with it
        if "noup" not in _dbg:
            ctx_up.upload(h_pts, None, out=(b["pts"], None))                                # H2D + normalise (points)
            ctx_up.upload(None, h_aff, out=(None, b["aff"]))                                # H2D + normalise (affines)
        ev_in[j].record(s_up)
        main.wait_event(ev_in[j])
        ctx.step_sharded(b["pts"], b["aff"], d_hyp, b["hyp_pt"], b["best"], b["labels"], fused["inliers"], b["ref"])
        ctx.step_sharded_finish()                                                            # this step's all-reduce + solves
        ev_res.record(main)
        s_down.wait_event(ev_res)
        with torch.cuda.stream(s_down):
            if "nodown" not in _dbg:
                h_labels2[i & 1].copy_(b["labels"], non_blocking=True)                       # D2H results
                h_ref2[i & 1].copy_(b["ref"], non_blocking=True)
        ev_done[j].record(s_down)
        last_done[0] = ev_done[j]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------------------------------
    for i in range(args.warmup):
        hot_pass(k2_ev[i])
    finish_pending()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launches
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    align = torch.zeros(1, device=dev)

    def align_ranks():
        # the host barrier releases the ranks up to a millisecond apart; a one-element all-reduce on the compute stream makes
        # the ranks' STREAMS start the timed region together (outside the timed region: it precedes the opening event)
        if world > 1:
            dist.all_reduce(align)

    align_ranks()
    t0.record()
    for i in range(args.steps):
        hot_pass(k2_ev[args.warmup + i])
    finish_pending()
    t1.record()
    barrier()
    sampler.stop_flag = True
    launches = ctx.launches - launches0
    ms_total = t0.elapsed_time(t1)
    k2_ms = float(np.mean([a.elapsed_time(b) for a, b in k2_ev[args.warmup:args.warmup + args.steps]]))
    t = torch.tensor([ms_total, k2_ms], dtype=torch.float64, device=dev)
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt)
    ms_total, k2_ms = float(t[0]), float(t[1])
    ms_per_step = ms_total / args.steps
    value = n_total * K_HYP / (ms_per_step * 1e-3)

    # ---- end-to-end timing (host buffers) -----------------------------------------------------------------------------------
    for _ in range(2):
        e2e_pass()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    align_ranks()
    e0.record()
    for _ in range(args.steps):
        e2e_pass()
    torch.cuda.current_stream().wait_event(last_done[0])   # the last step's downloads belong to the timed region
    e1.record()
    barrier()
    te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_total * K_HYP / (float(te[0]) / args.steps * 1e-3)
    h2d = (h_pts.numel() + h_aff.numel()) * 8 * world
    d2h = (h_labels.numel() * 4) * world + h_ref.numel() * 4 * world

    # ---- BASELINE.json configs[4]: independent 5k-correspondence pairs, round-robin over the ranks; on every GPU several host
    # threads each drive their own context (own stream), so the host graph-cut of one pair overlaps the kernels of another.
    # Every rank takes PAIRS_PER_RANK pairs (128: at 8 GPUs that is the 1024-pair workload of configs[4]); pairs/s = all pairs
    # over the slowest rank's wall time.
    batched = None
    if not args.no_batched:
        try:
            from concurrent.futures import ThreadPoolExecutor

            per_rank = int(os.environ.get("MH_PAIRS_PER_RANK", "128"))
            host_cores = os.cpu_count() or 8
            n_thr = max(2, min(16, host_cores // max(world, 1)))
            os.environ["MH_GC_THREADS"] = "1"   # throughput mode: one host thread per pair instead of a move pool per pair
            ids = [rank + world * i for i in range(per_rank)]          # round-robin: pair p goes to rank p % world
            scenes = [m.scenes.make_scene(5000, 3 + (i % 6), seed=0xB200 + 4 + i) for i in ids]
            ctxs = [m.Context(device=local, use_torch_stream=False) for _ in range(n_thr)]

            def run_thread(t):
                torch.cuda.set_device(local)
                return [int(ctxs[t].process(sc_.pts, sc_.aff, sc_.F)[2]) for sc_ in scenes[t::n_thr]]

            for c in ctxs:
                c.process(scenes[0].pts, scenes[0].aff, scenes[0].F)  # warm-up: allocations
            barrier()
            tb = time.perf_counter()
            with ThreadPoolExecutor(n_thr) as ex:
                planes = sum(ex.map(run_thread, range(n_thr)), [])
            dtb = time.perf_counter() - tb
            tt = torch.tensor([dtb, float(np.sum(planes))], dtype=torch.float64, device=dev)
            if world > 1:
                tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                dist.all_reduce(tt)
                dtb, planes_sum = float(tmax[0]), float(tt[1])
            else:
                planes_sum = float(tt[1])
            os.environ.pop("MH_GC_THREADS", None)
            batched = {"workload": f"cfg5: {per_rank * world} independent synthetic pairs x 5000 correspondences (3-8 planes) through "
                                   "mh_process (host buffers in, labels + homographies out), pairs round-robin over the ranks",
                       "pairs": per_rank * world, "pairs_per_rank": per_rank, "contexts_per_gpu": n_thr,
                       "pairs_per_s": per_rank * world / dtb, "ms_per_pair_amortised": dtb / (per_rank * world) * 1e3,
                       "mean_planes": planes_sum / (per_rank * world), "n_gpus": world}
        except Exception as e:
            batched = {"unavailable": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K2 fused): FP32 CUDA-core pipe --------------------------------------------------------
    peak_scalar = max(ctx.fp32_peak(0, 20000) for _ in range(3))
    peak_packed = max(ctx.fp32_peak(1, 20000) for _ in range(3))
    peak = max(peak_scalar, peak_packed)
    achieved = FLOP_PER_RESIDUAL * (n_loc * K_HYP) / (k2_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k2_fused_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch") * (n_loc / N_TOTAL)  # per 4M x 8192 launch
        except Exception:
            traffic = None
    roofline = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic,
                "kernel": ("cost_argmin_tc_kernel" if ctx.get_fast_config() >= 30 else "cost_argmin_kernel")
                          + f" (fast config {ctx.get_fast_config()})",
                "algorithmic": f"{FLOP_PER_RESIDUAL} flop/residual x {n_loc} x {K_HYP} per launch",
                "kernel_ms": k2_ms, "kernel_share_of_step": k2_ms / ms_per_step,
                "peak_source": "measured in this run: dependent-FFMA probe (mh_diag_fp32_peak), "
                               f"scalar FFMA {peak_scalar:.1f} / packed FFMA2 {peak_packed:.1f} TFLOP/s; "
                               "MEASURED_PEAKS.json carries only HBM and bf16-tensor peaks; nominal 148 SM x 128 lanes x "
                               "2 x 1.965 GHz = 74.4"}
    # the HBM-bound member of the same kernel family (materialised dense cost matrix), timed alone for the record
    kd = 1024
    nd = min(n_loc, 1 << 20)
    try:
        hbm_peak, hbm_src = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md"

    def time_dense(elem_bytes, dtype):
        od = torch.empty((nd, kd + 1), dtype=dtype, device=dev)
        for _ in range(2):
            ctx.data_cost_dense(d_pts[:nd], d_hyp[:kd], elem_bytes=elem_bytes, out=od)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            ctx.data_cost_dense(d_pts[:nd], d_hyp[:kd], elem_bytes=elem_bytes, out=od)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        return ms, nd * (kd + 1) * elem_bytes / (ms * 1e-3) / 1e9

    ms32, gbs32 = time_dense(4, torch.int32)
    ms16, gbs16 = time_dense(2, torch.int16)
    dense_traffic = None
    try:
        dt = json.load(open(os.path.join(ROOT, "profiles", "k2_dense_traffic.json")))["int32"]
        if dt["n"] == nd and dt["k"] == kd:
            dense_traffic = dt["dram_bytes_read"] + dt["dram_bytes_write"]
    except Exception:
        pass
    roofline_dense = {"bound": "hbm", "achieved": gbs32, "peak": hbm_peak, "unit": "GB/s", "frac": gbs32 / hbm_peak,
                      "traffic": dense_traffic, "kernel": "cost_dense_tiled_kernel<int32>", "peak_source": hbm_src,
                      "algorithmic": f"4 B/residual x {nd} x {kd + 1} per launch (the array GCO's setDataCost(int*) indexes)",
                      "kernel_ms": ms32,
                      "int16": {"achieved": gbs16, "frac": gbs16 / hbm_peak, "kernel_ms": ms16,
                                "note": "2 B/residual: bound by the kernel's FP32 arithmetic, not by HBM (with the row heads on the "
                                        "compute warps int16 and int32 launches take the same time); 0.70 of HBM would need "
                                        "2.3e12 residuals/s, more than the fused argmin kernel reaches"}}

    # ---- CPU baseline: the reference's own dataEnergy on a bounded sample ---------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline and world == 1:  # rank 0 at N=1 only
        hyp_host = ctx.hypotheses_to_host(d_hyp)
        rows = np.sort(np.random.Generator(np.random.Philox(SEED + 2)).choice(n_total, min(REF_SAMPLE, n_total), replace=False))
        reference_cost_rate(sc.pts[rows[:2048]], hyp_host, passes=1)
        rate, dtc, reps, cores, kind = reference_cost_rate(sc.pts[rows], hyp_host, min_seconds=10.0)
        cpu = {"value": rate, "unit": "residuals/s", "cores": cores, "kind": kind,
               "sample": f"{reps} x ({len(rows)} random correspondences of the scene x {K_HYP} hypotheses), "
                         + ("the reference source's dataEnergy (MultiH.cpp:473-504) compiled in place (oracle/_ref/libmultih_ref.so)"
                            if kind == "reference" else "FP64 oracle port of dataEnergy")
                         + f", rows split over {cores} host threads, {dtc:.1f} s; rate per residual (extrapolated to the scene)"}

    # ---- second half of BASELINE.json's metric: end-to-end ms per image pair (bundled barrsmith pair, configs[0]/[1]) --------
    # the RAW rows of the bundled pair that pass the F test, through mh_process exactly as the MultiH shims run it — refinement
    # filter, LM-polished 3PT fits, compatibility check: the reference's Process() — next to MultiH::Process() of the reference
    # source compiled in place, on this host
    pair = None
    if world == 1 and os.path.exists(os.path.join(ROOT, "tests", "golden", "golden_prefilter.npz")):
        rp = reference_pair(ROOT) if not args.no_cpu_baseline else None
        if rp is None:
            rp = {}
            g = np.load(os.path.join(ROOT, "tests", "golden", "golden_prefilter.npz"))
            F = g["barr_F"]
            x1 = np.c_[g["barr_pts"][:, :2], np.ones(len(g["barr_pts"]))]; x2 = np.c_[g["barr_pts"][:, 2:], np.ones(len(x1))]
            l = x1 @ F.T
            inl = np.abs(np.einsum("ij,ij->i", x2, l)) / np.hypot(l[:, 0], l[:, 1]) < 2.6
            rp.update(pts=g["barr_pts"][inl], aff=g["barr_aff"][inl], F=F)
        pctx = m.Context(m.capi.default_params(prefilter=1), device=local)
        pctx.process(rp["pts"], rp["aff"], rp["F"])  # warm-up (allocations, module load)
        reps, tp = 3, time.perf_counter()
        for _ in range(reps):
            lab, Hh, Kp = pctx.process(rp["pts"], rp["aff"], rp["F"])
        kept = lab > -2
        pair = {"workload": f"bundled barrsmith pair: the {len(lab)} raw correspondences that pass the F test, through mh_process as "
                            "MultiH::Process() runs (refinement filter, LM-polished 3PT fits, alternating optimisation with the host "
                            "graph-cut, compatibility check); host buffers in, labels + homographies out",
                "ms_per_pair": (time.perf_counter() - tp) / reps * 1e3, "kept": int(kept.sum()), "planes": int(Kp),
                "outlier_fraction": float((lab[kept] < 0).mean()), "iterations": pctx.iterations, "stage_ms": pctx.stage_ms(),
                "alternating_ms": pctx.alternating_ms(),
                "shipped_result": "Executable/results/barrsmith/result_barrsmith.txt (another build, OpenCV RANSAC, unseeded rand()): "
                                  "1094 kept, 5 planes, 17 % outliers"}
        if "labels" in rp:
            same = len(rp["labels"]) == int(kept.sum())
            pair["cpu_reference"] = {"ms_per_pair": rp["ms"], "planes": rp["planes"], "iterations": rp["iterations"],
                                     "kind": rp["kind"],
                                     "label_agreement": float((rp["labels"] == lab[kept]).mean()) if same else None}

    # ---- BASELINE.json configs[2]: the 20-plane scene with 100 000 correspondences through mh_process on one GPU -------------------
    cfg3 = None
    if world == 1 and not args.no_cfg3:
        try:
            sc3 = m.scenes.make_scene(100_000, 20, seed=0xB200 + 2)
            c3 = m.Context(device=local)
            t3 = time.perf_counter()
            lab3, H3, K3 = c3.process(sc3.pts, sc3.aff, sc3.F)
            dt3 = time.perf_counter() - t3
            cfg3 = {"workload": "cfg3: synthetic 20-plane scene, 100 000 affine correspondences, 0.5 px noise, 50 % outliers, through "
                                "mh_process (host buffers in, labels + homographies out)",
                    "s_per_scene": dt3, "planes": int(K3), "iterations": c3.iterations, "outlier_fraction": float((lab3 < 0).mean()),
                    "stage_ms": c3.stage_ms(), "alternating_ms": c3.alternating_ms(),
                    "note": "the alternating optimisation is the host alpha-expansion (north star: it stays on the host): 100 000 sites "
                            "x 70-200 labels per labelling step"}
            if not args.no_cpu_baseline:
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                from ref_pipeline import oracle_process

                n_sub = 4000   # the oracle pipeline (FP64 restatement + the reference's own GCO) on a sub-sample of the scene type
                scs = m.scenes.make_scene(n_sub, 20, seed=0xB200 + 2)
                tg = time.perf_counter()
                lg, Hg, Kg = c3.process(scs.pts, scs.aff, scs.F)
                tg = time.perf_counter() - tg
                to = time.perf_counter()
                lo_, Ho_, io_ = oracle_process(scs.pts, scs.aff, scs.F, compatibility_check=True, lm=True)
                to = time.perf_counter() - to
                cfg3["sub_sample"] = {"correspondences": n_sub, "planes_gpu": int(Kg), "planes_oracle": int(len(Ho_)),
                                      "label_agreement": float((lg == lo_).mean()), "s_gpu_path": tg, "s_oracle_pipeline": to,
                                      "kind": "same generator, 4000 correspondences: mh_process vs the oracle pipeline with the "
                                              "reference's own alpha-expansion (which equals MultiH::Process() of the reference source)"}
        except Exception as e:
            cfg3 = {"unavailable": repr(e)}

    line = {
        "metric": METRIC,
        "value": value, "unit": "residuals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "correspondences": n_total, "hypotheses": K_HYP, "per_rank": n_loc,
                   "step": "K1 HAF + K2 fused cost/argmin/inlier-count + K4 refit (+ NCCL bcast/all-reduce for N>1)",
                   "l2": l2_note},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": "residuals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": float(te[0]) / args.steps},
        "gpu_launches": int(lt[0]),
        "roofline": roofline, "roofline_dense": roofline_dense, "cpu_baseline": cpu, "pair_e2e": pair,
        "batched_pairs": batched, "cfg3_e2e": cfg3,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
