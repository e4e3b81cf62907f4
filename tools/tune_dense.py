"""A/B timing of the dense cost kernels (mh_data_cost_dense): variants 0..3 x int32/int16 x K, 1M correspondences."""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, multih_b200 as m
n = 1 << 20
sc, pick = bench.make_workload(n)
ctx = m.Context(); ctx.set_geometry(sc.F, sc.pts)
d_pts, d_aff = ctx.upload(sc.pts, sc.aff)
d_h = ctx.haf_hypotheses(d_pts, d_aff)
d_all = torch.cat([ctx.hypotheses_from_host(sc.planes), d_h[torch.from_numpy(pick[:3896] % n).cuda()]]).contiguous()  # 4096
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)
for K in [int(a) for a in sys.argv[1:]] or [1024, 1023, 2048]:
    d_hyp = d_all[:K].contiguous()
    for eb, dt in ((4, torch.int32), (2, torch.int16)):
        od = torch.empty((n, K + 1), dtype=dt, device="cuda")
        ref = None
        for v in (0, 3, 2, 1):
            ctx.set_dense_variant(v)
            for _ in range(2):
                ctx.data_cost_dense(d_pts, d_hyp, elem_bytes=eb, out=od)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                ctx.data_cost_dense(d_pts, d_hyp, elem_bytes=eb, out=od)
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            same = True
            if ref is None: ref = od.clone()
            else: same = bool(torch.equal(od, ref))
            gbs = n * (K + 1) * eb / ms / 1e6
            print(f"K {K} int{eb*8} variant {v}: {ms:.3f} ms {n*(K+1)/ms/1e9:.3f}e12 res/s {gbs:.0f} GB/s ({100*gbs/peak:.1f}% of {peak:.0f}) same {same}", flush=True)
ctx.set_dense_variant(1)
