"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of bench.py into markdown:
python tools/launch_summary.py gpurun_out/launches.csv profiles/rN_bench_launch_summary.md [bench_line.json]"""
import csv, io, json, sys, collections
src, out = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
agg = collections.OrderedDict()
for r in rows:
    name = r["Kernel Name"]
    short = name.split("(")[0]
    ns = float(r["Metric Value"].replace(",", ""))
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1; a[1] += ns
tot = sorted(agg.items(), key=lambda kv: -kv[1][1])
md = ["# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (1 B200)", "",
      "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` (raw: `%s`). Per-launch times are" % src.split("/")[-1],
      "cold-cache and serialised: compare shares, not absolutes.", "", "| kernel | launches | avg ms | total ms |", "|---|---|---|---|"]
for k, (n, ns) in tot:
    md.append(f"| `{k[:90]}` | {n} | {ns / n / 1e6:.3f} | {ns / 1e6:.3f} |")
step = ["haf_kernel", "fused_init_kernel", "cost_argmin_tc_kernel", "cost_argmin_kernel", "wild_", "label_count_kernel", "scan_kernel",
        "label_scatter_kernel", "haf_segment_accumulate_kernel", "acc_count_kernel", "haf_solve_kernel", "split_hyp", "labels_from_best"]
in_step = [(k, v) for k, v in tot if any(s in k for s in step)]
step_ms = sum(v[1] / v[0] for k, v in in_step) / 1e6
dom = max(in_step, key=lambda kv: kv[1][1] / kv[1][0])
md += ["", "One hot-path step = " + " + ".join(f"`{k.split('::')[-1][:40]}`" for k, _ in in_step) + f" = {step_ms:.3f} ms under ncu;",
       f"`{dom[0].split('::')[-1][:40]}` share = {100 * dom[1][1] / dom[1][0] / 1e6 / step_ms:.1f} %"]
if len(sys.argv) > 3:
    b = json.load(open(sys.argv[3]))
    md[-1] += (f" (bench.py's CUDA-event share in the same round: {100 * b['roofline']['kernel_share_of_step']:.1f} %, "
               f"kernel {b['roofline']['kernel_ms']:.2f} ms).")
md += ["", "The `fma_peak_kernel` launches are the in-run FP32 peak probe (roofline denominator), `cost_dense_*` the HBM-bound",
       "member timed for `roofline_dense`; both run after the timed region."]
open(out, "w").write("\n".join(md) + "\n")
print("\n".join(md))
