"""Stand-alone version of bench.py's batched-pairs leg with progress output (debug aid)."""
import faulthandler, os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(int(os.environ.get("PROBE_DUMP_S", "45")), exit=True)
import torch
import multih_b200 as m
from concurrent.futures import ThreadPoolExecutor
n_pairs, n_thr = int(sys.argv[1]), int(sys.argv[2])
own = (sys.argv[3] == "own") if len(sys.argv) > 3 else True
scenes = [m.scenes.make_scene(5000, 3 + (i % 6), seed=0xB200 + 4 + i) for i in range(n_pairs)]
ctxs = [m.Context(device=0, use_torch_stream=not own) for _ in range(n_thr)]
print("contexts ready", flush=True)
for c in ctxs:
    c.process(scenes[0].pts, scenes[0].aff, scenes[0].F)
print("warm", flush=True)
def run_thread(t):
    torch.cuda.set_device(0)
    out = []
    for k, sc_ in enumerate(scenes[t::n_thr]):
        t0 = time.perf_counter()
        out.append(int(ctxs[t].process(sc_.pts, sc_.aff, sc_.F)[2]))
        print(f"thread {t} pair {k} planes {out[-1]} {1e3 * (time.perf_counter() - t0):.0f} ms", flush=True)
    return out
tb = time.perf_counter()
with ThreadPoolExecutor(n_thr) as ex:
    planes = sum(ex.map(run_thread, range(n_thr)), [])
dt = time.perf_counter() - tb
print(f"{n_pairs} pairs, {n_thr} threads: {n_pairs / dt:.1f} pairs/s, {dt / n_pairs * 1e3:.1f} ms/pair amortised", flush=True)
