"""BASELINE configs[3]: 1M rounded-rect convex AA fills, throughput sweep over segment count
(r in {2,8,32,128,512,2048} -> N polygon points), GPU pipeline vs the compiled reference on one core.
Writes profiles/r01_config4_sweep.json.  python tools/sweep_config4.py [rects]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vg_renderer_b200 as vgb
from vg_renderer_b200 import workloads as W
from vg_renderer_b200.tables import BatchTables
from oracle.pyref import RefLib, have_ref

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ref = RefLib("sse") if have_ref("sse") else None
refs = RefLib("scalar") if have_ref("scalar") else None
rows = []
b = vgb.Batch(0)
for r in (2, 8, 32, 128, 512, 2048):
    q = W.config4_rrects(n, r)
    b.reset(1.0)
    b.add_rounded_rect_fills(q)
    b.upload()
    for _ in range(3):
        b.run()
    ms = []
    for _ in range(7):
        b.run()
        ms.append(b.last_run_ms())
    s = b.sizes()
    med = {k: float(np.median([m[k] for m in ms])) for k in ms[0]}
    V, I = int(s.num_vertices), int(s.num_indices)
    alg = 20 * n + 12 * V + 2 * I
    row = dict(r=r, polygon_points=int(s.num_points) // n, verts=V, indices=I, ms=med, gpu_mverts_s=V / med["total"] / 1e3,
               emit_gbs=alg / med["emit"] / 1e6)
    for name, lib in (("cpu_sse_1core_mverts_s", ref), ("cpu_scalar_1core_mverts_s", refs)):
        if lib is not None:
            m = 20000
            t = BatchTables.rounded_rect_fills(W.config4_rrects(m, r))
            lib.run_batch(t, store=False, path_range=(0, 1000))
            t0 = time.perf_counter()
            o = lib.run_batch(t, store=False)
            row[name] = o["num_vertices"] / (time.perf_counter() - t0) / 1e6
    rows.append(row)
    print(row, flush=True)
out = dict(workload=f"{n} rounded rects pathRoundedRect(10+i%7, 10, 4r+50, 3r+20, r) + strokerConvexFillAA, uint16 indices",
           note="CPU columns: compiled reference (path build + fill), one core, 20k rects", rows=rows)
os.makedirs("profiles", exist_ok=True)
json.dump(out, open("profiles/r01_config4_sweep.json", "w"), indent=1)
