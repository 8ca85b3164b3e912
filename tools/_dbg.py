import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vg_renderer_b200 as vgb
from vg_renderer_b200 import workloads as W
n = 1_000_000
b = vgb.Batch(0)
b.add_cubic_strokes(W.config3_cubics(n), width=4.0, cap=1, join=1)
b.upload()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
em = []
for i in range(reps):
    b.run()
    em.append(round(b.last_run_ms()["emit"], 2))
# compress: print run-length of values rounded to 0.1
out, prev, cnt = [], None, 0
for v in em:
    k = round(v, 1)
    if k == prev:
        cnt += 1
    else:
        if prev is not None:
            out.append(f"{prev}x{cnt}")
        prev, cnt = k, 1
out.append(f"{prev}x{cnt}")
print(os.environ.get("VGB_EMIT_PER_SM"), " ".join(out))
