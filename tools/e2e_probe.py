"""where does an end-to-end step spend its time (tuning helper)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vg_renderer_b200 as vgb
from vg_renderer_b200 import workloads as W
from vg_renderer_b200.api import pinned_empty
n = 1_000_000
p8 = W.config3_cubics(n)
b = vgb.Batch(0)
b.add_cubic_strokes(p8, width=4.0, cap=1, join=1); b.run(); b.sync()
s = b.sizes(); V, I = int(s.num_vertices), int(s.num_indices)
pos, _ = pinned_empty((V + 16, 2), np.float32); col, _ = pinned_empty((V + 16,), np.uint32); idx, _ = pinned_empty((I + 16,), np.uint16)
pin, _ = pinned_empty((n, 8), np.float32); pin[:] = p8
mvo, _ = pinned_empty((n + 1,), np.uint64); mio, _ = pinned_empty((n + 1,), np.uint64)
for it in range(3):
    t = [time.perf_counter()]
    b.reset(1.0); t.append(time.perf_counter())
    b.add_cubic_strokes(pin, width=4.0, cap=1, join=1); t.append(time.perf_counter())
    b.upload(); b.sync(); t.append(time.perf_counter())
    b.run(); b.sync(); t.append(time.perf_counter())
    b._chk(b.L.vgb_batch_download(b.h, pos.ctypes.data, col.ctypes.data, idx.ctypes.data, len(pos), len(idx), None, None)); t.append(time.perf_counter())
    b._chk(b.L.vgb_batch_download(b.h, None, None, None, 0, 0, mvo.ctypes.data, mio.ctypes.data)); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print("reset %.2f  record %.2f  upload %.2f  run %.2f  d2h(mesh) %.2f [%.1f GB/s]  d2h(table) %.2f  total %.2f ms" % (*d[:4], d[4], (12 * V + 2 * I) / d[4] / 1e6, d[5], d.sum()))
