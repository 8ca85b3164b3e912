"""Stage timing of the BASELINE workload on one GPU (tuning helper, not the bench).
   VGB_LIB=<variant .so> python tools/perf_probe.py [curves] [reps] [final]
   "final" also times the final-layout run (uv stream, rebased indices, draw-command table)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vg_renderer_b200 as vgb  # noqa: E402
from vg_renderer_b200 import workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
b = vgb.Batch(0)
b.add_cubic_strokes(W.config3_cubics(n), width=4.0, cap=1, join=1)
b.upload()


def timed(tag, bytes_per_vertex):
    for _ in range(3):
        b.run()
    acc = {}
    for _ in range(reps):
        b.run()
        for k, v in b.last_run_ms().items():
            acc.setdefault(k, []).append(v)
    s = b.sizes()
    med = {k: float(np.median(v)) for k, v in acc.items()}
    gb = (32 * n + bytes_per_vertex * s.num_vertices + 2 * s.num_indices) / 1e9
    print(tag, end=" ")
    return s, med, gb


s, med, gb = timed("local ", 12)
print(f"{os.path.basename(vgb.lib_path()):28s} total {med['total']:.3f} ms  flatten {med['flatten']:.3f}  count {med['stroke_count']:.3f}  "
      f"emit {med['emit']:.3f}  | {s.num_vertices / med['total'] / 1e3:.0f} Mverts/s  emit {gb / med['emit'] * 1e3:.0f} GB/s  launches {s.kernel_launches}")
if "final" in sys.argv[3:]:
    b.set_final_layout(vgb.make_final_cfg(65536))
    s, med, gb = timed("final ", 16)
    uv, cmds = b.download_final(want_uv=False)
    print(f"{os.path.basename(vgb.lib_path()):28s} total {med['total']:.3f} ms  flatten {med['flatten']:.3f}  count+segment {med['stroke_count']:.3f}  "
          f"emit+uv {med['emit']:.3f}  | {s.num_vertices / med['total'] / 1e3:.0f} Mverts/s  emit {gb / med['emit'] * 1e3:.0f} GB/s  "
          f"launches {s.kernel_launches}  draw commands {len(cmds)}")
