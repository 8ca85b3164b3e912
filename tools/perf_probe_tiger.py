"""Stage timing of the tiger-like scene x M instances on one GPU (BASELINE configs[4] shape)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vg_renderer_b200 as vgb
from vg_renderer_b200 import workloads as W
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
t0 = time.time()
ti = W.instance_tables(W.tiger_like(), M)
b = vgb.Batch(0)
b.add_tables(ti)
t1 = time.time()
b.upload()
for _ in range(3):
    b.run(index_u32=True)
acc = {}
for _ in range(5):
    b.run(index_u32=True)
    for k, v in b.last_run_ms().items():
        acc.setdefault(k, []).append(v)
s = b.sizes()
med = {k: float(np.median(v)) for k, v in acc.items()}
print(f"tiger-like x{M}: build {t1 - t0:.1f}s  total {med['total']:.3f} ms  flatten {med['flatten']:.3f}  count {med['stroke_count']:.3f}  emit {med['emit']:.3f} | "
      f"{s.num_vertices / med['total'] / 1e3:.0f} Mverts/s  verts {s.num_vertices}  idx {s.num_indices}  meshes {s.num_meshes}  launches {s.kernel_launches}")

# instance replay of the same scene (all instances at the cached scale): one tessellation + M transformed copies
src = vgb.Batch(0)
src.add_tables(W.tiger_like())
src.run(index_u32=True)
rng = np.random.default_rng(1)
a = rng.uniform(0, 2 * np.pi, M)
xf = np.stack([np.cos(a), np.sin(a), -np.sin(a), np.cos(a), rng.uniform(0, 4096, M), rng.uniform(0, 4096, M)], axis=1).astype(np.float32)
ts = []
for _ in range(6):
    src.mark(0)
    src.replay(xf)
    src.mark(1)
    ts.append(src.elapsed_ms(0, 1))
rs = src.replay_sizes()
ms = float(np.median(ts[1:]))
gb = (rs.num_vertices * 12 + rs.num_indices * 4) / 1e9
print(f"replay x{M}: {ms:.3f} ms  {rs.num_vertices / ms / 1e3:.0f} Mverts/s  write {gb / ms * 1e3:.0f} GB/s  verts {rs.num_vertices}  idx {rs.num_indices}")
