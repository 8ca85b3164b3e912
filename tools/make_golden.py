"""Generate tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/libvgref.so: the
reference's own path.cpp/stroker.cpp, VG_CONFIG_ENABLE_SIMD=0, bx shim with vgb_math.h).
Needs /root/reference (via `make -C oracle ref`); the outputs are committed so the GPU box and
CI can check the oracle and the kernels without the reference tree.

    python tools/make_golden.py

Every file records which shim variant produced it (variant: 0 = scalar + shared fp32 math)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyref import RefLib  # noqa: E402
from vg_renderer_b200 import workloads as W  # noqa: E402
from vg_renderer_b200.tables import BatchTables  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import mixed_recorder, polyline_cases  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
ref = RefLib("scalar")
assert ref.variant() == 0


def pack(prefix, out, d):
    for k in ("pos", "color", "index", "mesh_vtx_off", "mesh_idx_off"):
        d[f"{prefix}_{k}"] = out[k]
    d[f"{prefix}_points"] = np.uint64(out["num_points"])


# 1. config 1 KAT: single cubic, every op x cap x join
d = {"variant": np.uint32(ref.variant())}
pts, subs = ref.flatten([0, 2], W.config1_cubic()[0])
d["polyline"] = pts
for op in (0, 2, 3):
    for cap in (0, 1, 2):
        for join in (0, 1, 2):
            pos, col, idx = ref.mesh(op, pts, False, 0xFF0000FF, 4.0, cap, join)
            d[f"op{op}_c{cap}_j{join}_pos"] = pos
            d[f"op{op}_c{cap}_j{join}_idx"] = idx
            if col is not None:
                d[f"op{op}_c{cap}_j{join}_col"] = col
np.savez_compressed(os.path.join(OUT, "config1_kat.npz"), **d)

# 2. polyline edge cases x cap x join x closed x op
d = {"variant": np.uint32(ref.variant())}
for ci, (name, c) in enumerate(polyline_cases(small=True)):
    d[f"case{ci}_pts"] = np.asarray(c, np.float32)
    for op in (0, 2, 3):
        for cap in (0, 1, 2):
            for join in (0, 1, 2):
                for closed in (0, 1):
                    if closed and len(c) < 3:
                        continue
                    pos, col, idx = ref.mesh(op, c, bool(closed), 0x80FF8040, 6.0, cap, join)
                    k = f"case{ci}_op{op}_c{cap}_j{join}_z{closed}"
                    d[k + "_pos"] = pos
                    d[k + "_idx"] = idx
    if len(c) >= 3:
        for op in (1, 4):
            pos, col, idx = ref.mesh(op, c, True, 0xFF102030)
            d[f"case{ci}_op{op}_pos"] = pos
            d[f"case{ci}_op{op}_idx"] = idx
np.savez_compressed(os.path.join(OUT, "polylines.npz"), **d)

# 3. batches: random cubics (all cap/join), rounded rects, the mixed-builder recorder
d = {"variant": np.uint32(ref.variant())}
for cap in (0, 1, 2):
    for join in (0, 1, 2):
        t = BatchTables.cubic_strokes(W.config3_cubics(60, first=100 * (3 * cap + join)), width=4.0, cap=cap, join=join)
        pack(f"cubics_c{cap}_j{join}", ref.run_batch(t), d)
for r in (2, 8, 32, 128):
    pack(f"rrect_r{r}", ref.run_batch(BatchTables.rounded_rect_fills(W.config4_rrects(8, r))), d)
pack("mixed", ref.run_batch(mixed_recorder().tables()), d)
np.savez_compressed(os.path.join(OUT, "batches.npz"), **d)

# 4. totals + order-sensitive checksum of a larger stream (size-independent property anchor)
t = BatchTables.cubic_strokes(W.config3_cubics(20000), width=4.0, cap=1, join=1)
o = ref.run_batch(t, store=False)
np.savez_compressed(os.path.join(OUT, "stream20k.npz"), variant=np.uint32(ref.variant()),
                    totals=np.array([o["num_vertices"], o["num_indices"], o["num_points"], o["checksum"]], np.uint64))
for f in sorted(os.listdir(OUT)):
    print(f, os.path.getsize(os.path.join(OUT, f)))

# 5. final layout: the merge step over a stored batch, from the reference-side
#    wrapper (vgutil::batchTransformDrawIndices / memset32 are the reference's own functions)
from vg_renderer_b200.api import make_final_cfg  # noqa: E402

d = {"variant": np.uint32(ref.variant())}
t = BatchTables.cubic_strokes(W.config3_cubics(400, first=5), width=4.0, cap=1, join=1)
o = ref.run_batch(t)
FINAL_CASES = {"a": (65536, 3, 0, 0, 0, (127, 127)), "b": (4096, 1, 4000, 4000, 77, (5, -9)), "c": (1500, 0, 1499, 3, 1, (1, 1))}
for name, args in FINAL_CASES.items():
    idx, uv, cmds = ref.final_layout(o["mesh_vtx_off"], o["mesh_idx_off"], o["index"], make_final_cfg(*args))
    d[f"{name}_cfg"] = np.asarray(args[:5] + args[5], np.int64)
    d[f"{name}_index"] = idx
    d[f"{name}_cmds"] = cmds
np.savez_compressed(os.path.join(OUT, "final_layout.npz"), **d)
