"""How much of the parity claim rests on the shared fp32 math (vgb_math.h)?  The reference takes sin/cos/acos/atan2 from
bkaradzic/bx, which is neither vendored nor pinned; this repo defines them itself ("parity unpinned" w.r.t.
bx).  This script runs the SAME reference sources (path.cpp + stroker.cpp, compiled by oracle/Makefile) against the two
shims -- vgb_math.h (golden, libvgref.so) and libm (libvgref_libm.so) -- over the full benchmark stream (1 M cubics,
BASELINE configs[2]) and the rounded-rect sweep (configs[3]) and reports how often the integer decisions (vertex /
index counts per mesh) differ and how far positions move.  CPU only.  Writes profiles/r02_libm_sensitivity.json.
   python tools/libm_sensitivity.py [curves]"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.pyref import RefLib  # noqa: E402
from vg_renderer_b200 import workloads as W  # noqa: E402
from vg_renderer_b200.tables import BatchTables  # noqa: E402


def compare(t):
    a, b = RefLib("scalar").run_batch(t), RefLib("libm").run_batch(t)
    va, vb = np.diff(a["mesh_vtx_off"].astype(np.int64)), np.diff(b["mesh_vtx_off"].astype(np.int64))
    ia, ib = np.diff(a["mesh_idx_off"].astype(np.int64)), np.diff(b["mesh_idx_off"].astype(np.int64))
    diff = (va != vb) | (ia != ib)
    same = ~diff
    dpos, dmax_ulps = 0.0, 0
    pts_diff = int(a["num_points"] != b["num_points"])
    if same.any():
        # positions of the meshes whose counts agree (they are then comparable vertex by vertex)
        sa = np.repeat(same, va)
        sb = np.repeat(same, vb)
        pa, pb = a["pos"][sa], b["pos"][sb]
        dpos = float(np.max(np.abs(pa.astype(np.float64) - pb), initial=0.0))
        idx_same = bool(np.array_equal(a["index"][np.repeat(same, ia)], b["index"][np.repeat(same, ib)]))
    else:
        idx_same = True
    return dict(meshes=int(len(va)), meshes_with_different_counts=int(diff.sum()), vertex_count_delta=int(np.abs(va - vb).sum()),
                max_abs_position_delta=dpos, indices_equal_where_counts_agree=idx_same, points_total_differs=pts_diff,
                points=(a["num_points"], b["num_points"]))


def cubic_chunk(args):
    first, count = args
    return compare(BatchTables.cubic_strokes(W.config3_cubics(count, first=first), width=4.0, cap=1, join=1))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    per = 25_000
    with mp.get_context("fork").Pool(min(8, mp.cpu_count())) as pool:
        res = pool.map(cubic_chunk, [(k, min(per, n - k)) for k in range(0, n, per)], chunksize=1)
    cub = dict(curves=n, meshes=sum(r["meshes"] for r in res), meshes_with_different_counts=sum(r["meshes_with_different_counts"] for r in res),
               vertex_count_delta=sum(r["vertex_count_delta"] for r in res), max_abs_position_delta=max(r["max_abs_position_delta"] for r in res),
               indices_equal_where_counts_agree=all(r["indices_equal_where_counts_agree"] for r in res),
               chunks_with_different_point_totals=sum(r["points_total_differs"] for r in res))
    cub["fraction_of_meshes_with_different_counts"] = cub["meshes_with_different_counts"] / max(cub["meshes"], 1)
    rows = []
    for r in (2, 8, 32, 128, 512, 2048):
        # the sweep's rects differ only in x (10 + i % 7): 7 distinct shapes per radius
        c = compare(BatchTables.rounded_rect_fills(W.config4_rrects(7, r)))
        c["r"] = r
        c["polygon_points"] = (c["points"][0] // 7, c["points"][1] // 7)
        rows.append(c)
    out = dict(what="reference sources against vgb_math.h (golden) vs against libm: count / position sensitivity to the math library",
               cubics_aa_round_round=cub, rounded_rect_sweep=rows)
    os.makedirs("profiles", exist_ok=True)
    json.dump(out, open("profiles/r02_libm_sensitivity.json", "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
