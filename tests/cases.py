"""Shared test inputs (used by the GPU parity tests, the CPU oracle tests and tools/make_golden.py)."""
import numpy as np

from vg_renderer_b200.tables import Recorder


def polyline_cases(small=False):
    rng = np.random.default_rng(1234)
    cases = []
    cases.append(("two-point", [[0, 0], [10, 0]]))
    cases.append(("collinear", [[0, 0], [10, 0], [20, 0], [30, 0]]))
    cases.append(("hairpin", [[0, 0], [50, 0], [0, 0.5]]))
    cases.append(("right-angle-l", [[0, 0], [40, 0], [40, 40]]))
    cases.append(("right-angle-r", [[0, 0], [40, 0], [40, -40]]))
    cases.append(("duplicate", [[0, 0], [10, 10], [10, 10], [20, 0]]))
    cases.append(("near-duplicate", [[0, 0], [10, 10], [10.001, 10.001], [20, 0]]))
    cases.append(("zigzag", [[i * 7.0, (i % 2) * 9.0] for i in range(40)]))
    cases.append(("square", [[0, 0], [0, 30], [30, 30], [30, 0]]))
    cases.append(("triangle-cw", [[0, 0], [30, 0], [15, 25]]))
    cases.append(("star", [[50 + (40 if i % 2 else 15) * np.cos(i * np.pi / 5), 50 + (40 if i % 2 else 15) * np.sin(i * np.pi / 5)] for i in range(10)]))
    for k in range(2 if small else 6):
        n = int(rng.integers(3, 60 if small else 300))
        cases.append((f"random{k}", (rng.random((n, 2)) * 200).tolist()))
    if not small:
        cases.append(("long", (np.cumsum(rng.random((3000, 2)) * 10 - 3, axis=0)).tolist()))
    return [(n, np.asarray(c, np.float32).tolist()) for n, c in cases]


def mixed_recorder():
    """every path builder + transforms + fill-then-stroke on one path"""
    r = Recorder(fringe=1.0)
    r.begin_path().move_to(10, 10).line_to(100, 20).cubic_to(150, 80, 60, 160, 20, 120).quadratic_to(0, 60, 10, 10).close()
    r.convex_fill(0xFF112233, True).stroke(0, 0xFF445566, 3.0, 1, 1)
    r.begin_path(2.0, 0.125, [2, 0, 0, 2, 5, 7]).rounded_rect(0, 0, 120, 60, 12).stroke(0, 0xFFAA0000, 5.0, 0, 0)
    r.begin_path().rounded_rect(5, 5, 40, 40, 30).convex_fill(0xFF00AA00, True)  # becomes a circle
    r.begin_path().rounded_rect(5, 5, 40, 20, 0.05).stroke(0, 0xFF0000AA, 2.0, 2, 2)  # becomes a rect
    r.begin_path().rounded_rect_varying(0, 0, 200, 100, 30, 0.01, 10, 50).convex_fill(0xFF778899, True).stroke(0, 0xFF000000, 2.5, 0, 1)
    r.begin_path().circle(300, 300, 75).stroke(0, 0xFF123456, 8.0, 0, 1)
    r.begin_path().ellipse(300, 100, 90, 30).convex_fill(0xFF654321, True)
    r.begin_path().move_to(400, 400).arc(400, 300, 100, 0.3, 2.5, False).stroke(0, 0xFFFFFFFF, 4.0, 1, 1)
    r.begin_path().arc(100, 300, 50, 5.5, 1.0, True).line_to(100, 300).close().stroke(0, 0xFF00FFFF, 3.0, 0, 2)
    r.begin_path().move_to(20, 200).line_to(120, 200).arc_to(170, 200, 170, 260, 30).line_to(170, 300).stroke(0, 0xFFFF00FF, 6.0, 2, 1)
    r.begin_path().move_to(0, 0).polyline([[5, 5], [10, 0], [15, 5], [20, 0]]).stroke(0, 0xFF0F0F0F, 2.0, 1, 0)
    r.begin_path().move_to(0, 0).line_to(50, 0).move_to(0, 20).line_to(50, 20).line_to(50, 60).move_to(7, 7)
    r.stroke(0, 0xFFABCDEF, 4.0, 1, 1)  # three sub-paths, the last one has a single vertex (skipped)
    r.begin_path().rect(0, 0, 1e-7, 50).rect(10, 10, 30, 20).convex_fill(0xFF010203, True)  # first rect is a no-op
    r.begin_path(1.0, 0.25, [0.8, 0.6, -0.6, 0.8, 100, 50]).move_to(0, 0).cubic_to(300, -200, -100, 400, 200, 10).stroke(0, 0xFF333333, 10.0, 1, 1)
    return r
