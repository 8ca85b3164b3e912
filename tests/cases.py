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


def fuzz_recorder(seed, num_paths=120):
    """random paths (polylines with duplicates / collinear runs / hairpins, curves, shapes) x random draws
    (every op, cap, join, width incl. thin, scale / tolerance / transform): a broad net for the general kernels
    and for the oracle restatement; deterministic per seed"""
    rng = np.random.default_rng(seed)
    r = Recorder(fringe=float(rng.choice([1.0, 1.0, 0.5, 2.0])))
    for _ in range(num_paths):
        scale = float(rng.choice([1.0, 1.0, 0.5, 2.0, 3.7]))
        tol = float(rng.choice([0.25, 0.25, 0.1, 1.0]))
        xf = None
        if rng.random() < 0.3:
            a, s = rng.uniform(0, 6.28), rng.uniform(0.5, 2.0)
            xf = [float(v) for v in np.asarray([np.cos(a) * s, np.sin(a) * s, -np.sin(a) * s, np.cos(a) * s, rng.uniform(-50, 50), rng.uniform(-50, 50)], np.float32)]
        r.begin_path(scale, tol, xf)
        kind = rng.integers(0, 7)
        fillable = False
        if kind <= 2:  # polyline with degenerate features
            n = int(rng.integers(2, 40))
            pts = rng.uniform(0, 300, (n, 2)).astype(np.float32)
            for k in range(1, n):
                u = rng.random()
                if u < 0.08:
                    pts[k] = pts[k - 1]  # exact duplicate (deduped by the path)
                elif u < 0.14:
                    pts[k] = pts[k - 1] + np.float32(rng.uniform(-2e-3, 2e-3, 2))  # near duplicate around the epsilon
                elif u < 0.22 and k >= 2:
                    pts[k] = pts[k - 1] + (pts[k - 1] - pts[k - 2]) * np.float32(rng.uniform(0.2, 2))  # collinear continuation
                elif u < 0.28 and k >= 2:
                    pts[k] = pts[k - 2] + np.float32(rng.uniform(-0.3, 0.3, 2))  # hairpin
            r.move_to(float(pts[0, 0]), float(pts[0, 1]))
            for k in range(1, n):
                r.line_to(float(pts[k, 0]), float(pts[k, 1]))
            if rng.random() < 0.4:
                r.close()
        elif kind == 3:  # curves
            p = rng.uniform(0, 300, 2)
            r.move_to(*map(float, p))
            for _ in range(int(rng.integers(1, 5))):
                c = rng.uniform(-120, 120, (3, 2)) + p
                if rng.random() < 0.3:
                    r.quadratic_to(*map(float, c[0]), *map(float, c[2]))
                else:
                    r.cubic_to(*map(float, c.reshape(-1)))
                p = c[2]
            if rng.random() < 0.3:
                r.close()
        elif kind == 4:  # convex shapes
            fillable = True
            u = rng.random()
            if u < 0.3:
                r.circle(float(rng.uniform(0, 300)), float(rng.uniform(0, 300)), float(rng.uniform(0.5, 120)))
            elif u < 0.5:
                r.ellipse(float(rng.uniform(0, 300)), float(rng.uniform(0, 300)), float(rng.uniform(1, 90)), float(rng.uniform(1, 90)))
            elif u < 0.8:
                r.rounded_rect(float(rng.uniform(0, 100)), float(rng.uniform(0, 100)), float(rng.uniform(5, 200)), float(rng.uniform(5, 200)), float(rng.uniform(0, 60)))
            else:
                r.rect(float(rng.uniform(0, 100)), float(rng.uniform(0, 100)), float(rng.uniform(1, 200)), float(rng.uniform(1, 200)))
        elif kind == 5:  # arcs
            r.move_to(float(rng.uniform(0, 300)), float(rng.uniform(0, 300)))
            r.arc(float(rng.uniform(0, 300)), float(rng.uniform(0, 300)), float(rng.uniform(2, 100)), float(rng.uniform(0, 6.2)), float(rng.uniform(0, 6.2)), bool(rng.integers(0, 2)))
            r.arc_to(float(rng.uniform(0, 300)), float(rng.uniform(0, 300)), float(rng.uniform(0, 300)), float(rng.uniform(0, 300)), float(rng.uniform(1, 40)))
        else:  # several sub-paths in one path, some too short to draw
            for _ in range(int(rng.integers(1, 4))):
                r.move_to(float(rng.uniform(0, 300)), float(rng.uniform(0, 300)))
                for _ in range(int(rng.integers(0, 4))):
                    r.line_to(float(rng.uniform(0, 300)), float(rng.uniform(0, 300)))
        for _ in range(int(rng.integers(1, 4))):
            op = int(rng.choice([0, 0, 0, 2, 3, 1, 4])) if fillable else int(rng.choice([0, 0, 0, 2, 3]))
            color = int(rng.integers(0, 1 << 32))
            if op in (1, 4):
                r.convex_fill(color, op == 1)
            else:
                width = float(r.fringe) if op == 3 else float(rng.choice([1.5, 2.0, 4.0, 7.5, 20.0, 60.0]))
                r.stroke(op, color, width, int(rng.integers(0, 3)), int(rng.integers(0, 3)))
    return r
