"""The oracle itself (CPU only).  (1) the plain-C restatement (oracle/liboracle.so) against the
committed golden vectors produced by the compiled reference; (2) when oracle/_ref is present
(this container or a box that received the prebuilt .so), restatement vs compiled reference
directly, and the known-answer numbers of SURVEY.md App. C."""
import itertools
import os

import numpy as np
import pytest

from oracle.pyref import OracleLib, RefLib, have_ref
from vg_renderer_b200 import workloads as W
from vg_renderer_b200.tables import BatchTables

from cases import mixed_recorder, polyline_cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def orc():
    return OracleLib()


def same_batch(a, g, prefix):
    for k in ("pos", "color", "index", "mesh_vtx_off", "mesh_idx_off"):
        np.testing.assert_array_equal(a[k], g[f"{prefix}_{k}"], err_msg=f"{prefix} {k}")
    assert a["num_points"] == int(g[f"{prefix}_points"])


def test_golden_config1_kat(orc):
    g = np.load(os.path.join(GOLD, "config1_kat.npz"))
    assert int(g["variant"]) == 0
    pts, subs = orc.flatten([0, 2], W.config1_cubic()[0])
    assert len(pts) == 45 and subs == [(0, 45, False)]
    np.testing.assert_array_equal(pts, g["polyline"])
    np.testing.assert_allclose(pts[1], [104.6875, 95.674133], atol=1e-5)
    want = {(0, 0, 0): (180, 804), (0, 1, 1): (356, 1596), (0, 2, 2): (266, 1191)}
    for op in (0, 2, 3):
        for cap, join in itertools.product((0, 1, 2), (0, 1, 2)):
            pos, col, idx = orc.mesh(op, pts, False, 0xFF0000FF, 4.0, cap, join)
            k = f"op{op}_c{cap}_j{join}"
            np.testing.assert_array_equal(pos, g[k + "_pos"], err_msg=k)
            np.testing.assert_array_equal(idx, g[k + "_idx"], err_msg=k)
            if col is not None:
                np.testing.assert_array_equal(col, g[k + "_col"], err_msg=k)
            if (op, cap, join) in want:
                assert (len(pos), len(idx)) == want[(op, cap, join)]


def test_golden_polylines(orc):
    g = np.load(os.path.join(GOLD, "polylines.npz"))
    for ci, (name, c) in enumerate(polyline_cases(small=True)):
        np.testing.assert_array_equal(np.asarray(c, np.float32), g[f"case{ci}_pts"])
        for op, cap, join, closed in itertools.product((0, 2, 3), (0, 1, 2), (0, 1, 2), (0, 1)):
            if closed and len(c) < 3:
                continue
            pos, col, idx = orc.mesh(op, c, bool(closed), 0x80FF8040, 6.0, cap, join)
            k = f"case{ci}_op{op}_c{cap}_j{join}_z{closed}"
            np.testing.assert_array_equal(pos, g[k + "_pos"], err_msg=f"{name} {k}")
            np.testing.assert_array_equal(idx, g[k + "_idx"], err_msg=f"{name} {k}")
        if len(c) >= 3:
            for op in (1, 4):
                pos, col, idx = orc.mesh(op, c, True, 0xFF102030)
                np.testing.assert_array_equal(pos, g[f"case{ci}_op{op}_pos"])
                np.testing.assert_array_equal(idx, g[f"case{ci}_op{op}_idx"])


def test_golden_batches(orc):
    g = np.load(os.path.join(GOLD, "batches.npz"))
    for cap, join in itertools.product((0, 1, 2), (0, 1, 2)):
        t = BatchTables.cubic_strokes(W.config3_cubics(60, first=100 * (3 * cap + join)), width=4.0, cap=cap, join=join)
        same_batch(orc.run_batch(t), g, f"cubics_c{cap}_j{join}")
    for r in (2, 8, 32, 128):
        same_batch(orc.run_batch(BatchTables.rounded_rect_fills(W.config4_rrects(8, r))), g, f"rrect_r{r}")
    same_batch(orc.run_batch(mixed_recorder().tables()), g, "mixed")


def test_golden_stream_totals(orc):
    g = np.load(os.path.join(GOLD, "stream20k.npz"))
    o = orc.run_batch(BatchTables.cubic_strokes(W.config3_cubics(20000), width=4.0, cap=1, join=1), store=False)
    assert [o["num_vertices"], o["num_indices"], o["num_points"], o["checksum"]] == [int(v) for v in g["totals"]]


def test_workload_statistics(orc):
    """config 3 stream: ~23.5 points, ~184 vertices, ~821 indices per cubic"""
    n = 20000
    o = orc.run_batch(BatchTables.cubic_strokes(W.config3_cubics(n), width=4.0, cap=1, join=1), store=False)
    assert abs(o["num_points"] / n - 23.5) < 0.3
    assert abs(o["num_vertices"] / n - 184) < 3
    assert abs(o["num_indices"] / n - 821) < 12


def test_structural_properties(orc):
    """size-independent invariants of the mesh layout (also asserted on the GPU at full size)"""
    rng = np.random.default_rng(5)
    for n in (2, 3, 9, 100):
        p = (rng.random((n, 2)) * 300).astype(np.float32)
        for join in (0, 1, 2):
            pos, col, idx = orc.mesh(0, p, False, 0xFFFFFFFF, 5.0, 0, join)
            assert len(idx) % 3 == 0 and idx.max() < len(pos)
            if join == 0:
                assert len(pos) == 4 * n and len(idx) == 6 + 18 * (n - 2) + 24  # SURVEY A.3 totals
        if n >= 3:
            pos, col, idx = orc.mesh(1, p, True, 0xFFFFFFFF)
            assert len(pos) == 2 * n and len(idx) == 9 * n - 6
            assert set(np.unique(col).tolist()) <= {0xFFFFFFFF, 0x00FFFFFF}


needs_ref = pytest.mark.skipif(not have_ref("scalar"), reason="oracle/_ref not built")


@needs_ref
def test_restatement_vs_compiled_reference_meshes(orc):
    ref = RefLib("scalar")
    assert ref.variant() == 0
    cases = polyline_cases()
    for op, cap, join, closed in itertools.product((0, 2, 3), (0, 1, 2), (0, 1, 2), (False, True)):
        for w in (1.0, 3.5, 40.0):
            for name, c in cases[:14]:
                if closed and len(c) < 3:
                    continue
                a = ref.mesh(op, c, closed, 0x80FF8040, w, cap, join)
                b = orc.mesh(op, c, closed, 0x80FF8040, w, cap, join)
                np.testing.assert_array_equal(a[0], b[0], err_msg=f"{name} {op} {cap} {join} {closed} {w}")
                np.testing.assert_array_equal(a[2], b[2])
                if a[1] is not None:
                    np.testing.assert_array_equal(a[1], b[1])


@needs_ref
def test_restatement_vs_compiled_reference_batches(orc):
    ref = RefLib("scalar")
    for cap, join in itertools.product((0, 1, 2), (0, 1, 2)):
        t = BatchTables.cubic_strokes(W.config3_cubics(1500, first=31 * cap + join), width=4.0, cap=cap, join=join)
        a, b = ref.run_batch(t), orc.run_batch(t)
        for k in ("pos", "color", "index", "mesh_vtx_off", "mesh_idx_off"):
            np.testing.assert_array_equal(a[k], b[k])
    t = mixed_recorder().tables()
    a, b = ref.run_batch(t), orc.run_batch(t)
    for k in ("pos", "color", "index", "mesh_vtx_off", "mesh_idx_off"):
        np.testing.assert_array_equal(a[k], b[k])
    for tol in (0.25, 1e-2, 1e-4):
        cmd, data = [0, 2], [0, 0, 30000, -20000, -10000, 40000, 20000, 1000]
        pa, sa = ref.flatten(cmd, data, 1.0, tol)
        pb, sb = orc.flatten(cmd, data, 1.0, tol)
        assert sa == sb and len(pa) == {0.25: 376, 1e-2: 900, 1e-4: 2791}[tol]
        np.testing.assert_array_equal(pa, pb)


@needs_ref
def test_sse_build_same_triangles_different_order():
    """the stock x86 build (VG_CONFIG_ENABLE_SIMD=1) emits the same convex-fill triangle SET in a
    different ORDER and with approximate rsqrt/rcp positions: documented, not a parity target."""
    if not have_ref("sse"):
        pytest.skip("sse build missing")
    ref, sse = RefLib("scalar"), RefLib("sse")
    pts, _ = ref.flatten([6], [10, 10, 200, 100, 20])
    a, b = ref.mesh(1, pts, True, 0xFF00FF00), sse.mesh(1, pts, True, 0xFF00FF00)
    assert len(a[2]) == len(b[2]) == 210 and not np.array_equal(a[2], b[2])
    tri = lambda i: sorted(map(tuple, i.reshape(-1, 3).tolist()))  # noqa: E731
    assert tri(a[2]) == tri(b[2])
    assert np.abs(a[0] - b[0]).max() < 5e-3
    # the stroke path is the same code in both builds
    pl, _ = ref.flatten([0, 2], W.config1_cubic()[0])
    x, y = ref.mesh(0, pl, False, 0xFF0000FF, 4.0, 1, 1), sse.mesh(0, pl, False, 0xFF0000FF, 4.0, 1, 1)
    np.testing.assert_array_equal(x[0], y[0])
    np.testing.assert_array_equal(x[2], y[2])


@needs_ref
def test_libm_shim_sensitivity():
    """how much the choice of math library moves the result (bx is unpinned): counts agree on
    the cubic stream; where counts agree positions agree to 1e-4."""
    if not have_ref("libm"):
        pytest.skip("libm build missing")
    ref, lm = RefLib("scalar"), RefLib("libm")
    t = BatchTables.cubic_strokes(W.config3_cubics(3000), width=4.0, cap=1, join=1)
    a, b = ref.run_batch(t), lm.run_batch(t)
    nva, nvb = np.diff(a["mesh_vtx_off"]), np.diff(b["mesh_vtx_off"])
    agree = nva == nvb
    assert agree.mean() > 0.995
    if agree.all():
        np.testing.assert_array_equal(a["index"], b["index"])
        # 1e-4 abs, or 2 ulp where the coordinate's own float spacing exceeds that (|x| >= 1024)
        tol = np.maximum(1e-4, 2 * np.spacing(np.abs(a["pos"])))
        assert np.all(np.abs(a["pos"] - b["pos"]) <= tol)


# ---- final layout ------------------------------------------------------------------
def _final_cfgs():
    from vg_renderer_b200.api import make_final_cfg
    return [make_final_cfg(65536), make_final_cfg(65536, 4, 65000, 100, 12, (7, 8)), make_final_cfg(4096, 0, 4096, 4096, 0, (-1, 1)),
            make_final_cfg(1500, 2, 1499, 0, 3, (127, 127)), make_final_cfg(1000, 9, 0, 0, 1 << 20, (0, 0))]


def test_golden_final_layout(orc):
    from vg_renderer_b200.api import make_final_cfg
    g = np.load(os.path.join(GOLD, "final_layout.npz"))
    o = orc.run_batch(BatchTables.cubic_strokes(W.config3_cubics(400, first=5), width=4.0, cap=1, join=1))
    for name in "abc":
        c = [int(v) for v in g[name + "_cfg"]]
        idx, uv, cmds = orc.final_layout(o["mesh_vtx_off"], o["mesh_idx_off"], o["index"], make_final_cfg(*c[:5], uv=(c[5], c[6])))
        np.testing.assert_array_equal(idx, g[name + "_index"])
        assert cmds.tobytes() == g[name + "_cmds"].tobytes()
        assert np.all(uv == np.asarray(c[5:7], np.int16))


def test_final_layout_properties(orc):
    """what the merge contract guarantees, whatever the sizes: every command's indices address its own
    vertices only, commands tile the streams, no buffer holds more than max_vb_vertices"""
    o = orc.run_batch(mixed_recorder().tables())
    for cfg in _final_cfgs():
        if int(np.diff(o["mesh_vtx_off"]).max()) >= cfg.max_vb_vertices:
            with pytest.raises(RuntimeError):
                orc.final_layout(o["mesh_vtx_off"], o["mesh_idx_off"], o["index"], cfg)
            continue
        idx, uv, cmds = orc.final_layout(o["mesh_vtx_off"], o["mesh_idx_off"], o["index"], cfg)
        assert int(cmds["num_vertices"].sum()) == o["num_vertices"] and int(cmds["num_indices"].sum()) == o["num_indices"]
        for k, c in enumerate(cmds):
            lo = cfg.cmd_vertices if (k == 0 and c["first_vertex"] == cfg.vb_count and cfg.cmd_vertices) else 0
            part = idx[int(c["stream_index"]): int(c["stream_index"]) + int(c["num_indices"])]
            assert part.min() >= lo and part.max() < lo + c["num_vertices"]
            assert c["first_vertex"] + c["num_vertices"] <= cfg.max_vb_vertices
        assert np.all(np.diff(cmds["vb_id"].astype(np.int64)) == 1)


@needs_ref
def test_final_layout_restatement_vs_reference_functions(orc):
    ref = RefLib("scalar")
    for t in (BatchTables.cubic_strokes(W.config3_cubics(1200, first=9), width=4.0, cap=1, join=1), mixed_recorder().tables()):
        o = ref.run_batch(t)
        for cfg in _final_cfgs():
            try:
                a = ref.final_layout(o["mesh_vtx_off"], o["mesh_idx_off"], o["index"], cfg)
            except RuntimeError:
                with pytest.raises(RuntimeError):
                    orc.final_layout(o["mesh_vtx_off"], o["mesh_idx_off"], o["index"], cfg)
                continue
            b = orc.final_layout(o["mesh_vtx_off"], o["mesh_idx_off"], o["index"], cfg)
            np.testing.assert_array_equal(a[0], b[0])
            np.testing.assert_array_equal(a[1], b[1])
            assert a[2].tobytes() == b[2].tobytes()


@needs_ref
def test_transform_positions_restatement_vs_reference(orc):
    ref = RefLib("scalar")
    rng = np.random.default_rng(3)
    pos = rng.uniform(-4096, 4096, (10001, 2)).astype(np.float32)
    for k in range(8):
        m = rng.uniform(-3, 3, 6).astype(np.float32)
        np.testing.assert_array_equal(ref.transform_positions(pos, m), orc.transform_positions(pos, m))
    ident = np.asarray([1, 0, 0, 1, 0, 0], np.float32)
    np.testing.assert_array_equal(orc.transform_positions(pos, ident), pos)


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_fuzz_restatement_vs_compiled_reference(orc, seed):
    from cases import fuzz_recorder
    t = fuzz_recorder(seed).tables()
    a, b = RefLib("scalar").run_batch(t), orc.run_batch(t)
    assert a["num_vertices"] > 20000
    for k in ("mesh_vtx_off", "mesh_idx_off", "index", "color", "pos"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


@needs_ref
def test_degenerate_width_below_fringe_counts(orc):
    """AA stroke narrower than the fringe (vg.cpp never issues it, the per-object API can): hsw <= 0 drives
    acos out of range, da becomes 0 or 2*pi and pi/da infinite.  `(uint32_t)inf` is undefined in C; the
    reference's x86-64 build yields 0 (-> 2 cap points), a GPU conversion would saturate to 2^32-1.  All
    float -> count conversions go through vgb_f2u (x86 semantics) on both sides: same meshes."""
    ref = RefLib("scalar")
    pts = [[0, 0], [40, 5], [60, 50], [10, 70], [5, 30]]
    for scale, tol, fringe, width in ((2.0, 0.25, 2.0, 1.5), (1.0, 0.25, 2.0, 1.5), (1.0, 0.25, 2.0, 2.0), (3.7, 0.1, 2.0, 1.5), (1.0, 1.0, 0.5, 0.25)):
        for cap, join, closed in itertools.product((0, 1, 2), (0, 1, 2), (False, True)):
            a = ref.mesh(0, pts, closed, 0xFF204060, width, cap, join, scale, tol, fringe)
            b = orc.mesh(0, pts, closed, 0xFF204060, width, cap, join, scale, tol, fringe)
            assert len(a[0]) < 400
            np.testing.assert_array_equal(a[0], b[0])
            np.testing.assert_array_equal(a[2], b[2])
