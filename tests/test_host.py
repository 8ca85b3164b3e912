"""Host-side logic without a GPU: a record-only batch (device -1) builds the same sub-path
structure the reference's Path object produces, rejects what the reference cannot survive, and
every compute entry point fails loudly (no CPU fallback)."""
import numpy as np
import pytest

import vg_renderer_b200 as vgb
from oracle.pyref import OracleLib

from cases import mixed_recorder


def test_no_gpu_is_loud():
    if vgb.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(vgb.VgbError) as e:
        vgb.Batch(0)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)
    b = vgb.Batch(-1)
    b.begin_path().move_to(0, 0).line_to(1, 1)
    b.stroke()
    for fn in (b.upload, b.run, b.sync, b.sizes, b.download):
        with pytest.raises(vgb.VgbError) as e:
            fn()
        assert e.value.code == -4


def test_subpath_splitting_matches_reference_structure():
    r = mixed_recorder()
    t = r.tables()
    b = r.replay(vgb.Batch(-1))
    rec = b.recorded()
    orc = OracleLib()
    want = []  # sub-paths per path according to the algorithm restatement
    for p in range(t.num_paths):
        c0, c1 = t.path_cmd_off[p], t.path_cmd_off[p + 1]
        d0, d1 = t.path_data_off[p], t.path_data_off[p + 1]
        _, subs = orc.flatten(t.cmd[c0:c1], t.data[d0:d1], float(t.path_scale[p]), float(t.path_tol[p]))
        want.append(len(subs))
    got = np.bincount(rec["path_id"], minlength=t.num_paths).tolist()
    assert got == want
    assert rec["num_paths"] == t.num_paths and rec["num_draws"] == len(t.draw_op)
    # one job per (draw, sub-path), draws in order, sub-paths in order
    njobs = sum(int(t.draw_begin[p + 1] - t.draw_begin[p]) * want[p] for p in range(t.num_paths))
    assert rec["num_jobs"] == njobs
    assert np.all(np.diff(rec["job_draw"].astype(np.int64)) >= 0)
    # the no-op rect is dropped from the stream; every sub-path owns a contiguous command range
    assert np.all(rec["cmd_begin"][1:] == rec["cmd_end"][:-1])
    assert rec["cmd_end"][-1] == rec["num_cmds"]


def test_bulk_recorders_equal_per_call_recording():
    p8 = np.arange(40, dtype=np.float32).reshape(5, 8)
    a = vgb.Batch(-1)
    a.add_cubic_strokes(p8, width=3.0, cap=1, join=2)
    b = vgb.Batch(-1)
    for c in p8:
        b.begin_path().move_to(c[0], c[1]).cubic_to(*c[2:])
        b.stroke(0, 0xFF0000FF, 3.0, 1, 2)
    ra, rb = a.recorded(), b.recorded()
    for k in ("num_paths", "num_subpaths", "num_cmds", "num_data", "num_jobs"):
        assert ra[k] == rb[k], k
    for k in ("cmd_begin", "cmd_end", "data_begin", "path_id", "job_subpath"):
        np.testing.assert_array_equal(ra[k], rb[k])
    q = np.array([[0, 0, 50, 30, 5], [0, 0, 1e-9, 30, 0.01], [1, 1, 20, 20, 30]], np.float32)
    c = vgb.Batch(-1)
    c.add_rounded_rect_fills(q)
    rc = c.recorded()
    assert rc["num_paths"] == 3 and rc["num_subpaths"] == 2 and rc["num_jobs"] == 2  # middle rect is degenerate


def test_invalid_input_is_rejected():
    b = vgb.Batch(-1)
    b.begin_path()
    for call in (lambda: b.line_to(1, 1), lambda: b.cubic_to(1, 2, 3, 4, 5, 6), lambda: b.close()):
        with pytest.raises(vgb.VgbError) as e:
            call()
        assert e.value.code == -1  # moveTo must come first (path.cpp:82,88,709)
    b.move_to(0, 0)
    for bad in (float("nan"), float("inf"), -float("inf")):
        with pytest.raises(vgb.VgbError) as e:
            b.line_to(bad, 0)
        assert e.value.code == -2  # the reference never terminates on these
    with pytest.raises(vgb.VgbError):
        b.begin_path(0.0, 0.25)
    with pytest.raises(vgb.VgbError):
        b.begin_path(1.0, -1.0)
    b.begin_path().move_to(0, 0).line_to(5, 5)
    b.stroke()
    with pytest.raises(vgb.VgbError):
        b.line_to(9, 9)  # path already drawn
    with pytest.raises(vgb.VgbError):
        b.stroke(1)  # not a stroke op
    with pytest.raises(vgb.VgbError):
        b.stroke(0, 0xFFFFFFFF, 4.0, 7, 0)  # invalid cap (the reference silently does nothing)
    c = vgb.Batch(-1)
    with pytest.raises(vgb.VgbError):
        c.move_to(0, 0)  # outside begin_path
    c.begin_path().move_to(0, 0)
    for a0, a1 in ((1e9, 1.0), (0.5, -3e7)):
        with pytest.raises(vgb.VgbError) as e:
            c.arc(10, 10, 5, a0, a1, True)  # pathArc's `while (a > 2pi) a -= 2pi` never ends for such floats (path.cpp:635-645)
        assert e.value.code == -1
    c.arc(10, 10, 5, 1000.0, -999.0, False)  # large but harmless


def test_arc_subpath_rule_and_polyline():
    b = vgb.Batch(-1)
    b.begin_path().arc(0, 0, 10, 0, 1, 0)          # first command: pathArc -> moveTo
    b.arc(5, 5, 3, 0, 2, 1)                        # continues the same sub-path (lineTo)
    b.move_to(1, 1).polyline([[2, 2], [3, 3]])
    r = b.recorded()
    assert r["num_subpaths"] == 2 and r["num_cmds"] == 4
    assert r["num_data"] == 6 + 6 + 2 + (1 + 4)


def test_table_ingest_equals_per_call_recording():
    from vg_renderer_b200 import workloads as W
    t = mixed_recorder().tables()
    a = vgb.Batch(-1)
    a.add_tables(t)
    b = mixed_recorder().replay(vgb.Batch(-1))
    ra, rb = a.recorded(), b.recorded()
    for k in ("num_paths", "num_subpaths", "num_cmds", "num_data", "num_draws", "num_jobs"):
        assert ra[k] == rb[k], k
    for k in ("cmd_begin", "cmd_end", "data_begin", "path_id", "job_subpath", "job_draw"):
        np.testing.assert_array_equal(ra[k], rb[k])
    ti = W.instance_tables(W.tiger_like(40), 5)
    c = vgb.Batch(-1)
    c.add_tables(ti)
    rc = c.recorded()
    assert rc["num_paths"] == 200 and rc["num_draws"] == len(ti.draw_op)
    assert set(np.unique(ti.draw_op).tolist()) <= {0, 1, 3}
    assert np.all(ti.draw_width[ti.draw_op == 3] == 1.0)  # thin strokes carry the fringe as width


def test_bulk_zero_copy_recorded_counts():
    p8 = np.zeros((7, 8), np.float32)
    b = vgb.Batch(-1)
    b.set_cubic_strokes(p8)
    r = b.recorded()
    assert (r["num_paths"], r["num_subpaths"], r["num_jobs"], r["num_draws"]) == (7, 7, 7, 1)
    with pytest.raises(vgb.VgbError):
        b.begin_path()
    with pytest.raises(vgb.VgbError):
        b.add_cubic_strokes(p8)
    b.reset(1.0)
    b.begin_path().move_to(0, 0)


def test_scissor_table_and_marks():
    """vgb_batch_set_scissor: one table entry per distinct uint16-truncated rect (id 0 = never set), the force flag waits
    for a draw that has meshes, is dropped before the first draw, and reset() forgets everything"""
    b = vgb.Batch(-1)
    assert b.scissor_rects().tolist() == [[0, 0, 0, 0]]

    def draw(with_path=True):
        b.begin_path(1.0, 0.25)
        if with_path:
            b.move_to(0, 0)
            b.line_to(10, 0)
            b.line_to(10, 10)
        b.stroke(vgb.OP_STROKE_AA, 0xFFFFFFFF, 2.0, 0, 0)

    b.set_scissor([1.9, 2.2, 300.7, 200.0], True)      # before the first draw: nothing to break
    draw()
    b.set_scissor([1.1, 2.9, 300.2, 200.9], True)      # the same rect after truncation: same id, but the flag is up
    draw(with_path=False)                              # no sub-path, no mesh: the flag waits
    draw()
    b.set_scissor([5, 6, 7, 8], False)                 # another rect, no forced command (a caller that knows better)
    draw()
    b.set_scissor([1, 2, 300, 200], True)
    draw()
    ds = b.recorded_draw_scissor()
    np.testing.assert_array_equal(ds >> 31, [0, 0, 1, 0, 1])
    np.testing.assert_array_equal(ds & 0x7FFFFFFF, [1, 1, 1, 2, 1])
    assert b.scissor_rects().tolist() == [[0, 0, 0, 0], [1, 2, 300, 200], [5, 6, 7, 8]]
    np.testing.assert_array_equal(b.recorded_mesh_scissor() >> 31, [0, 1, 0, 1])   # one mesh per draw that has a sub-path
    with pytest.raises(vgb.VgbError):
        b.set_scissor([0, 0, float("nan"), 1], True)
    b.reset(1.0)
    assert b.scissor_rects().tolist() == [[0, 0, 0, 0]] and len(b.recorded_draw_scissor()) == 0
