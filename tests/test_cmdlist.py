"""Command-list front end: decoding vg-renderer's CommandList bytes into a batch.  CPU
tests compare what the product records from the bytes with what an independent float32 model of the
reference's executor (tests/cmdlist.py) records through explicit calls; the GPU test runs the decoded
batch and compares the meshes with the compiled reference run on the model's tables."""
import ctypes as C
import struct

import numpy as np
import pytest

import vg_renderer_b200 as vgb
from oracle.pyref import OracleLib

import cmdlist as CL


@pytest.fixture(scope="module")
def sincos():
    lib = OracleLib().lib
    lib.vgo_math_eval.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]

    def f(a):
        x = np.asarray([a], np.float32)
        s, c = np.empty(1, np.float32), np.empty(1, np.float32)
        lib.vgo_math_eval(0, x.ctypes.data, x.ctypes.data, s.ctypes.data, 1)
        lib.vgo_math_eval(1, x.ctypes.data, x.ctypes.data, c.ctypes.data, 1)
        return s[0], c[0]
    return f


KEYS = ("cmd", "data", "path_scale", "path_tol", "path_xform", "draw_op", "draw_color", "draw_width", "draw_cap", "draw_join", "draw_path",
        "cmd_begin", "cmd_end", "data_begin", "path_id", "job_subpath", "job_draw")


def recorded_equal(a, b):
    for k in KEYS:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


@pytest.mark.parametrize("cull", [False, True])
@pytest.mark.parametrize("xform,alpha", [((1, 0, 0, 1, 0, 0), 1.0), ((1.25, 0.5, -0.5, 1.25, 40, -10), 0.8)])
def test_decoded_list_equals_model_recording(sincos, cull, xform, alpha):
    w = CL.demo_scene()
    model = CL.Model(sincos, xform=xform, global_alpha=alpha, tol=0.25, fringe=1.0, cull=cull).run(w.ops)
    expect = vgb.Batch(-1)
    model.rec.replay(expect)
    got = vgb.Batch(-1)
    got.reset(1.0)
    ctx = vgb.make_cl_context(xform=xform, global_alpha=alpha, tess_tolerance=0.25, flags=2 if cull else 0)
    st = got.submit_command_list(w.bytes(), ctx)
    recorded_equal(got.recorded_streams(), expect.recorded_streams())
    for k, v in model.stats.items():
        assert getattr(st, k) == v, k
    assert st.commands == len(w.ops) and st.state_depth == 0
    assert (st.culled > 0) == cull and st.zero_alpha == 1 and st.unsupported >= 5
    # the list's state changes persist (VG_CONFIG_COMMAND_LIST_PRESERVE_STATE = 0)
    np.testing.assert_array_equal(np.asarray(list(ctx.xform), np.float32), np.asarray(model.s["m"], np.float32))
    assert ctx.global_alpha == np.float32(model.s["alpha"])
    np.testing.assert_array_equal(np.asarray(list(ctx.scissor), np.float32), np.asarray(model.s["sc"], np.float32))


def test_two_lists_accumulate_and_state_carries_over(sincos):
    a = CL.Writer().translate(10, 20).scale(3, 3).begin_path().circle(0, 0, 10).fill(0xFF010101, CL.fill_flags())
    b = CL.Writer().begin_path().move_to(0, 0).line_to(9, 9).stroke(0xFF020202, 1.0, CL.stroke_flags(0, 0, True))
    model = CL.Model(sincos).run(a.ops).run(b.ops)
    expect = vgb.Batch(-1)
    model.rec.replay(expect)
    got = vgb.Batch(-1)
    ctx = vgb.make_cl_context()
    got.submit_command_list(a.bytes(), ctx)
    got.submit_command_list(b.bytes(), ctx)
    recorded_equal(got.recorded_streams(), expect.recorded_streams())
    r = got.recorded_streams()
    assert r["path_scale"][1] == np.float32(3.0) and r["draw_width"][1] == np.float32(3.0) and r["draw_op"][1] == 0


def test_malformed_lists_are_rejected():
    b = vgb.Batch(-1)
    ctx = vgb.make_cl_context()
    good = CL.Writer().begin_path().move_to(1, 2).bytes()
    for bad in (good[:-4], good[:20], struct.pack("<II", 99, 0) + b"\0" * 8, struct.pack("<II", CL.MOVE_TO, 0) + b"\0" * 8,
                struct.pack("<II", CL.MOVE_TO, 64) + b"\0" * 24, struct.pack("<II", CL.POLYLINE, 16) + b"\0" * 8 + struct.pack("<I", 1000) + b"\0" * 12,
                CL.Writer().pop_state().bytes(), CL.Writer().move_to(1, 2).bytes()):
        with pytest.raises(vgb.VgbError):
            b.reset(1.0)
            b.submit_command_list(bad, ctx)
    w = CL.Writer()
    for _ in range(40):
        w.push_state()
    with pytest.raises(vgb.VgbError):
        b.submit_command_list(w.bytes(), ctx)
    with pytest.raises(vgb.VgbError):  # non-finite coordinates are refused at record time like everywhere else
        b.reset(1.0)
        b.submit_command_list(CL.Writer().begin_path().move_to(float("nan"), 0).bytes(), ctx)
    b.reset(1.0)
    assert b.submit_command_list(b"", ctx).commands == 0


@pytest.mark.gpu
def test_decoded_list_meshes_match_reference(sincos, gpu_batch, ref):
    w = CL.demo_scene()
    xform = (1.25, 0.5, -0.5, 1.25, 40, -10)
    model = CL.Model(sincos, xform=xform, global_alpha=0.9).run(w.ops)
    exp = ref.run_batch(model.rec.tables())
    gpu_batch.reset(1.0)
    st = gpu_batch.submit_command_list(w.bytes(), vgb.make_cl_context(xform=xform, global_alpha=0.9))
    assert st.draws == model.stats["draws"]
    gpu_batch.run()
    out = gpu_batch.download()
    for k in ("mesh_vtx_off", "mesh_idx_off", "index", "color", "pos"):
        np.testing.assert_array_equal(out[k], exp[k], err_msg=k)


def _random_list(seed, n_ops=400, concave=True):
    """random but well-formed command stream: every path command sits inside a BeginPath, pops never underflow"""
    rng = np.random.default_rng(seed)
    w = CL.Writer()
    depth, in_path, has_pt, drawn = 0, False, False, False
    for _ in range(n_ops):
        u = rng.random()
        if u < 0.12 or not in_path or (drawn and u < 0.52):  # geometry after a draw needs a new path (VG_CHECK, vg.cpp:2986)
            w.begin_path()
            in_path, has_pt, drawn = True, False, False
        elif u < 0.45:
            k = int(rng.integers(0, 9))
            f = lambda lo=-200, hi=400: float(np.float32(rng.uniform(lo, hi)))
            if k == 0 or not has_pt:
                w.move_to(f(), f())
            elif k == 1:
                w.line_to(f(), f())
            elif k == 2:
                w.cubic_to(f(), f(), f(), f(), f(), f())
            elif k == 3:
                w.quadratic_to(f(), f(), f(), f())
            elif k == 4:
                w.arc_to(f(), f(), f(), f(), f(1, 40))
            elif k == 5:
                w.arc(f(), f(), f(2, 90), f(0, 6), f(0, 6), bool(rng.integers(0, 2)))
            elif k == 6:
                w.polyline(rng.uniform(0, 300, 2 * int(rng.integers(1, 6))))
            elif k == 7:
                w.close_path()
            else:
                w.line_to(f(), f())
            has_pt = True
        elif u < 0.52:
            k = int(rng.integers(0, 5))
            f = lambda lo=0, hi=300: float(np.float32(rng.uniform(lo, hi)))
            [lambda: w.rect(f(), f(), f(1, 200), f(1, 200)), lambda: w.rounded_rect(f(), f(), f(5, 200), f(5, 200), f(0, 40)),
             lambda: w.circle(f(), f(), f(1, 90)), lambda: w.ellipse(f(), f(), f(1, 90), f(1, 90)),
             lambda: w.rounded_rect_varying(f(), f(), f(20, 200), f(20, 200), f(0, 30), f(0, 30), f(0, 30), f(0, 30))][k]()
            has_pt = True
        elif u < 0.66:
            drawn = True
            w.stroke(int(rng.integers(0, 1 << 32)), float(np.float32(rng.choice([0.1, 0.5, 1.0, 2.5, 8.0, 300.0]))),
                     CL.stroke_flags(int(rng.integers(0, 3)), int(rng.integers(0, 3)), bool(rng.integers(0, 2)), bool(rng.random() < 0.2)))
        elif u < 0.74:
            drawn = True
            w.fill(int(rng.integers(0, 1 << 32)), CL.fill_flags(concave=bool(rng.random() < 0.15) and concave, aa=bool(rng.integers(0, 2))))
        elif u < 0.80:
            if depth < 20:
                w.push_state()
                depth += 1
        elif u < 0.85:
            if depth > 0:
                w.pop_state()
                depth -= 1
        elif u < 0.97:
            k = int(rng.integers(0, 8))
            f = lambda lo, hi: float(np.float32(rng.uniform(lo, hi)))
            [lambda: w.translate(f(-50, 50), f(-50, 50)), lambda: w.scale(f(0.3, 2.5), f(0.3, 2.5)), lambda: w.rotate(f(-7, 7)),
             lambda: w.mult([f(-1.5, 1.5) for _ in range(4)] + [f(-30, 30), f(-30, 30)], bool(rng.integers(0, 2))),
             lambda: w.global_alpha(f(0, 1)), lambda: w.set_scissor(f(-100, 600), f(-100, 400), f(0, 3), f(0, 300)),
             lambda: w.intersect_scissor(f(-100, 600), f(-100, 400), f(0, 500), f(0, 300)), lambda: w.reset_scissor()][k]()
        else:
            w.identity() if rng.random() < 0.5 else w.set_view_box(0, 0, float(np.float32(rng.uniform(100, 2000))), float(np.float32(rng.uniform(100, 2000))))
    return w


@pytest.mark.parametrize("seed", range(8))
def test_random_lists_equal_model_recording(sincos, seed):
    w = _random_list(seed)
    cull = bool(seed % 2)
    model = CL.Model(sincos, fringe=1.0, cull=cull).run(w.ops)
    expect = vgb.Batch(-1)
    model.rec.replay(expect)
    got = vgb.Batch(-1)
    ctx = vgb.make_cl_context(flags=2 if cull else 0)
    st = got.submit_command_list(w.bytes(), ctx)
    recorded_equal(got.recorded_streams(), expect.recorded_streams())
    for k, v in model.stats.items():
        assert getattr(st, k) == v, k
    assert st.state_depth == len(model.stack) - 1


def test_nested_lists(sincos):
    icon = CL.Writer().push_state().scale(0.5, 0.5).begin_path().rounded_rect(0, 0, 40, 40, 6).fill(0xFF336699, CL.fill_flags())
    icon.stroke(0xFF000000, 2.0, CL.stroke_flags(1, 1, True)).pop_state().translate(50, 0)  # leaves a translation behind
    row = CL.Writer().submit(3).submit(3).submit(9).rotate(0.2).submit(3)  # 9: not a list of the context -> skipped
    deep = CL.Writer().translate(1, 1).begin_path().circle(0, 0, 3).fill(0xFFFFFFFF, CL.fill_flags(aa=False)).submit(7)  # submits itself
    culled = CL.Writer().set_scissor(0, 0, 0.2, 0.2).begin_path().rect(0, 0, 5, 5).fill(0xFF00FF00, CL.fill_flags())  # AllowCommandCulling
    lists = {3: (icon, 0), 5: (row, 0), 7: (deep, 0), 8: (culled, 2)}
    top = CL.Writer().translate(10, 20).submit(5).push_state().scale(2, 2).submit(5).pop_state().submit(7).submit(8)
    top.begin_path().rect(1, 1, 9, 9).fill(0xFFABCDEF, CL.fill_flags())  # drawn: the culling of list 8 ended with it

    model = CL.Model(sincos, lists={k: (w.ops, fl) for k, (w, fl) in lists.items()}).run(top.ops)
    expect = vgb.Batch(-1)
    model.rec.replay(expect)
    got = vgb.Batch(-1)
    ctx = vgb.make_cl_context(lists={k: (w.bytes(), fl) for k, (w, fl) in lists.items()})
    st = got.submit_command_list(top.bytes(), ctx)
    recorded_equal(got.recorded_streams(), expect.recorded_streams())
    for k, v in model.stats.items():
        assert getattr(st, k) == v, k
    assert st.nested == 2 + 6 + 16 + 1  # top: 5, 5 | each row: 3 x list 3 | 7 from the top + its 15 executions each submitting itself | 8
    assert st.culled == 1 and st.draws == 2 * 6 + 15 + 1
    np.testing.assert_array_equal(np.asarray(list(ctx.xform), np.float32), np.asarray(model.s["m"], np.float32))
    # without a list table nesting is reported as unsupported, nothing else changes
    plain = vgb.Batch(-1)
    st2 = plain.submit_command_list(top.bytes(), vgb.make_cl_context())
    assert st2.unsupported == 4 and st2.nested == 0 and st2.draws == 1
    # a depth limit of 1 executes the top list only
    lim = vgb.Batch(-1)
    st3 = lim.submit_command_list(top.bytes(), vgb.make_cl_context(lists={k: (w.bytes(), fl) for k, (w, fl) in lists.items()}, max_depth=1))
    assert st3.draws == 1 and st3.nested == 4


def test_collapsed_transform_does_not_abort_the_list():
    """A path begun under TransformScale(0, 0) (average scale 0) is legal in the reference (tessTol = inf, zero-alpha
    strokes, zero-area fills: nothing visible) and the list keeps executing.  The decoder must not reject it: the dead
    path's commands and draws are skipped, everything after PopState is recorded as usual."""
    w = (CL.Writer().push_state().scale(0, 0).begin_path().move_to(0, 0).cubic_to(10, 0, 20, 5, 30, 30).line_to(5, 40).close_path()
         .fill(0xFF112233, CL.fill_flags()).stroke(0xFF445566, 3.0, CL.stroke_flags(1, 1, True)).pop_state()
         .begin_path().rect(10, 10, 50, 40).fill(0xFF778899, CL.fill_flags()).set_scissor(0, 0, 100, 100)
         .begin_path().circle(50, 50, 20).stroke(0xFF000000, 2.0, CL.stroke_flags(0, 0, True)))
    got = vgb.Batch(-1)
    ctx = vgb.make_cl_context()
    st = got.submit_command_list(w.bytes(), ctx)
    assert st.paths == 3 and st.draws == 2 and st.culled == 2 and st.state_depth == 0
    assert st.scissor_changes == 1  # the second recorded draw runs under another scissor rect than the first
    expect = vgb.Batch(-1)
    expect.begin_path(1.0, 0.25, [1, 0, 0, 1, 0, 0])
    expect.rect(10, 10, 50, 40)
    expect.convex_fill(0xFF778899, True)
    expect.begin_path(1.0, 0.25, [1, 0, 0, 1, 0, 0])
    expect.circle(50, 50, 20)
    expect.stroke(0, 0xFF000000, 2.0, 0, 0)
    a, b = got.recorded_streams(), expect.recorded_streams()
    for k in ("cmd", "data", "draw_op", "draw_color", "draw_width", "path_scale"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


def test_skipped_draws_leave_an_ordering_marker():
    """a draw the decoder does not handle (gradient fill, concave fill, text ...) sits between two colour draws: the batch
    must not let them share a draw command -- the caller's own geometry for the skipped draw goes in between.  The mark is
    the one a scissor change leaves (vgb_batch_recorded_draw_scissor bit 31)."""
    import struct
    w = CL.Writer()
    w.begin_path().rect(10, 10, 50, 40).fill(0xFF778899, CL.fill_flags())
    w.begin_path().circle(100, 100, 30).fill(0xFF112233, CL.fill_flags())                       # merges with the first
    w.begin_path().rect(0, 0, 20, 20).raw(CL.FILL_GRADIENT, struct.pack("<IHH", 4, 0, 0))       # skipped: another draw type
    w.begin_path().circle(200, 100, 30).fill(0xFF445566, CL.fill_flags())                       # new command
    w.begin_path().move_to(0, 0).line_to(50, 0).line_to(25, 40).line_to(60, 60).close_path().fill(0xFF000000, CL.fill_flags(concave=True))  # skipped
    w.stroke(0xFFFFFFFF, 2.0, CL.stroke_flags(1, 1, True))                                      # new command
    w.begin_path().rect(5, 5, 9, 9).stroke(0xFF00FF00, 1.5, CL.stroke_flags(0, 0, True))        # merges
    got = vgb.Batch(-1)
    st = got.submit_command_list(w.bytes(), vgb.make_cl_context())
    assert st.unsupported == 2 and st.draws == 5 and st.scissor_changes == 2
    np.testing.assert_array_equal(got.recorded_draw_scissor() >> 31, [0, 0, 1, 1, 0])
