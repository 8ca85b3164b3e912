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
