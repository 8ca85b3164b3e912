"""The reference's WHOLE context as the oracle for the rows either side of the hot path: src/vg.cpp
itself -- command-list recorder and executor (vg.cpp:2323-2966, 4273-4637), ctxFillPathColor / ctxStrokePathColor
(3061-3492), the merge step (5207-5405) and vg::end (1076-1288) -- compiled where it lies against a record-only bgfx
stand-in (oracle/vgctx_capi.cpp, oracle/bgfx_stub).  Checked here, on the CPU:
  * tests/cmdlist.py's writer produces the bytes vg::clXXX really records,
  * the product's command-list decoder records the draws ctxSubmitCommandList really issues: the compiled hot path
    (oracle/_ref/libvgref.so) run on the decoder's recording gives vg.cpp's vertex streams,
  * the restated merge step (ref_capi.cpp: vgref_final_layout) gives vg.cpp's index buffer and draw-command table.
The -m gpu test compares the GPU's final layout with vg.cpp's output directly."""
import numpy as np
import pytest

import vg_renderer_b200 as vgb
from oracle.pyref import CtxLib, OracleLib, RefLib, have_ctx, have_ref
from vg_renderer_b200.tables import tables_from_recorded

import cmdlist as CL

pytestmark = pytest.mark.skipif(not (have_ctx() and have_ref("scalar")), reason="oracle/_ref/libvgctx.so not built (needs /root/reference)")


def scene(seed=3, shapes=60):
    """transforms, state stack, global alpha, scaled / fixed-width / thin / non-AA strokes, AA and plain fills of every path
    builder -- one scissor throughout (the batch does not carry scissor rects) -- and enough vertices to roll vertex buffers"""
    rng = np.random.default_rng(seed)
    w = CL.Writer()
    w.translate(20, 10).scale(1.5, 1.5)
    for k in range(shapes):
        if k % 7 == 0:
            w.push_state().rotate(float(np.float32(rng.uniform(-0.6, 0.6)))).translate(float(np.float32(rng.uniform(-30, 30))), 5.0)
        if k % 11 == 0:
            w.global_alpha(float(np.float32(rng.uniform(0.3, 1.0))))
        x, y = [float(np.float32(v)) for v in rng.uniform(20, 500, 2)]
        w.begin_path()
        kind = k % 8
        if kind == 0:
            w.rounded_rect(x, y, 120, 60, float(np.float32(rng.uniform(2, 25))))
        elif kind == 1:
            w.circle(x, y, float(np.float32(rng.uniform(5, 60))))
        elif kind == 2:
            w.ellipse(x, y, 50, 20)
        elif kind == 3:
            w.move_to(x, y).cubic_to(x + 50, y - 80, x + 120, y + 90, x + 180, y).quadratic_to(x + 200, y + 60, x + 90, y + 70).close_path()
        elif kind == 4:
            w.rect(x, y, 70, 30)
        elif kind == 5:
            w.move_to(x, y).arc_to(x + 60, y, x + 60, y + 60, 20).line_to(x + 10, y + 80)
        elif kind == 6:
            w.move_to(x + 40, y).arc(x, y, 40, 0.0, 4.0, bool(k & 8)).line_to(x, y)
        else:
            w.move_to(x, y).polyline(rng.uniform(0, 200, 10).astype(np.float32) + np.float32(x))
        if kind in (0, 1, 2, 3, 4):
            w.fill(0xFF000000 | int(rng.integers(0, 1 << 24)), CL.fill_flags(aa=bool(k % 3)))
        if kind != 4:
            width = float(np.float32(rng.uniform(0.3, 9.0)))
            w.stroke(0xC0000000 | int(rng.integers(0, 1 << 24)), width, CL.stroke_flags(int(rng.integers(0, 3)), int(rng.integers(0, 3)), bool((k + 1) % 5), fixed=bool(k % 6 == 0)))
        if k % 7 == 3:
            w.pop_state()
    return w


# meaningful payload bytes per command type (what the cl* functions CMD_WRITE, vg.cpp:2402-2960); the rest of a record --
# the header's 8 padding bytes and the payload's tail up to the next multiple of 16 -- is whatever the allocator left there
_PAYLOAD = {CL.BEGIN_PATH: 0, CL.MOVE_TO: 8, CL.LINE_TO: 8, CL.CUBIC_TO: 24, CL.QUAD_TO: 16, CL.ARC_TO: 20, CL.ARC: 24, CL.RECT: 16, CL.ROUNDED_RECT: 20,
            CL.ROUNDED_RECT_VARYING: 32, CL.CIRCLE: 12, CL.ELLIPSE: 16, CL.CLOSE_PATH: 0, CL.FILL_COLOR: 8, CL.STROKE_COLOR: 12, CL.PUSH_STATE: 0,
            CL.POP_STATE: 0, CL.RESET_SCISSOR: 0, CL.SET_SCISSOR: 16, CL.INTERSECT_SCISSOR: 16, CL.T_IDENTITY: 0, CL.T_SCALE: 8, CL.T_TRANSLATE: 8,
            CL.T_ROTATE: 4, CL.T_MULT: 28, CL.SET_VIEW_BOX: 16, CL.SET_GLOBAL_ALPHA: 4}


def records(buf):
    """[(type, size, meaningful payload bytes)] of a command-list byte stream"""
    out, i = [], 0
    while i < len(buf):
        t, n = np.frombuffer(buf[i:i + 8], np.uint32)
        p = buf[i + 16:i + 16 + int(n)]
        m = 4 + 8 * int(np.frombuffer(p[:4], np.uint32)[0]) if t == CL.POLYLINE else 2 if t == CL.SUBMIT_LIST else _PAYLOAD[int(t)]
        out.append((int(t), int(n), bytes(p[:m])))
        i += 16 + int(n)
    assert i == len(buf)
    return out


def decoded_tables(w, canvas):
    got = vgb.Batch(-1)
    ctx = vgb.make_cl_context(canvas=(float(canvas[0]), float(canvas[1])))
    st = got.submit_command_list(w.bytes(), ctx)
    assert st.unsupported == 0 and st.scissor_changes == 0
    return tables_from_recorded(got.recorded_streams()), st


def scissor_scene(seed=5, shapes=80):
    """scene() with the scissor moving underneath it: Set / Intersect / Reset between draws, PopState restoring a different
    and the same rect, a scissor command followed by a draw that has no vertices (the force flag must wait for the next
    mesh that has some, vg.cpp:5368-5402), a change between the fill and the stroke of one path, repeated rects"""
    rng = np.random.default_rng(seed)
    w = CL.Writer()
    w.translate(10, 10)
    for k in range(shapes):
        x, y = [float(np.float32(v)) for v in rng.uniform(20, 500, 2)]
        r = k % 13
        if r == 1:
            w.set_scissor(30.0 + k, 20.0, 600.0, 400.0)
        elif r == 3:
            w.push_state().intersect_scissor(100.0, 80.0, 500.0 - k, 300.0)
        elif r == 5:
            w.pop_state()                                              # back to the rect of r == 1: differs from the last draw's
        elif r == 6:
            w.push_state().translate(3.0, 4.0)
        elif r == 7:
            w.pop_state()                                              # same rect as the last draw's: no new command
        elif r == 8:
            w.reset_scissor()
            w.begin_path().move_to(x, y).stroke(0xFF000000, 2.0, CL.stroke_flags(0, 0, True))   # one point: no vertices
        elif r == 10:
            w.set_scissor(30.0, 20.0, 600.0, 400.0)                    # a rect seen before
        w.begin_path()
        if k % 3 == 0:
            w.rounded_rect(x, y, 90, 50, 9.0)
        elif k % 3 == 1:
            w.circle(x, y, float(np.float32(rng.uniform(5, 50))))
        else:
            w.move_to(x, y).cubic_to(x + 40, y - 60, x + 90, y + 70, x + 150, y).line_to(x + 100, y + 90)
        if k % 3 != 2:
            w.fill(0xFF000000 | int(rng.integers(0, 1 << 24)), CL.fill_flags(aa=bool(k % 2)))
        if r == 11:
            w.intersect_scissor(0.0, 0.0, 900.0, 500.0)                # between the fill and the stroke of the same path
        w.stroke(0xD0000000 | int(rng.integers(0, 1 << 24)), float(np.float32(rng.uniform(0.5, 6.0))), CL.stroke_flags(int(rng.integers(0, 3)), int(rng.integers(0, 3)), True))
    return w


def assert_cmds_equal_real(real, cmds, rects, canvas=(1280, 720)):
    d = real["draws"]
    assert len(d) == len(cmds)
    for k in ("vb_id", "first_vertex", "first_index", "num_vertices", "num_indices"):
        np.testing.assert_array_equal(d[k], cmds[k], err_msg=k)
    # vg::end (vg.cpp:1156-1186) starts from "previous rect = the whole canvas, cached bgfx scissor = none": commands with
    # the canvas rect that come before any other rect reach bgfx with no scissor at all, which the stand-in records as zeros
    # (a genuinely empty rect {0, 0, 0, 0} looks the same there: accept either for an all-zero record)
    sc, ours = d["scissor"].copy(), rects[cmds["scissor_id"]]
    zero = (sc == 0).all(axis=1)
    sc[zero & ~(ours == 0).all(axis=1)] = (0, 0, canvas[0], canvas[1])
    np.testing.assert_array_equal(sc, ours)


@pytest.mark.parametrize("max_vb", [65536, 3000])
def test_scissor_changes_open_draw_commands_like_the_real_context(max_vb):
    """the batch carries the scissor (vgb_batch_set_scissor, driven by the decoder): the restated merge step with the
    per-mesh marks gives the real context's draw-command table, rects included, and its rebased indices"""
    canvas = (1280, 720)
    w = scissor_scene()
    real = CtxLib().run(w.bytes(), canvas=canvas, max_vb_vertices=max_vb)
    got = vgb.Batch(-1)
    st = got.submit_command_list(w.bytes(), vgb.make_cl_context(canvas=(1280.0, 720.0)))
    assert st.unsupported == 0 and st.scissor_changes > 20
    ref = RefLib("scalar")
    exp = ref.run_batch(tables_from_recorded(got.recorded_streams()))
    np.testing.assert_array_equal(real["pos"], exp["pos"])
    ms = got.recorded_mesh_scissor()
    assert int((ms >> 31).sum()) == st.scissor_changes and (exp["mesh_vtx_off"][1:] == exp["mesh_vtx_off"][:-1]).any()
    cfg = vgb.make_final_cfg(max_vb, uv=(int(real["uv"][0, 0]), int(real["uv"][0, 1])))
    rects = got.scissor_rects()
    for lib in (ref, OracleLib()):
        idx, uv, cmds = lib.final_layout(exp["mesh_vtx_off"], exp["mesh_idx_off"], exp["index"], cfg, mesh_scissor=ms)
        np.testing.assert_array_equal(real["index"], idx)
        assert_cmds_equal_real(real, cmds, rects)
    assert len(np.unique(real["draws"]["vb_id"])) > 1 or max_vb == 65536
    assert len(real["draws"]) > len(np.unique(real["draws"]["vb_id"])) + 15


@pytest.mark.parametrize("max_vb", [65536, 5000])
def test_decoder_and_merge_step_against_the_real_context(max_vb):
    canvas = (1280, 720)
    w = scene()
    real = CtxLib().run(w.bytes(), canvas=canvas, max_vb_vertices=max_vb)
    # (1) the byte format: what vg::clXXX recorded is what the writer produced
    assert len(real["cl_bytes"]) == len(w.bytes()) and records(real["cl_bytes"]) == records(w.bytes())
    # (2) the executor: the hot path run on the DECODER's recording gives the real context's vertex streams
    t, st = decoded_tables(w, canvas)
    ref = RefLib("scalar")
    exp = ref.run_batch(t)
    assert exp["num_vertices"] == len(real["pos"])
    np.testing.assert_array_equal(real["pos"], exp["pos"])
    np.testing.assert_array_equal(real["color"], exp["color"])
    assert len(np.unique(real["uv"], axis=0)) == 1  # every path vertex samples the white pixel
    # (3) the merge step: restated book-keeping (ref_capi.cpp) against the real createDrawCommand_VertexColor / vg::end
    cfg = vgb.make_final_cfg(max_vb, uv=(int(real["uv"][0, 0]), int(real["uv"][0, 1])))
    idx, uv, cmds = ref.final_layout(exp["mesh_vtx_off"], exp["mesh_idx_off"], exp["index"], cfg)
    np.testing.assert_array_equal(real["index"], idx)
    np.testing.assert_array_equal(real["uv"], uv)
    d = real["draws"]
    assert len(d) == len(cmds) and (max_vb == 65536 or len(d) > 1)
    for k in ("vb_id", "first_vertex", "first_index", "num_vertices", "num_indices"):
        np.testing.assert_array_equal(d[k], cmds[k], err_msg=k)
    np.testing.assert_array_equal(real["vb_counts"], np.bincount(cmds["vb_id"], weights=cmds["num_vertices"]).astype(np.uint32)[cmds["vb_id"].min():])


def test_scissor_changes_are_the_real_contexts_extra_draw_commands():
    """the decoder's scissor_changes stat = the draw commands vg.cpp adds because the scissor rect changed"""
    w = CL.Writer()
    for k in range(6):
        if k in (2, 4):
            w.set_scissor(10.0 * k, 5.0, 300.0, 200.0)
        w.begin_path().circle(100 + 20 * k, 100, 30).fill(0xFF336699, CL.fill_flags())
    real = CtxLib().run(w.bytes())
    got = vgb.Batch(-1)
    st = got.submit_command_list(w.bytes(), vgb.make_cl_context())
    assert st.scissor_changes == 2 and len(real["draws"]) == 1 + st.scissor_changes
    np.testing.assert_array_equal(got.recorded_draw_scissor() >> 31, [0, 0, 1, 0, 1, 0])
    np.testing.assert_array_equal(got.scissor_rects()[got.recorded_draw_scissor() & 0x7FFFFFFF][2:], real["draws"]["scissor"][[1, 1, 2, 2]])


@pytest.mark.gpu
@pytest.mark.parametrize("max_vb", [65536, 3000])
def test_gpu_final_layout_with_scissor_changes_equals_the_real_context(max_vb):
    """scissor changes inside the list: the GPU's draw-command table splits where vg.cpp's does and names the same rects"""
    w = scissor_scene(seed=12, shapes=120)
    real = CtxLib().run(w.bytes(), canvas=(1280, 720), max_vb_vertices=max_vb)
    b = vgb.Batch(0)
    st = b.submit_command_list(w.bytes(), vgb.make_cl_context(canvas=(1280.0, 720.0)))
    assert st.scissor_changes > 20
    b.set_final_layout(vgb.make_final_cfg(max_vb, uv=(int(real["uv"][0, 0]), int(real["uv"][0, 1]))))
    b.run()
    out = b.download()
    uv, cmds = b.download_final()
    np.testing.assert_array_equal(out["pos"], real["pos"])
    np.testing.assert_array_equal(out["color"], real["color"])
    np.testing.assert_array_equal(uv, real["uv"])
    np.testing.assert_array_equal(out["index"], real["index"])
    assert_cmds_equal_real(real, cmds, b.scissor_rects())


@pytest.mark.gpu
@pytest.mark.parametrize("max_vb", [65536, 5000])
def test_gpu_final_layout_equals_the_real_context(max_vb):
    """command-list bytes -> decoder -> GPU pipeline with the final layout == what the reference's own vg.cpp hands to
    bgfx for the same list: vertex streams, uv stream, rebased uint16 indices, draw-command table"""
    canvas = (1280, 720)
    w = scene(seed=9, shapes=90)
    real = CtxLib().run(w.bytes(), canvas=canvas, max_vb_vertices=max_vb)
    b = vgb.Batch(0)
    st = b.submit_command_list(w.bytes(), vgb.make_cl_context(canvas=(float(canvas[0]), float(canvas[1]))))
    assert st.scissor_changes == 0
    b.set_final_layout(vgb.make_final_cfg(max_vb, uv=(int(real["uv"][0, 0]), int(real["uv"][0, 1]))))
    b.run()
    out = b.download()
    uv, cmds = b.download_final()
    np.testing.assert_array_equal(out["pos"], real["pos"])
    np.testing.assert_array_equal(out["color"], real["color"])
    np.testing.assert_array_equal(uv, real["uv"])
    np.testing.assert_array_equal(out["index"], real["index"])
    d = real["draws"]
    assert len(d) == len(cmds)
    for k in ("vb_id", "first_vertex", "first_index", "num_vertices", "num_indices"):
        np.testing.assert_array_equal(d[k], cmds[k], err_msg=k)


@pytest.mark.parametrize("cull", [0, 2])
def test_culling_scissor_and_matrix_commands_against_the_real_context(cull):
    """AllowCommandCulling (vg.cpp:4307, 4338-4341) with zero-sized / restored scissors, IntersectScissor, SetViewBox,
    TransformMult (both orders), TransformIdentity: the decoder's recording run through the compiled hot path must give
    the vertices the real executor produced (draw-command boundaries differ: several scissor rects)."""
    w = CL.Writer()
    w.set_view_box(0.0, 0.0, 640.0, 360.0)
    w.begin_path().circle(100, 100, 40).fill(0xFF102030, CL.fill_flags())
    w.push_state().set_scissor(50.0, 50.0, 0.5, 300.0)                       # narrower than a pixel: strokers are culled when allowed
    w.begin_path().rect(10, 10, 80, 40).fill(0xFF405060, CL.fill_flags()).stroke(0xFF000000, 2.0, CL.stroke_flags(1, 1, True))
    w.intersect_scissor(0.0, 0.0, 10.0, 10.0)
    w.begin_path().ellipse(200, 100, 60, 20).stroke(0xFF778899, 3.0, CL.stroke_flags(0, 2, True))
    w.pop_state()
    w.mult([1.0, 0.25, -0.25, 1.0, 12.0, -7.0], False).mult([0.9, 0.0, 0.0, 1.1, 3.0, 4.0], True)
    w.begin_path().move_to(20, 200).cubic_to(80, 120, 160, 280, 240, 200).stroke(0xFFAA5500, 6.0, CL.stroke_flags(2, 0, True))
    w.set_scissor(100.0, 50.0, 200.0, 150.0).intersect_scissor(150.0, 60.0, 400.0, 30.0)
    w.begin_path().rounded_rect(120, 60, 150, 90, 12).fill(0xFF00FF00, CL.fill_flags(aa=False)).stroke(0x80FFFFFF, 0.4, CL.stroke_flags(0, 0, True))
    w.reset_scissor().identity().global_alpha(0.5)
    w.begin_path().circle(300, 300, 25).fill(0xFFFF00FF, CL.fill_flags())
    real = CtxLib().run(w.bytes(), list_flags=cull, canvas=(1280, 720))
    assert records(real["cl_bytes"]) == records(w.bytes())
    got = vgb.Batch(-1)
    st = got.submit_command_list(w.bytes(), vgb.make_cl_context(canvas=(1280.0, 720.0), flags=cull))
    assert (st.culled > 0) == (cull == 2)
    exp = RefLib("scalar").run_batch(tables_from_recorded(got.recorded_streams()))
    np.testing.assert_array_equal(real["pos"], exp["pos"])
    np.testing.assert_array_equal(real["color"], exp["color"])
    assert int(real["draws"]["num_indices"].sum()) == exp["num_indices"]


def plain_scene(seed=21, shapes=24):
    """paths, fills and strokes with no transform or state commands of their own: one matrix for the whole list"""
    rng = np.random.default_rng(seed)
    w = CL.Writer()
    for k in range(shapes):
        x, y = [float(np.float32(v)) for v in rng.uniform(-150, 150, 2)]
        w.begin_path()
        if k % 4 == 0:
            w.rounded_rect(x, y, 80, 45, 11.0)
        elif k % 4 == 1:
            w.circle(x, y, float(np.float32(rng.uniform(4, 40))))
        elif k % 4 == 2:
            w.move_to(x, y).cubic_to(x + 30, y - 50, x + 70, y + 60, x + 120, y).quadratic_to(x + 140, y + 50, x + 60, y + 55).close_path()
        else:
            w.move_to(x, y).line_to(x + 40, y + 10).arc_to(x + 90, y + 10, x + 90, y + 60, 15).line_to(x + 20, y + 70)
        if k % 4 != 3:
            w.fill(0xFF000000 | int(rng.integers(0, 1 << 24)), CL.fill_flags(aa=bool(k % 3)))
        w.stroke(0xE0000000 | int(rng.integers(0, 1 << 24)), float(np.float32(rng.uniform(0.6, 5.0))), CL.stroke_flags(int(rng.integers(0, 3)), int(rng.integers(0, 3)), True))
    return w


def test_shape_cache_replay_is_the_restated_instance_replay():
    """SURVEY 8 f3 against the reference's own shape cache: a Cacheable list submitted under n matrices of equal average
    scale.  Submission 0 tessellates (= the decoder's recording under that matrix through the compiled hot path) and
    caches Tinv(mesh) (addCachedCommand, vg.cpp:5803-5841); submissions 1.. are submitCachedMesh (vg.cpp:6137-6166):
    pos = T_k(Tinv(pos_0)) by vgutil::batchTransformPositions, colours and local indices copied -- the contract
    vgb_batch_replay implements (its GPU output is checked against the same two reference functions in
    tests/test_gpu_parity.py::test_replay_matches_reference_transform)."""
    s = 1.5
    xf = np.asarray([[s, 0, 0, s, 400.25, 300.5], [0, s, -s, 0, 900.0, 120.0], [-s, 0, 0, -s, 1000.5, 650.0], [s, 0, 0, s, 77.0, 500.125],
                     [0, -s, s, 0, 333.0, 444.0]], np.float32)
    w = plain_scene()
    real = CtxLib().run(w.bytes(), list_flags=1, canvas=(1280, 720), xforms=xf)
    got = vgb.Batch(-1)
    st = got.submit_command_list(w.bytes(), vgb.make_cl_context(xform=[float(v) for v in xf[0]], canvas=(1280.0, 720.0)))
    assert st.unsupported == 0
    ref = RefLib("scalar")
    exp = ref.run_batch(tables_from_recorded(got.recorded_streams()))
    V, I, n = exp["num_vertices"], exp["num_indices"], len(xf)
    assert len(real["pos"]) == n * V and len(real["index"]) == n * I
    np.testing.assert_array_equal(real["pos"][:V], exp["pos"])
    cached = ref.transform_positions(exp["pos"], real["inv0"])
    for k in range(1, n):
        np.testing.assert_array_equal(real["pos"][k * V:(k + 1) * V], ref.transform_positions(cached, xf[k]), err_msg=f"instance {k}")
        np.testing.assert_array_equal(real["color"][k * V:(k + 1) * V], exp["color"])
    # the index buffer: every instance's local indices, rebased by the merge step like any other mesh
    mvo = np.concatenate([exp["mesh_vtx_off"][:-1] + k * V for k in range(n)] + [[n * V]]).astype(np.uint64)
    mio = np.concatenate([exp["mesh_idx_off"][:-1] + k * I for k in range(n)] + [[n * I]]).astype(np.uint64)
    idx, _, cmds = ref.final_layout(mvo, mio, np.tile(exp["index"], n), vgb.make_final_cfg(65536))
    np.testing.assert_array_equal(real["index"], idx)
    assert len(cmds) == len(real["draws"])
    # a matrix of another average scale invalidates the cache (vg.cpp:4289-4299): that submission is a fresh tessellation
    xf2 = np.vstack([xf[:2], [[2.0, 0, 0, 2.0, 50.0, 60.0]]]).astype(np.float32)
    real2 = CtxLib().run(w.bytes(), list_flags=1, canvas=(1280, 720), xforms=xf2)
    got2 = vgb.Batch(-1)
    got2.submit_command_list(w.bytes(), vgb.make_cl_context(xform=[float(v) for v in xf2[2]], canvas=(1280.0, 720.0)))
    exp2 = ref.run_batch(tables_from_recorded(got2.recorded_streams()))
    np.testing.assert_array_equal(real2["pos"][2 * V:], exp2["pos"])


def test_nested_command_lists_against_the_real_context():
    """SubmitCommandList inside lists (vg.cpp:4611-4618): state left behind by a child (VG_CONFIG_COMMAND_LIST_PRESERVE_STATE
    is 0), a child submitted under different matrices, a handle that is not a list (skipped), a list that submits itself
    until ContextConfig::m_MaxCommandListDepth (16, vg.cpp:4278-4283) stops it, a child with AllowCommandCulling."""
    icon = CL.Writer().push_state().scale(0.5, 0.5).begin_path().rounded_rect(0, 0, 40, 40, 6).fill(0xFF336699, CL.fill_flags())
    icon.stroke(0xFF000000, 2.0, CL.stroke_flags(1, 1, True)).pop_state().translate(50, 0)      # leaves a translation behind
    row = CL.Writer().submit(0).submit(0).submit(9).rotate(0.2).submit(0)                         # 9: not a list -> skipped
    deep = CL.Writer().translate(7, 5).begin_path().circle(20, 20, 6).fill(0xFFFFFFFF, CL.fill_flags(aa=False)).submit(2)  # submits itself
    culled = CL.Writer().set_scissor(0, 0, 0.2, 0.2).begin_path().rect(0, 0, 5, 5).fill(0xFF00FF00, CL.fill_flags())
    kids = [(icon, 0), (row, 0), (deep, 0), (culled, 2)]
    top = CL.Writer().translate(100, 200).submit(1).push_state().scale(2, 2).submit(1).pop_state().submit(2).submit(3)
    top.reset_scissor().begin_path().rect(1, 1, 90, 90).fill(0xFFABCDEF, CL.fill_flags())
    real = CtxLib().run(top.bytes(), canvas=(1280, 720), children=[(w.bytes(), fl) for w, fl in kids])
    assert records(real["cl_bytes"]) == records(top.bytes())
    got = vgb.Batch(-1)
    ctx = vgb.make_cl_context(canvas=(1280.0, 720.0), lists={k: (w.bytes(), fl) for k, (w, fl) in enumerate(kids)})
    st = got.submit_command_list(top.bytes(), ctx)
    assert st.nested == 2 + 6 + 16 + 1 and st.culled == 1 and st.draws == 2 * 6 + 15 + 1
    ref = RefLib("scalar")
    exp = ref.run_batch(tables_from_recorded(got.recorded_streams()))
    np.testing.assert_array_equal(real["pos"], exp["pos"])
    np.testing.assert_array_equal(real["color"], exp["color"])
    idx, uv, cmds = ref.final_layout(exp["mesh_vtx_off"], exp["mesh_idx_off"], exp["index"], vgb.make_final_cfg(65536), mesh_scissor=got.recorded_mesh_scissor())
    np.testing.assert_array_equal(real["index"], idx)
    assert_cmds_equal_real(real, cmds, got.scissor_rects())


@pytest.mark.parametrize("seed", range(12))
def test_random_lists_against_the_real_context(seed):
    """the fuzz lists of tests/test_cmdlist.py (every path builder, strokes of all kinds, fills, state stack, transforms,
    scissor commands, global alpha, view box; culling on for odd seeds) through the real vg.cpp: vertex streams from the
    decoder's recording + compiled hot path, index buffer and draw-command table from the restated merge step"""
    from test_cmdlist import _random_list
    w = _random_list(seed, n_ops=500, concave=False)
    cull = 2 if seed % 2 else 0
    real = CtxLib().run(w.bytes(), list_flags=cull, canvas=(1280, 720))
    assert records(real["cl_bytes"]) == records(w.bytes())
    got = vgb.Batch(-1)
    st = got.submit_command_list(w.bytes(), vgb.make_cl_context(canvas=(1280.0, 720.0), flags=cull))
    assert st.unsupported == 0
    ref = RefLib("scalar")
    exp = ref.run_batch(tables_from_recorded(got.recorded_streams()))
    assert len(real["pos"]) == exp["num_vertices"]
    np.testing.assert_array_equal(real["pos"], exp["pos"])
    np.testing.assert_array_equal(real["color"], exp["color"])
    idx, uv, cmds = ref.final_layout(exp["mesh_vtx_off"], exp["mesh_idx_off"], exp["index"], vgb.make_final_cfg(65536), mesh_scissor=got.recorded_mesh_scissor())
    np.testing.assert_array_equal(real["index"], idx)
    assert_cmds_equal_real(real, cmds, got.scissor_rects())


@pytest.mark.parametrize("seed", range(6))
def test_random_nested_lists_against_the_real_context(seed):
    """three fuzz lists as children (one with AllowCommandCulling, later ones submitting earlier ones), a top list that
    submits them under changing state, an id that is no list: decoder recording + compiled hot path + restated merge step
    against the real context (400 more seeds of this generator were run once offline: no difference)"""
    from test_cmdlist import _random_list
    rng = np.random.default_rng(seed)
    kids = [(_random_list(seed * 7 + k, n_ops=int(rng.integers(20, 120)), concave=False), int(rng.choice([0, 0, 2]))) for k in range(3)]
    for w, _ in kids:  # balanced state stacks: the reference does not check overflow in release builds (VG_CHECK, vg.cpp:3936)
        depth = sum(1 for t, _ in w.ops if t == CL.PUSH_STATE) - sum(1 for t, _ in w.ops if t == CL.POP_STATE)
        for _ in range(max(depth, 0)):
            w.pop_state()
    for k in (1, 2):
        if rng.random() < 0.6:
            kids[k][0].submit(int(rng.integers(0, k)))
    top = CL.Writer().translate(float(np.float32(rng.uniform(0, 50))), 3.0)
    for _ in range(int(rng.integers(2, 7))):
        top.submit(int(rng.choice([0, 1, 2, 9])))  # 9: not a list (3 would be the top list itself)
        if rng.random() < 0.5:
            top.begin_path().circle(float(np.float32(rng.uniform(50, 400))), 100.0, 20.0).fill(0xFF336699, CL.fill_flags())
        if rng.random() < 0.3:
            top.push_state().scale(1.5, 0.75).submit(int(rng.integers(0, 3))).pop_state()
    cull = 2 if seed % 2 else 0
    got = vgb.Batch(-1)
    got.submit_command_list(top.bytes(), vgb.make_cl_context(canvas=(1280.0, 720.0), flags=cull, lists={k: (w.bytes(), fl) for k, (w, fl) in enumerate(kids)}))
    ref = RefLib("scalar")
    exp = ref.run_batch(tables_from_recorded(got.recorded_streams()))
    real = CtxLib().run(top.bytes(), list_flags=cull, canvas=(1280, 720), children=[(w.bytes(), fl) for w, fl in kids])
    np.testing.assert_array_equal(real["pos"], exp["pos"])
    np.testing.assert_array_equal(real["color"], exp["color"])
    idx, uv, cmds = ref.final_layout(exp["mesh_vtx_off"], exp["mesh_idx_off"], exp["index"], vgb.make_final_cfg(65536), mesh_scissor=got.recorded_mesh_scissor())
    np.testing.assert_array_equal(real["index"], idx)
    assert_cmds_equal_real(real, cmds, got.scissor_rects())
