"""What the reference's own vg.cpp handed to bgfx for two command lists (tests/golden/ctx_scenes.npz, written by
tools/make_golden_ctx.py from oracle/_ref/libvgctx.so), checked WITHOUT the reference tree: the product's command-list decoder
records the list, this repo's plain-C restatement (oracle/vg_oracle.c) runs the recording and the merge step, and the result
must be the real context's vertex streams, index buffer and draw-command table -- scissor-forced commands and rects included."""
import os

import numpy as np
import pytest

import vg_renderer_b200 as vgb
from oracle.pyref import OracleLib
from vg_renderer_b200.tables import tables_from_recorded

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ctx_scenes.npz")


@pytest.mark.parametrize("name", ["plain", "scissor"])
def test_decoder_and_restated_merge_step_equal_the_recorded_real_context(name):
    g = np.load(GOLD)
    canvas = tuple(int(v) for v in g["canvas"])
    buf = g[f"{name}_list"].tobytes()
    got = vgb.Batch(-1)
    st = got.submit_command_list(buf, vgb.make_cl_context(canvas=(float(canvas[0]), float(canvas[1]))))
    assert st.unsupported == 0
    orc = OracleLib()
    exp = orc.run_batch(tables_from_recorded(got.recorded_streams()))
    np.testing.assert_array_equal(exp["pos"], g[f"{name}_pos"])
    np.testing.assert_array_equal(exp["color"], g[f"{name}_color"])
    uv0 = g[f"{name}_uv"][0]
    cfg = vgb.make_final_cfg(int(g[f"{name}_max_vb"]), uv=(int(uv0[0]), int(uv0[1])))
    idx, uv, cmds = orc.final_layout(exp["mesh_vtx_off"], exp["mesh_idx_off"], exp["index"], cfg, mesh_scissor=got.recorded_mesh_scissor())
    np.testing.assert_array_equal(idx, g[f"{name}_index"])
    np.testing.assert_array_equal(uv, g[f"{name}_uv"])
    assert len(cmds) == len(g[f"{name}_draw_vb_id"]) and (name == "plain" or st.scissor_changes > 10)
    for k in ("vb_id", "first_vertex", "first_index", "num_vertices", "num_indices"):
        np.testing.assert_array_equal(cmds[k], g[f"{name}_draw_{k}"], err_msg=k)
    # the rects: bgfx sees "no scissor" (zeros in the record) for canvas-rect commands that precede any other rect (vg.cpp:1156-1186)
    sc, ours = g[f"{name}_draw_scissor"].copy(), got.scissor_rects()[cmds["scissor_id"]]
    sc[(sc == 0).all(axis=1) & ~(ours == 0).all(axis=1)] = (0, 0, canvas[0], canvas[1])
    np.testing.assert_array_equal(sc, ours)
