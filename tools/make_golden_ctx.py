"""Generate tests/golden/ctx_scenes.npz from the reference's WHOLE context (oracle/_ref/libvgctx.so: src/vg.cpp itself against
the record-only bgfx stand-in, see oracle/vgctx_capi.cpp).  Needs /root/reference (`make -C oracle ref`); the output is
committed so that the command-list decoder, the restated merge step and the scissor-forced draw commands stay pinned to what
the real vg.cpp produced even where the reference tree is absent (tests/test_vgctx_golden.py).

    python tools/make_golden_ctx.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.pyref import CtxLib  # noqa: E402
from test_vgctx import scene, scissor_scene  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ctx_scenes.npz")
CANVAS = (1280, 720)
cases = {"plain": (scene(seed=3, shapes=40), 5000), "scissor": (scissor_scene(seed=5, shapes=60), 3000)}
d = {"canvas": np.asarray(CANVAS, np.uint32)}
ctx = CtxLib()
for name, (w, max_vb) in cases.items():
    real = ctx.run(w.bytes(), canvas=CANVAS, max_vb_vertices=max_vb)
    assert real["cl_bytes"] is not None
    d[f"{name}_list"] = np.frombuffer(w.bytes(), np.uint8)
    d[f"{name}_max_vb"] = np.uint32(max_vb)
    for k in ("pos", "color", "index", "uv"):
        d[f"{name}_{k}"] = real[k]
    for k in ("vb_id", "first_vertex", "num_vertices", "first_index", "num_indices", "scissor"):
        d[f"{name}_draw_{k}"] = np.ascontiguousarray(real["draws"][k])
    print(name, "vertices", len(real["pos"]), "indices", len(real["index"]), "draw commands", len(real["draws"]))
np.savez_compressed(OUT, **d)
print("wrote", OUT, os.path.getsize(OUT), "bytes")
