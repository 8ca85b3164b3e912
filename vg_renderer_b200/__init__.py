"""vg_renderer_b200 -- B200-native batched path flattening + stroke / convex-fill mesh generation.

Thin ctypes mirror of libvgb200.so's C ABI (include/vgb.h).  All computation happens in the
hand-written sm_100a CUDA kernels inside the shared library; Python is harness only.  There is no
CPU fallback: creating a Batch without a visible GPU raises.
"""
from .api import (Batch, Path, Stroker, VgbError, lib_path, load_library, device_count,  # noqa: F401
                  CAP_BUTT, CAP_ROUND, CAP_SQUARE, JOIN_MITER, JOIN_ROUND, JOIN_BEVEL,
                  OP_STROKE_AA, OP_CONVEX_FILL_AA, OP_STROKE, OP_STROKE_AA_THIN, OP_CONVEX_FILL,
                  WINDING_CCW, WINDING_CW, INDEX_U16, INDEX_U32, FinalCfg, DrawCmd, DRAW_CMD_DTYPE, make_final_cfg,
                  ClContext, ClStats, ClList, make_cl_context, Group, Merged, copy_from_device)
from .tables import BatchTables, Recorder  # noqa: F401
