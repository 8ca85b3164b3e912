"""ctypes binding of libvgb200.so (include/vgb.h).

Classes mirror the reference's interfaces for this path:
  Batch    -- record / run / read for many paths (the batched entry the GPU needs; recording
              calls are named after vg::beginPath / moveTo / ... / strokePath / fillPath)
  Path     -- include/vg/path.h:19-38     (createPath .. pathGetNumSubPaths)
  Stroker  -- include/vg/stroker.h:11-70  (createStroker .. strokerConvexFillAA)
The library must be built first (`python -c "import __graft_entry__ as g; g.build()"` or
`make -C vg_renderer_b200/csrc`); loading fails loudly if it is missing -- there is no fallback.
"""
import ctypes as C
import os

import numpy as np

CAP_BUTT, CAP_ROUND, CAP_SQUARE = 0, 1, 2
JOIN_MITER, JOIN_ROUND, JOIN_BEVEL = 0, 1, 2
WINDING_CCW, WINDING_CW = 0, 1
OP_STROKE_AA, OP_CONVEX_FILL_AA, OP_STROKE, OP_STROKE_AA_THIN, OP_CONVEX_FILL = range(5)
INDEX_U16, INDEX_U32 = 0, 1
FILL_ORDER_SSE = 2

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    # VGB_LIB selects another build of the same library (tuning variants under lib/variants/)
    return os.environ.get("VGB_LIB") or os.path.join(_HERE, "lib", "libvgb200.so")


class VgbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vgb error {code}: {msg}")
        self.code = code


class Sizes(C.Structure):
    _fields_ = [("num_vertices", C.c_uint64), ("num_indices", C.c_uint64), ("num_points", C.c_uint64),
                ("num_subpaths", C.c_uint32), ("num_meshes", C.c_uint32), ("kernel_launches", C.c_uint32),
                ("reserved", C.c_uint32)]


class FinalCfg(C.Structure):
    """vgb_final_cfg: the context state the merge step starts from (include/vgb.h)."""
    _fields_ = [("max_vb_vertices", C.c_uint32), ("vb_id", C.c_uint32), ("vb_count", C.c_uint32),
                ("cmd_vertices", C.c_uint32), ("ib_count", C.c_uint32), ("uv", C.c_int16 * 2)]


class DrawCmd(C.Structure):
    """vgb_draw_cmd"""
    _fields_ = [("vb_id", C.c_uint32), ("first_vertex", C.c_uint32), ("first_index", C.c_uint32),
                ("num_vertices", C.c_uint32), ("num_indices", C.c_uint32), ("first_mesh", C.c_uint32),
                ("num_meshes", C.c_uint32), ("scissor_id", C.c_uint32), ("stream_vertex", C.c_uint64),
                ("stream_index", C.c_uint64)]


DRAW_CMD_DTYPE = np.dtype([("vb_id", "<u4"), ("first_vertex", "<u4"), ("first_index", "<u4"), ("num_vertices", "<u4"),
                           ("num_indices", "<u4"), ("first_mesh", "<u4"), ("num_meshes", "<u4"), ("scissor_id", "<u4"),
                           ("stream_vertex", "<u8"), ("stream_index", "<u8")])


def make_final_cfg(max_vb_vertices=65536, vb_id=0, vb_count=0, cmd_vertices=0, ib_count=0, uv=(16, 16)):
    c = FinalCfg()
    c.max_vb_vertices, c.vb_id, c.vb_count, c.cmd_vertices, c.ib_count = max_vb_vertices, vb_id, vb_count, cmd_vertices, ib_count
    c.uv[0], c.uv[1] = uv
    return c


class ClList(C.Structure):
    """vgb_cl_list: one of the context's command lists (for SubmitCommandList)"""
    _fields_ = [("commands", C.c_void_p), ("size", C.c_uint32), ("flags", C.c_uint32)]


class ClContext(C.Structure):
    """vgb_cl_context"""
    _fields_ = [("xform", C.c_float * 6), ("scissor", C.c_float * 4), ("global_alpha", C.c_float), ("tess_tolerance", C.c_float),
                ("canvas_width", C.c_float), ("canvas_height", C.c_float), ("flags", C.c_uint32), ("num_lists", C.c_uint32),
                ("lists", C.POINTER(ClList)), ("max_depth", C.c_uint32), ("reserved", C.c_uint32)]


class ClStats(C.Structure):
    """vgb_cl_stats"""
    _fields_ = [(k, C.c_uint32) for k in ("commands", "paths", "draws", "culled", "zero_alpha", "unsupported", "state_depth", "nested", "scissor_changes")]


def make_cl_context(xform=(1, 0, 0, 1, 0, 0), global_alpha=1.0, tess_tolerance=0.25, canvas=(1280.0, 720.0), scissor=None, flags=0,
                    lists=None, max_depth=0):
    """lists: {handle index: (bytes, flags)} -- the context's command lists a SubmitCommandList may name"""
    c = ClContext()
    if lists:
        n = max(lists) + 1
        arr = (ClList * n)()
        keep = []
        for i, (buf, fl) in lists.items():
            a = np.frombuffer(bytes(buf), np.uint8)
            keep.append(a)
            arr[i].commands = a.ctypes.data if len(a) else None
            arr[i].size, arr[i].flags = len(a), fl
        c.num_lists, c.lists = n, arr
        c._keep = (arr, keep)  # the C struct only borrows the buffers
    c.max_depth = max_depth
    for i, v in enumerate(xform):
        c.xform[i] = v
    sc = scissor if scissor is not None else (0.0, 0.0, canvas[0], canvas[1])
    for i, v in enumerate(sc):
        c.scissor[i] = v
    c.global_alpha, c.tess_tolerance, c.canvas_width, c.canvas_height, c.flags = global_alpha, tess_tolerance, canvas[0], canvas[1], flags
    return c


class SubPath(C.Structure):
    _fields_ = [("first_vertex", C.c_uint32), ("num_vertices", C.c_uint32), ("is_closed", C.c_uint8)]


class Merged(C.Structure):
    """vgb_merged: this rank's place in the group's merged vertex / index space"""
    _fields_ = [("vertex_base", C.c_uint64), ("index_base", C.c_uint64), ("total_vertices", C.c_uint64), ("total_indices", C.c_uint64)]


class Mesh(C.Structure):
    _fields_ = [("pos", C.POINTER(C.c_float)), ("color", C.POINTER(C.c_uint32)), ("index", C.POINTER(C.c_uint16)),
                ("num_vertices", C.c_uint32), ("num_indices", C.c_uint32)]


_lib = None


def load_library():
    """dlopen libvgb200.so once; declares argument types for the entry points used here."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} not built: run __graft_entry__.build() / make -C vg_renderer_b200/csrc")
    L = C.CDLL(p)
    f, u32, u64, vp, i = C.c_float, C.c_uint32, C.c_uint64, C.c_void_p, C.c_int
    L.vgb_last_error.restype = C.c_char_p
    L.vgb_version.restype = C.c_char_p
    L.vgb_batch_create.restype = vp
    L.vgb_batch_create.argtypes = [i]
    L.vgb_batch_destroy.argtypes = [vp]
    L.vgb_batch_reset.argtypes = [vp, f]
    L.vgb_batch_begin_path.argtypes = [vp, f, f, vp]
    L.vgb_batch_move_to.argtypes = [vp, f, f]
    L.vgb_batch_line_to.argtypes = [vp, f, f]
    L.vgb_batch_cubic_to.argtypes = [vp] + [f] * 6
    L.vgb_batch_quadratic_to.argtypes = [vp] + [f] * 4
    L.vgb_batch_arc_to.argtypes = [vp] + [f] * 5
    L.vgb_batch_rect.argtypes = [vp] + [f] * 4
    L.vgb_batch_rounded_rect.argtypes = [vp] + [f] * 5
    L.vgb_batch_rounded_rect_varying.argtypes = [vp] + [f] * 8
    L.vgb_batch_circle.argtypes = [vp] + [f] * 3
    L.vgb_batch_ellipse.argtypes = [vp] + [f] * 4
    L.vgb_batch_arc.argtypes = [vp] + [f] * 5 + [u32]
    L.vgb_batch_polyline.argtypes = [vp, vp, u32]
    L.vgb_batch_close.argtypes = [vp]
    L.vgb_batch_stroke.argtypes = [vp, u32, u32, f, u32, u32]
    L.vgb_batch_convex_fill.argtypes = [vp, u32, i]
    L.vgb_batch_add_cubic_strokes.argtypes = [vp, vp, u32, f, f, u32, u32, f, u32, u32]
    L.vgb_batch_add_rounded_rect_fills.argtypes = [vp, vp, u32, f, f, u32, i]
    L.vgb_batch_add_paths.argtypes = [vp, u32] + [vp] * 13
    L.vgb_batch_set_cubic_strokes.argtypes = [vp, vp, u32, f, f, u32, u32, f, u32, u32]
    L.vgb_batch_set_polylines.argtypes = [vp, vp, u32, vp, vp, vp, u32, f, f]
    L.vgb_batch_draw_polyline.argtypes = [vp, u32, u32, u32, f, u32, u32]
    L.vgb_batch_draw_all.argtypes = [vp, u32, u32, f, u32, u32]
    L.vgb_batch_recorded_counts.argtypes = [vp, vp]
    L.vgb_batch_recorded_subpaths.argtypes = [vp, vp, vp, vp, vp]
    L.vgb_batch_recorded_jobs.argtypes = [vp, vp, vp]
    L.vgb_batch_recorded_streams.argtypes = [vp] + [vp] * 11
    L.vgb_batch_set_path_xform.argtypes = [vp, vp]
    L.vgb_batch_set_scissor.argtypes = [vp, vp, C.c_int]
    L.vgb_batch_scissor_rects.argtypes = [vp, vp, u32, vp]
    L.vgb_batch_recorded_draw_scissor.argtypes = [vp, vp]
    L.vgb_batch_fringe.restype = f
    L.vgb_batch_fringe.argtypes = [vp]
    L.vgb_batch_submit_command_list.argtypes = [vp, vp, u32, C.POINTER(ClContext), C.POINTER(ClStats)]
    L.vgb_batch_upload.argtypes = [vp]
    L.vgb_batch_run.argtypes = [vp, u32, u64]
    L.vgb_batch_sync.argtypes = [vp]
    L.vgb_batch_get_sizes.argtypes = [vp, C.POINTER(Sizes)]
    L.vgb_batch_device_ptrs.argtypes = [vp] + [C.POINTER(vp)] * 7
    L.vgb_batch_download.argtypes = [vp, vp, vp, vp, u64, u64, vp, vp]
    L.vgb_batch_download_paths.argtypes = [vp, vp, u64, vp, u32]
    L.vgb_batch_set_final_layout.argtypes = [vp, C.POINTER(FinalCfg)]
    L.vgb_batch_download_final.argtypes = [vp, vp, u64, vp, u32, C.POINTER(u32), C.POINTER(vp)]
    L.vgb_batch_replay.argtypes = [vp, vp, u32, vp]
    L.vgb_batch_replay_sizes.argtypes = [vp, C.POINTER(Sizes)]
    L.vgb_batch_replay_device_ptrs.argtypes = [vp] + [C.POINTER(vp)] * 3
    L.vgb_batch_download_replay.argtypes = [vp, vp, vp, vp, u64, u64]
    L.vgb_batch_execute.argtypes = [vp, u32, vp, vp, vp, u64, u64, vp, vp, C.POINTER(Sizes)]
    L.vgb_batch_last_run_ms.argtypes = [vp] + [C.POINTER(f)] * 4
    L.vgb_batch_mark.argtypes = [vp, u32]
    L.vgb_batch_elapsed_ms.argtypes = [vp, u32, u32, C.POINTER(f)]
    L.vgb_host_alloc.restype = vp
    L.vgb_host_alloc.argtypes = [C.c_size_t]
    L.vgb_host_free.argtypes = [vp]
    L.vgb_batch_stream.restype = vp
    L.vgb_batch_stream.argtypes = [vp]
    # several GPUs, one merged vertex/index space
    L.vgb_group_unique_id.argtypes = [vp]
    L.vgb_group_create.argtypes = [C.POINTER(vp), i, i, i, vp]
    L.vgb_group_destroy.argtypes = [vp]
    L.vgb_batch_run_merged.argtypes = [vp, vp, u32]
    L.vgb_batch_merged_info.argtypes = [vp, C.POINTER(Merged)]
    L.vgb_group_gather.argtypes = [vp, vp, i] + [C.POINTER(vp)] * 3 + [C.POINTER(Merged), C.POINTER(f)]
    # per-object API
    L.vgb_path_create.restype = vp
    L.vgb_path_create.argtypes = [vp, vp]
    L.vgb_path_destroy.argtypes = [vp]
    L.vgb_path_reset.argtypes = [vp, f, f]
    L.vgb_path_move_to.argtypes = [vp, f, f]
    L.vgb_path_line_to.argtypes = [vp, f, f]
    L.vgb_path_cubic_to.argtypes = [vp] + [f] * 6
    L.vgb_path_quadratic_to.argtypes = [vp] + [f] * 4
    L.vgb_path_arc_to.argtypes = [vp] + [f] * 5
    L.vgb_path_rect.argtypes = [vp] + [f] * 4
    L.vgb_path_rounded_rect.argtypes = [vp] + [f] * 5
    L.vgb_path_rounded_rect_varying.argtypes = [vp] + [f] * 8
    L.vgb_path_circle.argtypes = [vp] + [f] * 3
    L.vgb_path_ellipse.argtypes = [vp] + [f] * 4
    L.vgb_path_arc.argtypes = [vp] + [f] * 5 + [u32]
    L.vgb_path_polyline.argtypes = [vp, vp, u32]
    L.vgb_path_close.argtypes = [vp]
    L.vgb_path_get_vertices.restype = C.POINTER(f)
    L.vgb_path_get_vertices.argtypes = [vp]
    L.vgb_path_get_num_vertices.restype = u32
    L.vgb_path_get_num_vertices.argtypes = [vp]
    L.vgb_path_get_subpaths.restype = C.POINTER(SubPath)
    L.vgb_path_get_subpaths.argtypes = [vp]
    L.vgb_path_get_num_subpaths.restype = u32
    L.vgb_path_get_num_subpaths.argtypes = [vp]
    L.vgb_stroker_create.restype = vp
    L.vgb_stroker_create.argtypes = [vp, vp]
    L.vgb_stroker_destroy.argtypes = [vp]
    L.vgb_stroker_reset.argtypes = [vp, f, f, f]
    L.vgb_stroker_polyline_stroke.argtypes = [vp, C.POINTER(Mesh), vp, u32, i, f, u32, u32]
    L.vgb_stroker_polyline_stroke_aa.argtypes = [vp, C.POINTER(Mesh), vp, u32, i, u32, f, u32, u32]
    L.vgb_stroker_polyline_stroke_aa_thin.argtypes = [vp, C.POINTER(Mesh), vp, u32, i, u32, u32, u32]
    L.vgb_stroker_convex_fill.argtypes = [vp, C.POINTER(Mesh), vp, u32]
    L.vgb_stroker_convex_fill_aa.argtypes = [vp, C.POINTER(Mesh), vp, u32, u32]
    _lib = L
    return L


def device_count():
    return load_library().vgb_device_count()


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Batch:
    """Batched record -> run -> read. One CUDA stream and one set of device buffers per Batch."""

    def __init__(self, device=0, fringe=1.0):
        self.L = load_library()
        self.h = self.L.vgb_batch_create(device)
        if not self.h:
            raise VgbError(-4, self.L.vgb_last_error().decode())
        self.index_u32 = False
        self._chk(self.L.vgb_batch_reset(self.h, fringe))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.vgb_batch_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise VgbError(rc, self.L.vgb_last_error().decode())
        return self

    # ---- recording (names follow vg::beginPath/moveTo/.../strokePath/fillPath) ----
    def reset(self, fringe=1.0): return self._chk(self.L.vgb_batch_reset(self.h, fringe))

    def begin_path(self, scale=1.0, tol=0.25, xform=None):
        m = None if xform is None else (C.c_float * 6)(*[float(v) for v in xform])
        return self._chk(self.L.vgb_batch_begin_path(self.h, scale, tol, C.cast(m, C.c_void_p) if m is not None else None))

    def move_to(self, x, y): return self._chk(self.L.vgb_batch_move_to(self.h, x, y))
    def line_to(self, x, y): return self._chk(self.L.vgb_batch_line_to(self.h, x, y))
    def cubic_to(self, a, b, c, d, x, y): return self._chk(self.L.vgb_batch_cubic_to(self.h, a, b, c, d, x, y))
    def quadratic_to(self, cx, cy, x, y): return self._chk(self.L.vgb_batch_quadratic_to(self.h, cx, cy, x, y))
    def arc_to(self, x1, y1, x2, y2, r): return self._chk(self.L.vgb_batch_arc_to(self.h, x1, y1, x2, y2, r))
    def rect(self, x, y, w, h): return self._chk(self.L.vgb_batch_rect(self.h, x, y, w, h))
    def rounded_rect(self, x, y, w, h, r): return self._chk(self.L.vgb_batch_rounded_rect(self.h, x, y, w, h, r))
    def rounded_rect_varying(self, x, y, w, h, a, b, c, d): return self._chk(self.L.vgb_batch_rounded_rect_varying(self.h, x, y, w, h, a, b, c, d))
    def circle(self, x, y, r): return self._chk(self.L.vgb_batch_circle(self.h, x, y, r))
    def ellipse(self, x, y, rx, ry): return self._chk(self.L.vgb_batch_ellipse(self.h, x, y, rx, ry))
    def arc(self, x, y, r, a0, a1, dir_): return self._chk(self.L.vgb_batch_arc(self.h, x, y, r, a0, a1, dir_))
    def close(self): return self._chk(self.L.vgb_batch_close(self.h))

    def polyline(self, coords):
        c = np.ascontiguousarray(coords, np.float32).reshape(-1)
        return self._chk(self.L.vgb_batch_polyline(self.h, _ptr(c), len(c) // 2))

    def stroke(self, op=OP_STROKE_AA, color=0xFF0000FF, width=4.0, cap=CAP_BUTT, join=JOIN_MITER):
        return self._chk(self.L.vgb_batch_stroke(self.h, op, color, width, cap, join))

    def convex_fill(self, color=0xFF00FF00, aa=True):
        return self._chk(self.L.vgb_batch_convex_fill(self.h, color, int(aa)))

    def add_cubic_strokes(self, p8, scale=1.0, tol=0.25, op=OP_STROKE_AA, color=0xFF0000FF, width=4.0, cap=CAP_ROUND, join=JOIN_ROUND):
        p8 = np.ascontiguousarray(p8, np.float32).reshape(-1, 8)
        return self._chk(self.L.vgb_batch_add_cubic_strokes(self.h, _ptr(p8), len(p8), scale, tol, op, color, width, cap, join))

    def add_tables(self, t):
        """ingest a tables.BatchTables (numpy) in one call: vgb_batch_add_paths"""
        return self._chk(self.L.vgb_batch_add_paths(self.h, t.num_paths, _ptr(t.path_cmd_off), _ptr(t.path_data_off), _ptr(t.cmd), _ptr(t.data),
                                                     _ptr(t.path_scale), _ptr(t.path_tol), _ptr(t.path_xform), _ptr(t.draw_begin), _ptr(t.draw_op),
                                                     _ptr(t.draw_color), _ptr(t.draw_width), _ptr(t.draw_cap), _ptr(t.draw_join)))

    def set_cubic_strokes(self, p8, scale=1.0, tol=0.25, op=OP_STROKE_AA, color=0xFF0000FF, width=4.0, cap=CAP_ROUND, join=JOIN_ROUND):
        """zero-copy bulk input: p8 (float32 [n,8], ideally pinned) is BORROWED until the next upload/run/execute"""
        assert p8.dtype == np.float32 and p8.flags["C_CONTIGUOUS"]
        self._bulk_ref = p8  # keep the caller's buffer alive
        return self._chk(self.L.vgb_batch_set_cubic_strokes(self.h, _ptr(p8), p8.size // 8, scale, tol, op, color, width, cap, join))

    def add_rounded_rect_fills(self, xywhr, scale=1.0, tol=0.25, color=0xFF00FF00, aa=True):
        q = np.ascontiguousarray(xywhr, np.float32).reshape(-1, 5)
        return self._chk(self.L.vgb_batch_add_rounded_rect_fills(self.h, _ptr(q), len(q), scale, tol, color, int(aa)))

    def set_polylines(self, pts, first, num, closed, scale=1.0, tol=0.25):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        first = np.ascontiguousarray(first, np.uint32)
        num = np.ascontiguousarray(num, np.uint32)
        closed = np.ascontiguousarray(closed, np.uint8)
        return self._chk(self.L.vgb_batch_set_polylines(self.h, _ptr(pts), len(pts), _ptr(first), _ptr(num), _ptr(closed), len(first), scale, tol))

    def draw_polyline(self, sub, op=OP_STROKE_AA, color=0xFF0000FF, width=4.0, cap=CAP_BUTT, join=JOIN_MITER):
        return self._chk(self.L.vgb_batch_draw_polyline(self.h, sub, op, color, width, cap, join))

    def draw_all(self, op=OP_STROKE_AA, color=0xFF0000FF, width=4.0, cap=CAP_BUTT, join=JOIN_MITER):
        return self._chk(self.L.vgb_batch_draw_all(self.h, op, color, width, cap, join))

    def recorded(self):
        """host tables as recorded (works on a record-only batch, device=-1)"""
        c = np.zeros(6, np.uint32)
        self._chk(self.L.vgb_batch_recorded_counts(self.h, _ptr(c)))
        ns, nj = int(c[1]), int(c[5])
        a = [np.zeros(max(ns, 1), np.uint32) for _ in range(4)]
        self._chk(self.L.vgb_batch_recorded_subpaths(self.h, *[_ptr(x) for x in a]))
        js, jd = np.zeros(max(nj, 1), np.uint32), np.zeros(max(nj, 1), np.uint32)
        self._chk(self.L.vgb_batch_recorded_jobs(self.h, _ptr(js), _ptr(jd)))
        return dict(num_paths=int(c[0]), num_subpaths=ns, num_cmds=int(c[2]), num_data=int(c[3]), num_draws=int(c[4]), num_jobs=nj,
                    cmd_begin=a[0][:ns], cmd_end=a[1][:ns], data_begin=a[2][:ns], path_id=a[3][:ns], job_subpath=js[:nj], job_draw=jd[:nj])

    def recorded_streams(self):
        """everything recorded, as arrays: cmd, data, path_scale, path_tol, path_xform[P,6], draw_{op,color,width,cap,join,path}"""
        r = self.recorded()
        P, D = r["num_paths"], r["num_draws"]
        out = dict(cmd=np.zeros(max(r["num_cmds"], 1), np.uint8), data=np.zeros(max(r["num_data"], 1), np.float32),
                   path_scale=np.zeros(max(P, 1), np.float32), path_tol=np.zeros(max(P, 1), np.float32),
                   path_xform=np.zeros((max(P, 1), 6), np.float32), draw_op=np.zeros(max(D, 1), np.uint32),
                   draw_color=np.zeros(max(D, 1), np.uint32), draw_width=np.zeros(max(D, 1), np.float32),
                   draw_cap=np.zeros(max(D, 1), np.uint8), draw_join=np.zeros(max(D, 1), np.uint8), draw_path=np.zeros(max(D, 1), np.uint32))
        self._chk(self.L.vgb_batch_recorded_streams(self.h, *[_ptr(out[k]) for k in (
            "cmd", "data", "path_scale", "path_tol", "path_xform", "draw_op", "draw_color", "draw_width", "draw_cap", "draw_join", "draw_path")]))
        n = dict(cmd=r["num_cmds"], data=r["num_data"], path_scale=P, path_tol=P, path_xform=P)
        return {k: v[: n.get(k, D)] for k, v in out.items()} | r

    def set_scissor(self, rect, force_new_command=True):
        """scissor {x, y, w, h} of the draws recorded from now on (vg.cpp State::m_ScissorRect); force_new_command
        mirrors m_ForceNewDrawCommand: the next non-empty mesh opens a draw command in the final layout"""
        r = np.ascontiguousarray(rect, np.float32).reshape(4)
        self._chk(self.L.vgb_batch_set_scissor(self.h, _ptr(r), 1 if force_new_command else 0))

    def scissor_rects(self):
        """the scissor table vgb_draw_cmd::scissor_id indexes: [n,4] uint16 {x, y, w, h}; row 0 = no scissor set"""
        n = C.c_uint32()
        self._chk(self.L.vgb_batch_scissor_rects(self.h, None, 0, C.byref(n)))
        out = np.zeros((n.value, 4), np.uint16)
        self._chk(self.L.vgb_batch_scissor_rects(self.h, _ptr(out), n.value, C.byref(n)))
        return out

    def recorded_draw_scissor(self):
        """per recorded draw: scissor id | 1 << 31 when the draw opens a new draw command"""
        d = self.recorded()["num_draws"]
        out = np.zeros(max(d, 1), np.uint32)
        self._chk(self.L.vgb_batch_recorded_draw_scissor(self.h, _ptr(out)))
        return out[:d]

    def recorded_mesh_scissor(self):
        """the same per mesh (= job), the mark on the first mesh of its draw: what oracle final_layout(mesh_scissor=) takes"""
        jd = self.recorded()["job_draw"]
        ds = self.recorded_draw_scissor()
        if len(jd) == 0:
            return np.zeros(0, np.uint32)
        ms = ds[jd].copy()
        first = np.ones(len(jd), bool)
        first[1:] = jd[1:] != jd[:-1]
        ms[~first] &= np.uint32(0x7FFFFFFF)
        return ms

    def set_path_xform(self, xform6):
        m = np.ascontiguousarray(xform6, np.float32).reshape(6)
        return self._chk(self.L.vgb_batch_set_path_xform(self.h, _ptr(m)))

    def submit_command_list(self, buf, ctx):
        """decode a vg-renderer CommandList byte stream (bytes / uint8 array) into the batch; ctx: ClContext (updated in place)"""
        a = np.frombuffer(bytes(buf), np.uint8) if not isinstance(buf, np.ndarray) else np.ascontiguousarray(buf, np.uint8)
        st = ClStats()
        self._chk(self.L.vgb_batch_submit_command_list(self.h, _ptr(a) if len(a) else None, len(a), C.byref(ctx), C.byref(st)))
        return st

    # ---- execution ----
    def upload(self): return self._chk(self.L.vgb_batch_upload(self.h))

    def run(self, index_u32=False, vertex_base=0, fill_order_sse=False):
        """fill_order_sse: strokerConvexFillAA indices in the order of the reference's SIMD build (VGB_RUN_FILL_ORDER_SSE)"""
        self.index_u32 = bool(index_u32)
        return self._chk(self.L.vgb_batch_run(self.h, (INDEX_U32 if index_u32 else INDEX_U16) | (FILL_ORDER_SSE if fill_order_sse else 0), vertex_base))

    def run_merged(self, group):
        """run as one rank of `group`: merged uint32 indices, base from the in-library all-gather (asynchronous)"""
        self.index_u32 = True
        return self._chk(self.L.vgb_batch_run_merged(self.h, group.h, INDEX_U32))

    def merged_info(self):
        m = Merged()
        self._chk(self.L.vgb_batch_merged_info(self.h, C.byref(m)))
        return dict(vertex_base=int(m.vertex_base), index_base=int(m.index_base), total_vertices=int(m.total_vertices),
                    total_indices=int(m.total_indices))

    def sync(self): return self._chk(self.L.vgb_batch_sync(self.h))

    def sizes(self):
        s = Sizes()
        self._chk(self.L.vgb_batch_get_sizes(self.h, C.byref(s)))
        return s

    def last_run_ms(self):
        v = [C.c_float() for _ in range(4)]
        self._chk(self.L.vgb_batch_last_run_ms(self.h, *[C.byref(x) for x in v]))
        return dict(total=v[0].value, flatten=v[1].value, stroke_count=v[2].value, emit=v[3].value)

    def mark(self, slot): return self._chk(self.L.vgb_batch_mark(self.h, slot))

    def elapsed_ms(self, a, b):
        v = C.c_float()
        self._chk(self.L.vgb_batch_elapsed_ms(self.h, a, b, C.byref(v)))
        return v.value

    def device_ptrs(self):
        p = [C.c_void_p() for _ in range(7)]
        self._chk(self.L.vgb_batch_device_ptrs(self.h, *[C.byref(x) for x in p]))
        return dict(zip(("pos", "color", "index", "mesh_vtx_off", "mesh_idx_off", "points", "subpaths"), [x.value for x in p]))

    def download(self, want_mesh_table=True):
        """-> dict(pos[V,2] f32, color[V] u32, index[I] u16|u32, mesh_vtx_off, mesh_idx_off)"""
        s = self.sizes()
        nv, ni, nm = int(s.num_vertices), int(s.num_indices), int(s.num_meshes)
        pos = np.empty((nv, 2), np.float32)
        col = np.empty(nv, np.uint32)
        idx = np.empty(ni, np.uint32 if self.index_u32 else np.uint16)
        mvo = np.zeros(nm + 1, np.uint64) if want_mesh_table else None
        mio = np.zeros(nm + 1, np.uint64) if want_mesh_table else None
        self._chk(self.L.vgb_batch_download(self.h, _ptr(pos), _ptr(col), _ptr(idx), nv, ni, _ptr(mvo), _ptr(mio)))
        return dict(pos=pos, color=col, index=idx, mesh_vtx_off=mvo, mesh_idx_off=mio,
                    num_vertices=nv, num_indices=ni, num_meshes=nm, num_points=int(s.num_points))

    def set_final_layout(self, cfg):
        """cfg: FinalCfg (make_final_cfg) or None.  The next uint16 runs produce vg.cpp's merged layout."""
        self._chk(self.L.vgb_batch_set_final_layout(self.h, C.byref(cfg) if cfg is not None else None))

    def download_final(self, want_uv=True):
        """-> (uv[V,2] int16 | None, cmds: structured array of DRAW_CMD_DTYPE)"""
        s = self.sizes()
        nv = int(s.num_vertices)
        n = C.c_uint32()
        self._chk(self.L.vgb_batch_download_final(self.h, None, 0, None, 0, C.byref(n), None))
        cmds = np.zeros(n.value, DRAW_CMD_DTYPE)
        uv = np.empty((nv, 2), np.int16) if want_uv else None
        self._chk(self.L.vgb_batch_download_final(self.h, _ptr(uv), nv, _ptr(cmds), n.value, C.byref(n), None))
        return uv, cmds

    def replay(self, xforms, inv_xform=None):
        """write the meshes of the last run once per row of xforms[n,6] (instance replay)"""
        x = np.ascontiguousarray(xforms, np.float32).reshape(-1, 6)
        inv = np.ascontiguousarray(inv_xform, np.float32).reshape(6) if inv_xform is not None else None
        return self._chk(self.L.vgb_batch_replay(self.h, _ptr(x), len(x), _ptr(inv)))

    def replay_sizes(self):
        s = Sizes()
        self._chk(self.L.vgb_batch_replay_sizes(self.h, C.byref(s)))
        return s

    def download_replay(self):
        s = self.replay_sizes()
        nv, ni = int(s.num_vertices), int(s.num_indices)
        pos = np.empty((nv, 2), np.float32)
        col = np.empty(nv, np.uint32)
        idx = np.empty(ni, np.uint32 if self.index_u32 else np.uint16)
        self._chk(self.L.vgb_batch_download_replay(self.h, _ptr(pos), _ptr(col), _ptr(idx), nv, ni))
        return dict(pos=pos, color=col, index=idx, num_vertices=nv, num_indices=ni, num_meshes=int(s.num_meshes))

    def download_paths(self):
        """-> (pts[T,2] f32, [(first, num, closed), ...])"""
        s = self.sizes()
        pts = np.empty((int(s.num_points), 2), np.float32)
        subs = (SubPath * max(1, s.num_subpaths))()
        self._chk(self.L.vgb_batch_download_paths(self.h, _ptr(pts), len(pts), C.cast(subs, C.c_void_p), s.num_subpaths))
        return pts, [(subs[i].first_vertex, subs[i].num_vertices, bool(subs[i].is_closed)) for i in range(s.num_subpaths)]

    def execute_into(self, pos, col, idx, index_u32=False):
        """upload + run + D2H into caller (pinned) numpy buffers; returns Sizes."""
        self.index_u32 = bool(index_u32)
        s = Sizes()
        self._chk(self.L.vgb_batch_execute(self.h, INDEX_U32 if index_u32 else INDEX_U16, _ptr(pos), _ptr(col), _ptr(idx),
                                           len(pos), len(idx), None, None, C.byref(s)))
        return s


class Group:
    """vgb_group: the ranks (one process per GPU) that share ONE merged vertex / index space.
    Group.unique_id() on one rank -> 128 bytes -> every rank (by the launcher's own means) -> Group(...)."""

    @staticmethod
    def unique_id():
        L = load_library()
        buf = (C.c_uint8 * 128)()
        rc = L.vgb_group_unique_id(C.cast(buf, C.c_void_p))
        if rc != 0:
            raise VgbError(rc, L.vgb_last_error().decode())
        return bytes(buf)

    def __init__(self, device, rank, nranks, uid):
        self.L = load_library()
        self.rank, self.nranks = rank, nranks
        h = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(uid))
        rc = self.L.vgb_group_create(C.byref(h), device, rank, nranks, C.cast(buf, C.c_void_p))
        if rc != 0:
            raise VgbError(rc, self.L.vgb_last_error().decode())
        self.h = h

    def gather(self, batch, owner=0):
        """payload of every rank -> ONE buffer on `owner` (NVLink peer copies).  -> dict(pos/color/index device
        pointers on the owner, else None; info; ms = device time of the transfer)"""
        p = [C.c_void_p() for _ in range(3)]
        m, ms = Merged(), C.c_float()
        rc = self.L.vgb_group_gather(self.h, batch.h, owner, *[C.byref(x) for x in p], C.byref(m), C.byref(ms))
        if rc != 0:
            raise VgbError(rc, self.L.vgb_last_error().decode())
        return dict(pos=p[0].value, color=p[1].value, index=p[2].value, ms=ms.value, vertex_base=int(m.vertex_base), index_base=int(m.index_base),
                    total_vertices=int(m.total_vertices), total_indices=int(m.total_indices))

    def close(self):
        if getattr(self, "h", None):
            self.L.vgb_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def copy_from_device(ptr, nbytes, dtype):
    """device pointer -> numpy (test helper: cudaMemcpy through the library's own runtime)"""
    L = load_library()
    out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype)
    L.vgb_copy_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    rc = L.vgb_copy_to_host(_ptr(out), C.c_void_p(ptr), nbytes)
    if rc != 0:
        raise VgbError(rc, L.vgb_last_error().decode())
    return out


def pinned_empty(shape, dtype):
    """numpy array backed by cudaHostAlloc'd memory (kept alive by the returned array's base)."""
    L = load_library()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = L.vgb_host_alloc(max(n, 16))
    if not p:
        raise MemoryError("cudaHostAlloc failed")
    buf = (C.c_uint8 * max(n, 16)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr, p


class Path:
    """vg::Path (include/vg/path.h:19-38): same call pattern; getters trigger the GPU flatten."""

    def __init__(self):
        self.L = load_library()
        self.h = self.L.vgb_path_create(None, None)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.vgb_path_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def reset(self, scale=1.0, tol=0.25): self.L.vgb_path_reset(self.h, scale, tol)
    def move_to(self, x, y): self.L.vgb_path_move_to(self.h, x, y)
    def line_to(self, x, y): self.L.vgb_path_line_to(self.h, x, y)
    def cubic_to(self, a, b, c, d, x, y): self.L.vgb_path_cubic_to(self.h, a, b, c, d, x, y)
    def quadratic_to(self, cx, cy, x, y): self.L.vgb_path_quadratic_to(self.h, cx, cy, x, y)
    def arc_to(self, x1, y1, x2, y2, r): self.L.vgb_path_arc_to(self.h, x1, y1, x2, y2, r)
    def rect(self, x, y, w, h): self.L.vgb_path_rect(self.h, x, y, w, h)
    def rounded_rect(self, x, y, w, h, r): self.L.vgb_path_rounded_rect(self.h, x, y, w, h, r)
    def rounded_rect_varying(self, x, y, w, h, a, b, c, d): self.L.vgb_path_rounded_rect_varying(self.h, x, y, w, h, a, b, c, d)
    def circle(self, x, y, r): self.L.vgb_path_circle(self.h, x, y, r)
    def ellipse(self, x, y, rx, ry): self.L.vgb_path_ellipse(self.h, x, y, rx, ry)
    def arc(self, x, y, r, a0, a1, dir_): self.L.vgb_path_arc(self.h, x, y, r, a0, a1, dir_)
    def close(self): self.L.vgb_path_close(self.h)

    def polyline(self, coords):
        c = np.ascontiguousarray(coords, np.float32).reshape(-1)
        self.L.vgb_path_polyline(self.h, _ptr(c), len(c) // 2)

    def vertices(self):
        n = self.L.vgb_path_get_num_vertices(self.h)
        p = self.L.vgb_path_get_vertices(self.h)
        return np.ctypeslib.as_array(p, shape=(n, 2)).copy() if n else np.zeros((0, 2), np.float32)

    def subpaths(self):
        n = self.L.vgb_path_get_num_subpaths(self.h)
        s = self.L.vgb_path_get_subpaths(self.h)
        return [(s[i].first_vertex, s[i].num_vertices, bool(s[i].is_closed)) for i in range(n)]


class Stroker:
    """vg::Stroker (include/vg/stroker.h:11-70)."""

    def __init__(self):
        self.L = load_library()
        self.h = self.L.vgb_stroker_create(None, None)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.vgb_stroker_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def reset(self, scale=1.0, tol=0.25, fringe=1.0): self.L.vgb_stroker_reset(self.h, scale, tol, fringe)

    @staticmethod
    def _out(m):
        nv, ni = m.num_vertices, m.num_indices
        pos = np.ctypeslib.as_array(m.pos, shape=(nv, 2)).copy() if nv else np.zeros((0, 2), np.float32)
        col = np.ctypeslib.as_array(m.color, shape=(nv,)).copy() if (nv and m.color) else None
        idx = np.ctypeslib.as_array(m.index, shape=(ni,)).copy() if ni else np.zeros(0, np.uint16)
        return pos, col, idx

    def polyline_stroke_aa(self, vtx, closed, color, width, cap, join):
        v = np.ascontiguousarray(vtx, np.float32).reshape(-1, 2)
        m = Mesh()
        self.L.vgb_stroker_polyline_stroke_aa(self.h, C.byref(m), _ptr(v), len(v), int(closed), color, width, cap, join)
        return self._out(m)

    def polyline_stroke(self, vtx, closed, width, cap, join):
        v = np.ascontiguousarray(vtx, np.float32).reshape(-1, 2)
        m = Mesh()
        self.L.vgb_stroker_polyline_stroke(self.h, C.byref(m), _ptr(v), len(v), int(closed), width, cap, join)
        return self._out(m)

    def polyline_stroke_aa_thin(self, vtx, closed, color, cap, join):
        v = np.ascontiguousarray(vtx, np.float32).reshape(-1, 2)
        m = Mesh()
        self.L.vgb_stroker_polyline_stroke_aa_thin(self.h, C.byref(m), _ptr(v), len(v), int(closed), color, cap, join)
        return self._out(m)

    def convex_fill_aa(self, vtx, color):
        v = np.ascontiguousarray(vtx, np.float32).reshape(-1, 2)
        m = Mesh()
        self.L.vgb_stroker_convex_fill_aa(self.h, C.byref(m), _ptr(v), len(v), color)
        return self._out(m)

    def convex_fill(self, vtx):
        v = np.ascontiguousarray(vtx, np.float32).reshape(-1, 2)
        m = Mesh()
        self.L.vgb_stroker_convex_fill(self.h, C.byref(m), _ptr(v), len(v))
        return self._out(m)
