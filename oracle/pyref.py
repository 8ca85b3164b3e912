"""ctypes loader for the ORACLE side (test infrastructure, never the product path).

  RefLib("scalar"|"sse"|"libm")  -> oracle/_ref/libvgref*.so : the reference's own path.cpp /
                                    stroker.cpp compiled against oracle/bx_shim (oracle/Makefile)
  OracleLib()                     -> oracle/liboracle.so      : this repo's plain-C restatement

Both expose the same three calls: flatten(cmds) / mesh(op, polyline, ...) / run_batch(tables).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# command ids / arities mirror include/vgb.h
CMD_MOVE_TO, CMD_LINE_TO, CMD_CUBIC_TO, CMD_QUAD_TO, CMD_ARC_TO, CMD_RECT, CMD_ROUNDED_RECT, \
    CMD_ROUNDED_RECT_VARYING, CMD_CIRCLE, CMD_ELLIPSE, CMD_ARC, CMD_POLYLINE, CMD_CLOSE = range(13)
OP_STROKE_AA, OP_CONVEX_FILL_AA, OP_STROKE, OP_STROKE_AA_THIN, OP_CONVEX_FILL = range(5)
CAP_BUTT, CAP_ROUND, CAP_SQUARE = 0, 1, 2
JOIN_MITER, JOIN_ROUND, JOIN_BEVEL = 0, 1, 2


class SubPath(C.Structure):
    _fields_ = [("first_vertex", C.c_uint32), ("num_vertices", C.c_uint32), ("is_closed", C.c_uint8)]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class _Lib:
    prefix = "vgref"

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (needs /root/reference for _ref)")
        self.path = path
        self.lib = C.CDLL(path)
        self._flatten = getattr(self.lib, self.prefix + "_flatten")
        self._mesh = getattr(self.lib, self.prefix + "_mesh")
        self._run = getattr(self.lib, self.prefix + "_run_batch")
        for f in (self._flatten, self._mesh, self._run):
            f.restype = C.c_int

    def variant(self):
        f = getattr(self.lib, self.prefix + "_variant")
        f.restype = C.c_uint32
        return f()

    def flatten(self, cmds, data, scale=1.0, tol=0.25, cap=1 << 16):
        cmds = np.ascontiguousarray(cmds, np.uint8)
        data = np.ascontiguousarray(data, np.float32)
        while True:
            pts = np.empty((cap, 2), np.float32)
            sp = (SubPath * 4096)()
            npts, nsp = C.c_uint32(), C.c_uint32()
            rc = self._flatten(_p(cmds, C.c_uint8), C.c_uint32(len(cmds)), _p(data, C.c_float), C.c_float(scale), C.c_float(tol),
                               _p(pts, C.c_float), C.c_uint32(cap), sp, C.c_uint32(4096), C.byref(npts), C.byref(nsp))
            if rc == -5 and cap < (1 << 26):
                cap *= 8
                continue
            if rc != 0:
                raise RuntimeError(f"flatten rc={rc}")
            subs = [(sp[i].first_vertex, sp[i].num_vertices, bool(sp[i].is_closed)) for i in range(nsp.value)]
            return pts[: npts.value].copy(), subs

    def mesh(self, op, vtx, closed=False, color=0xFF0000FF, width=4.0, cap=CAP_BUTT, join=JOIN_MITER,
             scale=1.0, tol=0.25, fringe=1.0):
        vtx = np.ascontiguousarray(vtx, np.float32).reshape(-1, 2)
        vcap, icap = 1 << 17, 1 << 20
        pos = np.empty((vcap, 2), np.float32)
        col = np.zeros(vcap, np.uint32)
        idx = np.empty(icap, np.uint16)
        nv, ni, hc = C.c_uint32(), C.c_uint32(), C.c_int()
        rc = self._mesh(C.c_uint32(op), _p(vtx, C.c_float), C.c_uint32(len(vtx)), C.c_int(int(closed)), C.c_uint32(color),
                        C.c_float(width), C.c_uint32(cap), C.c_uint32(join), C.c_float(scale), C.c_float(tol), C.c_float(fringe),
                        _p(pos, C.c_float), _p(col, C.c_uint32), _p(idx, C.c_uint16), C.c_uint32(vcap), C.c_uint32(icap),
                        C.byref(nv), C.byref(ni), C.byref(hc))
        if rc != 0:
            raise RuntimeError(f"mesh rc={rc}")
        return pos[: nv.value].copy(), (col[: nv.value].copy() if hc.value else None), idx[: ni.value].copy()

    def run_batch(self, t, store=True, path_range=None):
        """t: vg_renderer_b200.tables.BatchTables-like object (numpy arrays). Returns dict."""
        p0, p1 = path_range if path_range is not None else (0, t.num_paths)
        totals = np.zeros(6, np.uint64)
        if store:
            # size by a dry run
            self.run_batch(t, store=False, path_range=(p0, p1))
            tot = self._last_totals
            nv, ni, nm = int(tot[0]), int(tot[1]), int(tot[3])
            pos = np.empty((max(nv, 1), 2), np.float32)
            col = np.empty(max(nv, 1), np.uint32)
            idx = np.empty(max(ni, 1), np.uint16)
            mvo = np.zeros(nm + 1, np.uint64)
            mio = np.zeros(nm + 1, np.uint64)
        else:
            nv = ni = nm = 0
            pos = col = idx = mvo = mio = None
        rc = self._run(C.c_uint32(p0), C.c_uint32(p1),
                       _p(t.path_cmd_off, C.c_uint32), _p(t.path_data_off, C.c_uint32), _p(t.cmd, C.c_uint8), _p(t.data, C.c_float),
                       _p(t.path_scale, C.c_float), _p(t.path_tol, C.c_float), _p(t.path_xform, C.c_float), C.c_float(t.fringe),
                       _p(t.draw_begin, C.c_uint32), _p(t.draw_op, C.c_uint32), _p(t.draw_color, C.c_uint32), _p(t.draw_width, C.c_float),
                       _p(t.draw_cap, C.c_uint8), _p(t.draw_join, C.c_uint8),
                       C.c_int(int(store)), _p(pos, C.c_float), _p(col, C.c_uint32), _p(idx, C.c_uint16),
                       C.c_uint64(nv), C.c_uint64(ni), _p(mvo, C.c_uint64), _p(mio, C.c_uint64), C.c_uint32(nm),
                       _p(totals, C.c_uint64))
        if rc != 0:
            raise RuntimeError(f"run_batch rc={rc}")
        self._last_totals = totals
        out = dict(num_vertices=int(totals[0]), num_indices=int(totals[1]), num_points=int(totals[2]),
                   num_meshes=int(totals[3]), checksum=int(totals[4]), num_subpaths=int(totals[5]))
        if store:
            out.update(pos=pos[:nv], color=col[:nv], index=idx[:ni], mesh_vtx_off=mvo, mesh_idx_off=mio)
        return out


    def final_layout(self, mesh_vtx_off, mesh_idx_off, index_local, cfg, mesh_scissor=None):
        """vg.cpp's merge step over stored meshes.  cfg: vg_renderer_b200.api.FinalCfg.
        mesh_scissor: per mesh, scissor id | 1 << 31 on the first mesh of a draw that follows a scissor command.
        -> (index[I] u16 rebased, uv[V,2] i16, cmds structured array) or raises on an oversize mesh."""
        from vg_renderer_b200.api import DRAW_CMD_DTYPE
        f = getattr(self.lib, self.prefix + "_final_layout_sc")
        f.restype = C.c_int
        mvo = np.ascontiguousarray(mesh_vtx_off, np.uint64)
        mio = np.ascontiguousarray(mesh_idx_off, np.uint64)
        il = np.ascontiguousarray(index_local, np.uint16)
        nm = len(mvo) - 1
        nv, ni = int(mvo[-1] - mvo[0]), int(mio[-1] - mio[0])
        idx = np.zeros(max(ni, 1), np.uint16)
        uv = np.zeros((max(nv, 1), 2), np.int16)
        cmds = np.zeros(nm + 1, DRAW_CMD_DTYPE)
        n = C.c_uint32()
        ms = None if mesh_scissor is None else np.ascontiguousarray(mesh_scissor, np.uint32)
        assert ms is None or len(ms) == nm
        rc = f(_p(mvo, C.c_uint64), _p(mio, C.c_uint64), C.c_uint32(nm), _p(il, C.c_uint16), C.byref(cfg),
               None if ms is None else _p(ms, C.c_uint32), _p(idx, C.c_uint16), _p(uv, C.c_int16), cmds.ctypes.data_as(C.c_void_p), C.c_uint32(nm + 1), C.byref(n))
        if rc != 0:
            raise RuntimeError(f"final_layout rc={rc}")
        return idx[:ni], uv[:nv], cmds[: n.value]


    def transform_positions(self, pos, mtx):
        """vgutil::batchTransformPositions: pos[n,2] x 2x3 matrix (a, b, c, d, e, f) -> [n,2]"""
        f = getattr(self.lib, self.prefix + "_transform_positions")
        f.restype = None
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 2)
        m = np.ascontiguousarray(mtx, np.float32).reshape(6)
        out = np.empty_like(pos)
        f(_p(pos, C.c_float), C.c_uint32(len(pos)), _p(m, C.c_float), _p(out, C.c_float))
        return out


# "dropin": ref_capi.cpp linked against libvgb200.so's vg:: symbols instead of the reference's path.cpp +
# stroker.cpp (the link-level drop-in test; needs a GPU at run time)
_REF_NAMES = {"scalar": "libvgref.so", "sse": "libvgref_sse.so", "libm": "libvgref_libm.so", "dropin": "libvgref_dropin.so"}


class RefLib(_Lib):
    """The compiled reference (oracle/_ref). kind: scalar (golden) | sse | libm."""
    prefix = "vgref"

    def __init__(self, kind="scalar"):
        super().__init__(os.path.join(HERE, "_ref", _REF_NAMES[kind]))

    def touch_all(self):
        """calls every vg::pathXXX / vg::strokerXXX symbol once (ref_capi.cpp:vgref_touch_all) -> 15 counters"""
        out = np.zeros(15, np.uint32)
        f = self.lib.vgref_touch_all
        f.restype = C.c_int
        rc = f(_p(out, C.c_uint32))
        if rc != 0:
            raise RuntimeError(f"vgref_touch_all rc={rc}")
        return out


class OracleLib(_Lib):
    """This repo's plain-C restatement of the algorithm (oracle/vg_oracle.c)."""
    prefix = "vgo"

    def __init__(self):
        super().__init__(os.path.join(HERE, "liboracle.so"))


def have_ref(kind="scalar"):
    return os.path.exists(os.path.join(HERE, "_ref", _REF_NAMES[kind]))


CTX_DRAW_DTYPE = np.dtype([("vb_id", "<u4"), ("first_vertex", "<u4"), ("num_vertices", "<u4"), ("first_index", "<u4"), ("num_indices", "<u4"),
                           ("program", "<u4"), ("scissor", "<u2", 4)])


def have_ctx():
    return os.path.exists(os.path.join(HERE, "_ref", "libvgctx.so"))


class CtxLib:
    """The reference's WHOLE context -- src/vg.cpp itself, compiled against a record-only bgfx stand-in
    (oracle/vgctx_capi.cpp, oracle/bgfx_stub): records a command list through vg::clXXX, submits it between
    vg::begin / vg::end and returns what bgfx was handed."""

    def __init__(self):
        p = os.path.join(HERE, "_ref", "libvgctx.so")
        if not os.path.exists(p):
            raise FileNotFoundError(f"{p} missing: run `make -C oracle ref` (needs /root/reference)")
        self.lib = C.CDLL(p)
        self.lib.vgctx_run_instances.restype = C.c_int

    def run(self, script, list_flags=0, canvas=(1280, 720), max_vb_vertices=65536, vcap=1 << 22, icap=1 << 24, xforms=None, children=()):
        """xforms [n,6]: submit the list n times, each inside pushState / transformMult(row, Pre) / popState (with
        list_flags = 1, Cacheable, the later submissions are the shape cache's replays); adds inv0 to the result"""
        script = np.frombuffer(bytes(script), np.uint8)
        xf = None if xforms is None else np.ascontiguousarray(xforms, np.float32).reshape(-1, 6)
        inv0 = np.zeros(6, np.float32)
        # children: [(bytes, flags)] -> the context's lists 0 .. n-1, which SubmitCommandList records may name
        kb = np.frombuffer(b"".join(bytes(c[0]) for c in children) + b"\0", np.uint8)
        ks = np.asarray([len(bytes(c[0])) for c in children] + [0], np.uint32)
        kf = np.asarray([c[1] for c in children] + [0], np.uint32)
        cl = np.zeros(len(script) + 4096, np.uint8)
        cl_size, ndraws, nvbs, nidx = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
        draws = np.zeros(1 << 16, CTX_DRAW_DTYPE)
        pos = np.zeros((vcap, 2), np.float32)
        uv = np.zeros((vcap, 2), np.int16)
        col = np.zeros(vcap, np.uint32)
        vbc = np.zeros(4096, np.uint32)
        idx = np.zeros(icap, np.uint16)
        rc = self.lib.vgctx_run_instances(_p(script, C.c_uint8), C.c_uint32(len(script)), C.c_uint32(list_flags), C.c_uint32(canvas[0]), C.c_uint32(canvas[1]),
                                C.c_uint32(max_vb_vertices), None if xf is None else _p(xf, C.c_float), C.c_uint32(0 if xf is None else len(xf)),
                                _p(inv0, C.c_float), _p(kb, C.c_uint8), _p(ks, C.c_uint32), _p(kf, C.c_uint32), C.c_uint32(len(children)), _p(cl, C.c_uint8), C.c_uint32(len(cl)), C.byref(cl_size),
                                draws.ctypes.data_as(C.c_void_p), C.c_uint32(len(draws)), C.byref(ndraws),
                                _p(pos, C.c_float), _p(uv, C.c_int16), _p(col, C.c_uint32), C.c_uint64(vcap), _p(vbc, C.c_uint32), C.c_uint32(len(vbc)),
                                C.byref(nvbs), _p(idx, C.c_uint16), C.c_uint64(icap), C.byref(nidx))
        if rc != 0:
            raise RuntimeError(f"vgctx_run rc={rc}")
        counts = vbc[: nvbs.value].copy()
        nv = int(counts.sum())
        return dict(cl_bytes=cl[: cl_size.value].tobytes(), draws=draws[: ndraws.value].copy(), vb_counts=counts, pos=pos[:nv].copy(),
                    uv=uv[:nv].copy(), color=col[:nv].copy(), index=idx[: nidx.value].copy(), inv0=inv0)
