"""Host-side description of a batch as plain numpy tables (the format oracle/pyref.py's
run_batch consumes) plus a Recorder that builds them with the same call names as the reference's
immediate-mode API (vg::moveTo / cubicTo / strokePath ..., src/vg.cpp:1335-1720).  A Recorder can
replay itself into a product Batch, so parity tests feed both sides from one description."""
import numpy as np

(CMD_MOVE_TO, CMD_LINE_TO, CMD_CUBIC_TO, CMD_QUAD_TO, CMD_ARC_TO, CMD_RECT, CMD_ROUNDED_RECT,
 CMD_ROUNDED_RECT_VARYING, CMD_CIRCLE, CMD_ELLIPSE, CMD_ARC, CMD_POLYLINE, CMD_CLOSE) = range(13)
ARITY = [2, 2, 6, 4, 5, 4, 5, 8, 3, 4, 6, -1, 0]


class BatchTables:
    """numpy tables: paths (command ranges + flatten params) and draws sorted by path."""

    def __init__(self, fringe=1.0):
        self.fringe = float(fringe)
        self.num_paths = 0
        self.path_cmd_off = np.zeros(1, np.uint32)
        self.path_data_off = np.zeros(1, np.uint32)
        self.cmd = np.zeros(0, np.uint8)
        self.data = np.zeros(0, np.float32)
        self.path_scale = np.zeros(0, np.float32)
        self.path_tol = np.zeros(0, np.float32)
        self.path_xform = None
        self.draw_begin = np.zeros(1, np.uint32)
        self.draw_op = np.zeros(0, np.uint32)
        self.draw_color = np.zeros(0, np.uint32)
        self.draw_width = np.zeros(0, np.float32)
        self.draw_cap = np.zeros(0, np.uint8)
        self.draw_join = np.zeros(0, np.uint8)

    @staticmethod
    def cubic_strokes(p8, op=0, color=0xFF0000FF, width=4.0, cap=1, join=1, scale=1.0, tol=0.25, fringe=1.0):
        """n paths: moveTo(p[0:2]) cubicTo(p[2:8]) + one stroke each (BASELINE config 3 shape)."""
        p8 = np.ascontiguousarray(p8, np.float32).reshape(-1, 8)
        n = len(p8)
        t = BatchTables(fringe)
        t.num_paths = n
        t.path_cmd_off = (np.arange(n + 1, dtype=np.uint32) * 2)
        t.path_data_off = (np.arange(n + 1, dtype=np.uint32) * 8)
        t.cmd = np.tile(np.array([CMD_MOVE_TO, CMD_CUBIC_TO], np.uint8), n)
        t.data = p8.reshape(-1).copy()
        t.path_scale = np.full(n, scale, np.float32)
        t.path_tol = np.full(n, tol, np.float32)
        t.draw_begin = np.arange(n + 1, dtype=np.uint32)
        t.draw_op = np.full(n, op, np.uint32)
        t.draw_color = np.full(n, color, np.uint32)
        t.draw_width = np.full(n, width, np.float32)
        t.draw_cap = np.full(n, cap, np.uint8)
        t.draw_join = np.full(n, join, np.uint8)
        return t

    @staticmethod
    def rounded_rect_fills(xywhr, color=0xFF00FF00, aa=True, scale=1.0, tol=0.25, fringe=1.0):
        """n paths: roundedRect(x,y,w,h,r) + convex fill (BASELINE config 4 shape)."""
        q = np.ascontiguousarray(xywhr, np.float32).reshape(-1, 5)
        n = len(q)
        t = BatchTables(fringe)
        t.num_paths = n
        t.path_cmd_off = np.arange(n + 1, dtype=np.uint32)
        t.path_data_off = (np.arange(n + 1, dtype=np.uint32) * 5)
        t.cmd = np.full(n, CMD_ROUNDED_RECT, np.uint8)
        t.data = q.reshape(-1).copy()
        t.path_scale = np.full(n, scale, np.float32)
        t.path_tol = np.full(n, tol, np.float32)
        t.draw_begin = np.arange(n + 1, dtype=np.uint32)
        t.draw_op = np.full(n, 1 if aa else 4, np.uint32)
        t.draw_color = np.full(n, color, np.uint32)
        t.draw_width = np.zeros(n, np.float32)
        t.draw_cap = np.zeros(n, np.uint8)
        t.draw_join = np.zeros(n, np.uint8)
        return t


def tables_from_recorded(r, fringe=1.0):
    """BatchTables out of what a (record-only) vg_renderer_b200.api.Batch recorded: Batch.recorded_streams().
    Lets the compiled reference run exactly what the product's recorder / command-list decoder produced."""
    t = BatchTables(fringe)
    P = int(r["num_paths"])
    t.num_paths = P
    ncmd, ndata = len(r["cmd"]), len(r["data"])
    coff = np.full(P + 1, ncmd, np.int64)
    doff = np.full(P + 1, ndata, np.int64)
    for s in range(int(r["num_subpaths"]) - 1, -1, -1):  # the first sub-path of a path gives the path's start
        p = int(r["path_id"][s])
        coff[p] = int(r["cmd_begin"][s])
        doff[p] = int(r["data_begin"][s])
    for p in range(P - 1, -1, -1):  # a path without commands starts where the next one does
        if coff[p] > coff[p + 1]:
            coff[p], doff[p] = coff[p + 1], doff[p + 1]
    t.path_cmd_off = coff.astype(np.uint32)
    t.path_data_off = doff.astype(np.uint32)
    t.cmd = np.asarray(r["cmd"], np.uint8)
    t.data = np.asarray(r["data"], np.float32)
    t.path_scale = np.asarray(r["path_scale"], np.float32)
    t.path_tol = np.asarray(r["path_tol"], np.float32)
    t.path_xform = np.asarray(r["path_xform"], np.float32).reshape(-1)
    dp = np.asarray(r["draw_path"], np.uint32)
    t.draw_begin = np.searchsorted(dp, np.arange(P + 1), side="left").astype(np.uint32)
    t.draw_op = np.asarray(r["draw_op"], np.uint32)
    t.draw_color = np.asarray(r["draw_color"], np.uint32)
    t.draw_width = np.asarray(r["draw_width"], np.float32)
    t.draw_cap = np.asarray(r["draw_cap"], np.uint8)
    t.draw_join = np.asarray(r["draw_join"], np.uint8)
    return t


class Recorder:
    """Immediate-mode style recorder producing BatchTables (and replayable into a Batch)."""

    def __init__(self, fringe=1.0):
        self.fringe = float(fringe)
        self.paths = []  # dict(scale, tol, xform, cmds=[(c, [floats])], draws=[(op,color,width,cap,join)])

    # --- path building (vg::beginPath .. vg::closePath) ---
    def begin_path(self, scale=1.0, tol=0.25, xform=None):
        self.paths.append(dict(scale=scale, tol=tol, xform=None if xform is None else [float(v) for v in xform], cmds=[], draws=[]))
        return self

    def _c(self, c, *v):
        self.paths[-1]["cmds"].append((c, [float(x) for x in v]))
        return self

    def move_to(self, x, y): return self._c(CMD_MOVE_TO, x, y)
    def line_to(self, x, y): return self._c(CMD_LINE_TO, x, y)
    def cubic_to(self, a, b, c, d, x, y): return self._c(CMD_CUBIC_TO, a, b, c, d, x, y)
    def quadratic_to(self, cx, cy, x, y): return self._c(CMD_QUAD_TO, cx, cy, x, y)
    def arc_to(self, x1, y1, x2, y2, r): return self._c(CMD_ARC_TO, x1, y1, x2, y2, r)
    def rect(self, x, y, w, h): return self._c(CMD_RECT, x, y, w, h)
    def rounded_rect(self, x, y, w, h, r): return self._c(CMD_ROUNDED_RECT, x, y, w, h, r)
    def rounded_rect_varying(self, x, y, w, h, a, b, c, d): return self._c(CMD_ROUNDED_RECT_VARYING, x, y, w, h, a, b, c, d)
    def circle(self, x, y, r): return self._c(CMD_CIRCLE, x, y, r)
    def ellipse(self, x, y, rx, ry): return self._c(CMD_ELLIPSE, x, y, rx, ry)
    def arc(self, x, y, r, a0, a1, cw): return self._c(CMD_ARC, x, y, r, a0, a1, 1.0 if cw else 0.0)
    def close(self): return self._c(CMD_CLOSE)

    def polyline(self, coords):
        coords = np.asarray(coords, np.float32).reshape(-1)
        self.paths[-1]["cmds"].append((CMD_POLYLINE, coords))
        return self

    # --- draws (vg::strokePath / vg::fillPath) ---
    def stroke(self, op=0, color=0xFF0000FF, width=4.0, cap=0, join=0):
        self.paths[-1]["draws"].append((op, color, float(width), cap, join))
        return self

    def convex_fill(self, color=0xFF00FF00, aa=True):
        self.paths[-1]["draws"].append((1 if aa else 4, color, 0.0, 0, 0))
        return self

    # --- outputs ---
    def tables(self):
        t = BatchTables(self.fringe)
        cmd, data, coff, doff = [], [], [0], [0]
        dbeg, dop, dcol, dw, dcap, djoin = [0], [], [], [], [], []
        anyx = any(p["xform"] is not None for p in self.paths)
        xf = []
        for p in self.paths:
            for c, v in p["cmds"]:
                cmd.append(c)
                if c == CMD_POLYLINE:
                    n = len(v) // 2
                    data.append(np.array([n], np.uint32).view(np.float32))
                    data.append(np.asarray(v, np.float32))
                else:
                    data.append(np.asarray(v, np.float32))
            coff.append(len(cmd))
            doff.append(int(sum(len(d) for d in data)))
            for d in p["draws"]:
                dop.append(d[0]); dcol.append(d[1]); dw.append(d[2]); dcap.append(d[3]); djoin.append(d[4])
            dbeg.append(len(dop))
            xf.append(p["xform"] if p["xform"] is not None else [1, 0, 0, 1, 0, 0])
        t.num_paths = len(self.paths)
        t.path_cmd_off = np.array(coff, np.uint32)
        t.path_data_off = np.array(doff, np.uint32)
        t.cmd = np.array(cmd, np.uint8)
        t.data = np.concatenate(data).astype(np.float32) if data else np.zeros(0, np.float32)
        t.path_scale = np.array([p["scale"] for p in self.paths], np.float32)
        t.path_tol = np.array([p["tol"] for p in self.paths], np.float32)
        t.path_xform = np.array(xf, np.float32).reshape(-1) if anyx else None
        t.draw_begin = np.array(dbeg, np.uint32)
        t.draw_op = np.array(dop, np.uint32)
        t.draw_color = np.array(dcol, np.uint32)
        t.draw_width = np.array(dw, np.float32)
        t.draw_cap = np.array(dcap, np.uint8)
        t.draw_join = np.array(djoin, np.uint8)
        return t

    def replay(self, batch):
        """issue the same calls on a product Batch (vg_renderer_b200.api.Batch)."""
        batch.reset(self.fringe)
        for p in self.paths:
            batch.begin_path(p["scale"], p["tol"], p["xform"])
            for c, v in p["cmds"]:
                if c == CMD_MOVE_TO: batch.move_to(*v)
                elif c == CMD_LINE_TO: batch.line_to(*v)
                elif c == CMD_CUBIC_TO: batch.cubic_to(*v)
                elif c == CMD_QUAD_TO: batch.quadratic_to(*v)
                elif c == CMD_ARC_TO: batch.arc_to(*v)
                elif c == CMD_RECT: batch.rect(*v)
                elif c == CMD_ROUNDED_RECT: batch.rounded_rect(*v)
                elif c == CMD_ROUNDED_RECT_VARYING: batch.rounded_rect_varying(*v)
                elif c == CMD_CIRCLE: batch.circle(*v)
                elif c == CMD_ELLIPSE: batch.ellipse(*v)
                elif c == CMD_ARC: batch.arc(v[0], v[1], v[2], v[3], v[4], 1 if v[5] != 0 else 0)
                elif c == CMD_POLYLINE: batch.polyline(v)
                elif c == CMD_CLOSE: batch.close()
            for op, color, width, cap, join in p["draws"]:
                if op in (1, 4):
                    batch.convex_fill(color, op == 1)
                else:
                    batch.stroke(op, color, width, cap, join)
        return batch
