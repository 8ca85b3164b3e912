"""Test-side writer of vg-renderer's CommandList byte format and an independent float32 model of the
state machine that executes it -- both written from the reference's source, not from the decoder:
  writer: clAllocCommand (vg.cpp:5694-5723) + the CMD_WRITE sequences of the cl* functions (vg.cpp:2402-2960)
  model : ctxSubmitCommandList (vg.cpp:4273-4637) with ctxBeginPath / ctxFillPathColor / ctxStrokePathColor /
          the state functions (vg.cpp:2969-2981, 3061-3132, 3401-3492, 3934-4127, 4927-4944).
The model drives a tables.Recorder with explicit (scale, tol, transform, colour, width) per path / draw;
the product decodes the bytes itself (vgb_batch_submit_command_list); tests compare the two recordings."""
import struct

import numpy as np

from vg_renderer_b200.tables import Recorder

F = np.float32
(BEGIN_PATH, MOVE_TO, LINE_TO, CUBIC_TO, QUAD_TO, ARC_TO, ARC, RECT, ROUNDED_RECT, ROUNDED_RECT_VARYING, CIRCLE, ELLIPSE, POLYLINE,
 CLOSE_PATH, FILL_COLOR, FILL_GRADIENT, FILL_IMAGE, STROKE_COLOR, STROKE_GRADIENT, STROKE_IMAGE, TRI_LIST, BEGIN_CLIP, END_CLIP,
 RESET_CLIP, LINEAR_GRADIENT, BOX_GRADIENT, RADIAL_GRADIENT, IMAGE_PATTERN, PUSH_STATE, POP_STATE, RESET_SCISSOR, SET_SCISSOR,
 INTERSECT_SCISSOR, T_IDENTITY, T_SCALE, T_TRANSLATE, T_ROTATE, T_MULT, SET_VIEW_BOX, SET_GLOBAL_ALPHA, TEXT, TEXT_BOX,
 SUBMIT_LIST) = range(43)


def stroke_flags(cap, join, aa, fixed=False):  # VG_STROKE_FLAGS + StrokeFlags::FixedWidth, vg.h:176-208
    return (int(aa) << 4) | (cap << 2) | join | ((1 << 5) if fixed else 0)


def fill_flags(concave=False, aa=True, evenodd=False):  # VG_FILL_FLAGS, vg.h:229
    return (int(evenodd) << 4) | (int(aa) << 2) | int(concave)


class Writer:
    def __init__(self):
        self.buf = bytearray()
        self.ops = []  # (type, payload tuple) for the model

    def _cmd(self, t, payload=b""):
        pad = (-len(payload)) % 16
        self.buf += struct.pack("<II", t, len(payload) + pad) + b"\0" * 8 + payload + b"\0" * pad

    def _f(self, t, *v):
        self.ops.append((t, tuple(float(F(x)) for x in v)))
        self._cmd(t, struct.pack("<%df" % len(v), *v))
        return self

    def begin_path(self): self.ops.append((BEGIN_PATH, ())); self._cmd(BEGIN_PATH); return self
    def close_path(self): self.ops.append((CLOSE_PATH, ())); self._cmd(CLOSE_PATH); return self
    def move_to(self, x, y): return self._f(MOVE_TO, x, y)
    def line_to(self, x, y): return self._f(LINE_TO, x, y)
    def cubic_to(self, *v): return self._f(CUBIC_TO, *v)
    def quadratic_to(self, *v): return self._f(QUAD_TO, *v)
    def arc_to(self, *v): return self._f(ARC_TO, *v)
    def rect(self, *v): return self._f(RECT, *v)
    def rounded_rect(self, *v): return self._f(ROUNDED_RECT, *v)
    def rounded_rect_varying(self, *v): return self._f(ROUNDED_RECT_VARYING, *v)
    def circle(self, *v): return self._f(CIRCLE, *v)
    def ellipse(self, *v): return self._f(ELLIPSE, *v)

    def arc(self, cx, cy, r, a0, a1, cw):
        self.ops.append((ARC, tuple(float(F(x)) for x in (cx, cy, r, a0, a1)) + (int(cw),)))
        self._cmd(ARC, struct.pack("<5fI", cx, cy, r, a0, a1, int(cw)))
        return self

    def polyline(self, coords):
        c = np.asarray(coords, np.float32).reshape(-1)
        self.ops.append((POLYLINE, tuple(float(x) for x in c)))
        self._cmd(POLYLINE, struct.pack("<I", len(c) // 2) + c.tobytes())
        return self

    def fill(self, color, flags):
        self.ops.append((FILL_COLOR, (flags, color)))
        self._cmd(FILL_COLOR, struct.pack("<II", flags, color))
        return self

    def stroke(self, color, width, flags):
        self.ops.append((STROKE_COLOR, (float(F(width)), flags, color)))
        self._cmd(STROKE_COLOR, struct.pack("<fII", width, flags, color))
        return self

    def raw(self, t, payload=b""):  # commands outside the hot path (gradients, text, clip ...)
        self.ops.append((t, payload))
        self._cmd(t, payload)
        return self

    def push_state(self): self.ops.append((PUSH_STATE, ())); self._cmd(PUSH_STATE); return self
    def pop_state(self): self.ops.append((POP_STATE, ())); self._cmd(POP_STATE); return self
    def reset_scissor(self): self.ops.append((RESET_SCISSOR, ())); self._cmd(RESET_SCISSOR); return self
    def set_scissor(self, *v): return self._f(SET_SCISSOR, *v)
    def intersect_scissor(self, *v): return self._f(INTERSECT_SCISSOR, *v)
    def identity(self): self.ops.append((T_IDENTITY, ())); self._cmd(T_IDENTITY); return self
    def scale(self, x, y): return self._f(T_SCALE, x, y)
    def translate(self, x, y): return self._f(T_TRANSLATE, x, y)
    def rotate(self, a): return self._f(T_ROTATE, a)
    def set_view_box(self, *v): return self._f(SET_VIEW_BOX, *v)
    def global_alpha(self, a): return self._f(SET_GLOBAL_ALPHA, a)

    def mult(self, m, post):
        self.ops.append((T_MULT, tuple(float(F(x)) for x in m) + (int(post),)))
        self._cmd(T_MULT, struct.pack("<6fI", *m, int(post)))
        return self

    def submit(self, list_id):  # clSubmitCommandList, vg.cpp:2959-2966
        self.ops.append((SUBMIT_LIST, (int(list_id),)))
        self._cmd(SUBMIT_LIST, struct.pack("<H", list_id))
        return self

    def bytes(self):
        return bytes(self.buf)


def _clamp(v, lo, hi):
    return min(max(v, lo), hi)


class Model:
    """float32 restatement of what the reference's executor does with the ops of a Writer."""

    def __init__(self, sincos, xform=(1, 0, 0, 1, 0, 0), global_alpha=1.0, tol=0.25, fringe=1.0, canvas=(1280.0, 720.0), scissor=None,
                 cull=False, lists=None, max_depth=16):
        self.lists = lists or {}  # handle index -> (ops, flags): the lists a SubmitCommandList may name
        self.depth, self.max_depth = 0, max_depth
        self.sincos = sincos  # a -> (sin a, cos a) with the build's shared math
        self.rec = Recorder(fringe=fringe)
        self.fringe, self.tol, self.canvas, self.cull = F(fringe), tol, (F(canvas[0]), F(canvas[1])), cull
        sc = scissor if scissor is not None else (0.0, 0.0, canvas[0], canvas[1])
        self.stack = [dict(m=[F(v) for v in xform], alpha=F(global_alpha), sc=[F(v) for v in sc])]
        self.stats = dict(draws=0, culled=0, zero_alpha=0, unsupported=0, paths=0, nested=0)
        self.skip = False
        self.drawn = False
        self.path = None
        self.in_clip = False

    @property
    def s(self):
        return self.stack[-1]

    def avg_scale(self):
        t = self.s["m"]
        sx = np.sqrt(F(t[0] * t[0] + t[2] * t[2]), dtype=F)
        sy = np.sqrt(F(t[1] * t[1] + t[3] * t[3]), dtype=F)
        return F((sx + sy) * F(0.5))

    def _first_draw(self):
        if not self.drawn:
            self.path["xform"] = [float(v) for v in self.s["m"]]
            self.drawn = True

    def run(self, ops, cull=None):
        """one execution of one list (vg.cpp:4273-4637); skipCmds and the culling flag are per execution"""
        if self.depth >= self.max_depth:
            return self
        self.depth += 1
        outer = (self.skip, self.cull)
        self.skip, self.cull = False, (self.cull if cull is None else cull)
        self._run(ops)
        self.skip, self.cull = outer
        self.depth -= 1
        return self

    def _run(self, ops):
        r = self.rec
        for t, v in ops:
            if self.skip and FILL_COLOR <= t <= STROKE_IMAGE:
                self.stats["culled"] += 1
                continue
            m = self.s["m"]
            if t == BEGIN_PATH:
                r.begin_path(float(self.avg_scale()), self.tol, None)
                self.path = r.paths[-1]
                self.drawn = False
                self.stats["paths"] += 1
            elif t == CLOSE_PATH: r.close()
            elif t == MOVE_TO: r.move_to(*v)
            elif t == LINE_TO: r.line_to(*v)
            elif t == CUBIC_TO: r.cubic_to(*v)
            elif t == QUAD_TO: r.quadratic_to(*v)
            elif t == ARC_TO: r.arc_to(*v)
            elif t == ARC: r.arc(*v)
            elif t == RECT: r.rect(*v)
            elif t == ROUNDED_RECT: r.rounded_rect(*v)
            elif t == ROUNDED_RECT_VARYING: r.rounded_rect_varying(*v)
            elif t == CIRCLE: r.circle(*v)
            elif t == ELLIPSE: r.ellipse(*v)
            elif t == POLYLINE: r.polyline(v)
            elif t == FILL_COLOR:
                flags, color = v
                if self.in_clip or (flags & 1):
                    self.stats["unsupported"] += 1
                    continue
                a = int(F(self.s["alpha"] * F(color >> 24)))  # (uint8_t)(globalAlpha * alpha)
                if a == 0:
                    self.stats["zero_alpha"] += 1
                    continue
                self._first_draw()
                r.convex_fill((color & 0xFFFFFF) | (a << 24), bool(flags & 4))
                self.stats["draws"] += 1
            elif t == STROKE_COLOR:
                width, flags, color = v
                if self.in_clip:
                    self.stats["unsupported"] += 1
                    continue
                scaled = F(width) if flags & 32 else F(_clamp(F(F(width) * self.avg_scale()), F(0), F(200)))
                thin = scaled <= self.fringe
                alpha_scale = self.s["alpha"]
                if thin:
                    c = F(_clamp(scaled, F(0), self.fringe))
                    alpha_scale = F(self.s["alpha"] * F(c * c))
                a = int(F(alpha_scale * F(color >> 24)))
                if a == 0:
                    self.stats["zero_alpha"] += 1
                    continue
                aa = bool(flags & 16)
                op = 2 if not aa else (3 if thin else 0)
                self._first_draw()
                r.stroke(op, (color & 0xFFFFFF) | (a << 24), float(self.fringe if thin else scaled), (flags >> 2) & 3, flags & 3)
                self.stats["draws"] += 1
            elif t == BEGIN_CLIP:
                self.in_clip = True
                self.stats["unsupported"] += 1
            elif t == END_CLIP:
                self.in_clip = False
            elif t == SUBMIT_LIST:
                if not self.lists:
                    self.stats["unsupported"] += 1
                elif v[0] in self.lists:
                    self.stats["nested"] += 1
                    child_ops, child_flags = self.lists[v[0]]
                    self.run(child_ops, cull=bool(child_flags & 2))
            elif t in (FILL_GRADIENT, FILL_IMAGE, STROKE_GRADIENT, STROKE_IMAGE, TRI_LIST, LINEAR_GRADIENT, BOX_GRADIENT, RADIAL_GRADIENT,
                       IMAGE_PATTERN, TEXT, TEXT_BOX, RESET_CLIP):
                self.stats["unsupported"] += 1
            elif t == PUSH_STATE:
                self.stack.append(dict(m=list(self.s["m"]), alpha=self.s["alpha"], sc=list(self.s["sc"])))
            elif t == POP_STATE:
                self.stack.pop()
                if self.cull:
                    self.skip = bool(self.s["sc"][2] < 1 or self.s["sc"][3] < 1)
            elif t == RESET_SCISSOR:
                self.s["sc"] = [F(0), F(0), self.canvas[0], self.canvas[1]]
                self.skip = False
            elif t in (SET_SCISSOR, INTERSECT_SCISSOR):
                x, y, w, h = (F(a) for a in v)
                px, py = F(F(m[0] * x + m[2] * y) + m[4]), F(F(m[1] * x + m[3] * y) + m[5])
                sx, sy = F(m[0] * w + m[2] * h), F(m[1] * w + m[3] * h)
                sc = self.s["sc"]
                if t == SET_SCISSOR:
                    minx, miny = _clamp(px, F(0), self.canvas[0]), _clamp(py, F(0), self.canvas[1])
                    maxx, maxy = _clamp(F(px + sx), F(0), self.canvas[0]), _clamp(F(py + sy), F(0), self.canvas[1])
                    self.s["sc"] = [minx, miny, F(maxx - minx), F(maxy - miny)]
                    if self.cull:
                        self.skip = bool(self.s["sc"][2] < 1 or self.s["sc"][3] < 1)
                else:
                    minx, miny = max(px, sc[0]), max(py, sc[1])
                    maxx, maxy = min(F(px + sx), F(sc[0] + sc[2])), min(F(py + sy), F(sc[1] + sc[3]))
                    nw, nh = max(F(0), F(maxx - minx)), max(F(0), F(maxy - miny))
                    self.s["sc"] = [minx, miny, nw, nh]
                    if self.cull:
                        self.skip = not (nw >= 1 and nh >= 1)
            elif t == T_IDENTITY:
                self.s["m"] = [F(1), F(0), F(0), F(1), F(0), F(0)]
            elif t == T_SCALE:
                x, y = F(v[0]), F(v[1])
                self.s["m"] = [F(x * m[0]), F(x * m[1]), F(y * m[2]), F(y * m[3]), m[4], m[5]]
            elif t == T_TRANSLATE:
                x, y = F(v[0]), F(v[1])
                self.s["m"] = [m[0], m[1], m[2], m[3], F(m[4] + F(F(m[0] * x) + F(m[2] * y))), F(m[5] + F(F(m[1] * x) + F(m[3] * y)))]
            elif t == T_ROTATE:
                sn, c = self.sincos(F(v[0]))
                self.s["m"] = [F(F(c * m[0]) + F(sn * m[2])), F(F(c * m[1]) + F(sn * m[3])), F(F(-sn * m[0]) + F(c * m[2])),
                               F(F(-sn * m[1]) + F(c * m[3])), m[4], m[5]]
            elif t == T_MULT:
                b6, post = [F(x) for x in v[:6]], v[6]
                a, b = (m, b6) if post else (b6, m)
                self.s["m"] = [F(F(a[0] * b[0]) + F(a[2] * b[1])), F(F(a[1] * b[0]) + F(a[3] * b[1])), F(F(a[0] * b[2]) + F(a[2] * b[3])),
                               F(F(a[1] * b[2]) + F(a[3] * b[3])), F(F(F(a[0] * b[4]) + F(a[2] * b[5])) + a[4]),
                               F(F(F(a[1] * b[4]) + F(a[3] * b[5])) + a[5])]
            elif t == SET_VIEW_BOX:
                x, y, w, h = (F(a) for a in v)
                sx, sy = F(self.canvas[0] / w), F(self.canvas[1] / h)
                n = [F(sx * m[0]), F(sx * m[1]), F(sy * m[2]), F(sy * m[3]), m[4], m[5]]
                n[4] = F(n[4] - F(F(n[0] * x) + F(n[2] * y)))
                n[5] = F(n[5] - F(F(n[1] * x) + F(n[3] * y)))
                self.s["m"] = n
            elif t == SET_GLOBAL_ALPHA:
                self.s["alpha"] = F(v[0])
        return self


def demo_scene():
    """a command list that exercises every supported command, the lazy path transform, thin / fixed-width
    strokes, alpha, culling and the commands that are skipped"""
    w = Writer()
    w.begin_path().move_to(10, 10).cubic_to(60, -20, 120, 90, 200, 30).line_to(220, 80).stroke(0xFF3366CC, 4.0, stroke_flags(1, 1, True))
    w.push_state().translate(100, 50).rotate(0.7).scale(1.5, 0.75)
    w.begin_path().rounded_rect(0, 0, 120, 60, 9).fill(0x80FF8800, fill_flags(aa=True)).stroke(0xFF000000, 2.0, stroke_flags(0, 0, True))
    w.begin_path().circle(40, 40, 25).translate(5, 5).fill(0xFF00FF00, fill_flags(aa=False))  # transform changes between begin and draw
    w.scale(2, 2).stroke(0xFFFFFFFF, 0.2, stroke_flags(2, 2, True))  # same path, later transform: reuses the first; thin stroke
    w.global_alpha(0.5)
    w.begin_path().move_to(0, 0).quadratic_to(30, 60, 80, 10).arc_to(120, 40, 160, 0, 12).close_path()
    w.stroke(0xFF123456, 3.0, stroke_flags(1, 2, False)).stroke(0xC0654321, 1.0, stroke_flags(0, 1, True, fixed=True))
    w.mult((0.9, 0.1, -0.1, 0.9, 3, 4), True).mult((1, 0, 0, 1, -2, 6), False)
    w.begin_path().ellipse(50, 20, 30, 12).rect(5, 5, 40, 30).rounded_rect_varying(60, 60, 80, 50, 4, 8, 0, 12).arc(10, 90, 15, 0.2, 2.4, 1)
    w.polyline([0, 0, 10, 5, 20, -5, 30, 5]).fill(0xFF0000FF, fill_flags(aa=True)).fill(0xFFABCDEF, fill_flags(concave=True))
    w.raw(FILL_GRADIENT, struct.pack("<IHH", 4, 0, 0)).raw(LINEAR_GRADIENT, struct.pack("<4f2I", 0, 0, 1, 1, 0xFF, 0xFF00))
    w.pop_state()
    w.global_alpha(0.001).begin_path().rect(1, 1, 5, 5).fill(0xFFFFFFFF, fill_flags()).global_alpha(1.0)  # alpha rounds to 0: dropped
    w.push_state().set_scissor(10, 10, 0.5, 100).begin_path().rect(0, 0, 9, 9).fill(0xFF112233, fill_flags())  # culled if allowed
    w.intersect_scissor(0, 0, 50, 50).stroke(0xFF445566, 2, stroke_flags(0, 0, True)).pop_state()
    w.set_view_box(0, 0, 640, 360).begin_path().move_to(1, 1).line_to(50, 70).line_to(90, 10).stroke(0xFFEEDDCC, 2.5, stroke_flags(2, 0, True))
    w.raw(BEGIN_CLIP, struct.pack("<I", 0)).begin_path().rect(0, 0, 10, 10).fill(0xFF000000, fill_flags(aa=False)).raw(END_CLIP)
    w.identity().reset_scissor().begin_path().circle(300, 300, 80).stroke(0xFF00FFFF, 6, stroke_flags(1, 1, True))
    return w
