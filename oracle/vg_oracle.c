/*
 * vg_oracle.c -- plain-C restatement of the reference's hot path, used as the CHECKER for the
 * CUDA kernels.  TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load liboracle.so; the product (libvgb200.so) never
 * links, loads or calls it.
 *
 * What it restates (file:line are in /root/reference):
 *   path accumulator, append/dedupe/close .......... src/path.cpp:44-84, 707-726, 748-784
 *   cubic / quadratic flattening .................... src/path.cpp:86-201
 *   arcTo, rect, rounded rects, circle/ellipse, arc,
 *   polyline ........................................ src/path.cpp:203-705
 *   vec2Dir, calcExtrusionVector .................... src/stroker.cpp:31-53
 *   strokerPolylineStroke (non-AA) .................. src/stroker.cpp:239-273, 1008-1388
 *   strokerPolylineStrokeAA ......................... src/stroker.cpp:275-309, 1390-1991
 *   strokerPolylineStrokeAAThin ..................... src/stroker.cpp:311-332, 1993-2314
 *   strokerConvexFill / ConvexFillAA (scalar) ....... src/stroker.cpp:334-365, 713-806
 *   transformPath / batchTransformPositions (scalar)  src/vg.cpp:4957-4975, src/vg_util.h:24-28
 *   per-sub-path driver loops ....................... src/vg.cpp:3099-3132, 3448-3485
 * It is written as a serial "running cross-section" state machine (one generic ribbon walker
 * parameterised by the number of ids per cross-section), not as a transcription: the reference
 * has three hand-expanded templates.  Pinning: tests/test_oracle.py checks it against the
 * compiled reference itself (oracle/_ref) on the KATs of SURVEY.md App. C, on every
 * cap x join x closed combination, and against committed golden vectors (tests/golden/).
 *
 * Numerics: bx is absent and unpinned, so transcendental functions come from
 * vg_renderer_b200/csrc/vgb_math.h (the same explicit fp32 routines the kernels and the _ref
 * build use).  With respect to bx's own approximations: "parity unpinned" (DESIGN.md section 2);
 * with respect to the reference's sources compiled here (oracle/_ref): pinned.
 * Build with -ffp-contract=off (oracle/Makefile): every operation individually rounded.
 */
#include "vg_oracle.h"
#include "vgb_math.h"

#include <stdlib.h>
#include <string.h>

/* =========================================================================================
 * growable output
 * ======================================================================================= */
typedef struct { float x, y; } v2;

typedef struct {
	v2* pos;
	uint32_t* col;
	uint16_t* idx;
	uint32_t nv, ni, vcap, icap;
} MeshBuf;

static void mb_reserve(MeshBuf* m, uint32_t v, uint32_t i)
{
	if (m->nv + v > m->vcap) {
		m->vcap = (m->nv + v) * 2 + 64;
		m->pos = (v2*)realloc(m->pos, sizeof(v2) * m->vcap);
		m->col = (uint32_t*)realloc(m->col, sizeof(uint32_t) * m->vcap);
	}
	if (m->ni + i > m->icap) {
		m->icap = (m->ni + i) * 2 + 64;
		m->idx = (uint16_t*)realloc(m->idx, sizeof(uint16_t) * m->icap);
	}
}

static uint32_t mb_v(MeshBuf* m, float x, float y, uint32_t c)
{
	mb_reserve(m, 1, 0);
	m->pos[m->nv].x = x;
	m->pos[m->nv].y = y;
	m->col[m->nv] = c;
	return m->nv++;
}

static void mb_tri(MeshBuf* m, uint32_t a, uint32_t b, uint32_t c)
{
	mb_reserve(m, 0, 3);
	m->idx[m->ni++] = (uint16_t)a;
	m->idx[m->ni++] = (uint16_t)b;
	m->idx[m->ni++] = (uint16_t)c;
}

/* =========================================================================================
 * path (src/path.cpp)
 * ======================================================================================= */
typedef struct {
	v2* v;
	uint32_t n, cap;
	vgb_subpath* sp;
	uint32_t nsp, spcap;
	int cur; /* index of current sub-path or -1 */
	float scale, tol;
} PathBuf;

static void pb_reset(PathBuf* p, float scale, float tol)
{
	p->n = 0;
	p->nsp = 0;
	p->cur = -1;
	p->scale = scale;
	p->tol = tol;
}

static void pb_raw(PathBuf* p, float x, float y)
{
	if (p->n == p->cap) {
		p->cap = p->cap * 2 + 64;
		p->v = (v2*)realloc(p->v, sizeof(v2) * p->cap);
	}
	p->v[p->n].x = x;
	p->v[p->n].y = y;
	++p->n;
	++p->sp[p->cur].num_vertices;
}

/* path.cpp:761-784: drop a vertex closer than sqrt(1e-5) to the last KEPT one */
static void pb_add(PathBuf* p, float x, float y)
{
	vgb_subpath* s = &p->sp[p->cur];
	if (s->num_vertices != 0) {
		const v2 l = p->v[s->first_vertex + s->num_vertices - 1];
		const float dx = l.x - x, dy = l.y - y;
		if (dx * dx + dy * dy < VGB_EPSILON) {
			return;
		}
	}
	pb_raw(p, x, y);
}

/* path.cpp:62-78 */
static void pb_move(PathBuf* p, float x, float y)
{
	if (p->cur < 0 || p->sp[p->cur].num_vertices != 0) {
		if (p->nsp == p->spcap) {
			p->spcap = p->spcap * 2 + 16;
			p->sp = (vgb_subpath*)realloc(p->sp, sizeof(vgb_subpath) * p->spcap);
		}
		p->cur = (int)p->nsp++;
		p->sp[p->cur].first_vertex = p->n;
		p->sp[p->cur].num_vertices = 0;
		p->sp[p->cur].is_closed = 0;
	}
	pb_add(p, x, y);
}

static v2 pb_last(const PathBuf* p)
{
	const vgb_subpath* s = &p->sp[p->cur];
	return p->v[s->first_vertex + s->num_vertices - 1];
}

/* path.cpp:707-726 */
static void pb_close(PathBuf* p)
{
	vgb_subpath* s = &p->sp[p->cur];
	if (s->is_closed || s->num_vertices <= 2) {
		return;
	}
	s->is_closed = 1;
	const v2 f = p->v[s->first_vertex];
	const v2 l = p->v[s->first_vertex + s->num_vertices - 1];
	const float dx = l.x - f.x, dy = l.y - f.y;
	if (dx * dx + dy * dy < VGB_EPSILON) {
		--s->num_vertices;
		--p->n;
	}
}

/* path.cpp:86-182.  A node carries (P1..P4); the explicit stack holds right halves still to
 * visit (at most 10).  Flat -> emit P4; not flat and room -> descend left; not flat and no
 * room -> the node is abandoned with no vertex. */
static void pb_cubic(PathBuf* p, float c1x, float c1y, float c2x, float c2y, float ex, float ey)
{
	float pend[10][8];
	int depth = 0;
	const v2 s = pb_last(p);
	float a[8] = { s.x, s.y, c1x, c1y, c2x, c2y, ex, ey };
	const float tt = p->tol / (p->scale * p->scale);
	for (;;) {
		const float dx = a[6] - a[0], dy = a[7] - a[1];
		const float d2 = vgb_abs((a[2] - a[6]) * dy - (a[3] - a[7]) * dx);
		const float d3 = vgb_abs((a[4] - a[6]) * dy - (a[5] - a[7]) * dx);
		const float d23 = d2 + d3;
		if (d23 * d23 <= tt * (dx * dx + dy * dy)) {
			pb_add(p, a[6], a[7]);
			if (depth == 0) {
				return;
			}
			memcpy(a, pend[--depth], sizeof(a));
		} else if (depth < 10) {
			const float x12 = (a[0] + a[2]) * 0.5f, y12 = (a[1] + a[3]) * 0.5f;
			const float x23 = (a[2] + a[4]) * 0.5f, y23 = (a[3] + a[5]) * 0.5f;
			const float x34 = (a[4] + a[6]) * 0.5f, y34 = (a[5] + a[7]) * 0.5f;
			const float x123 = (x12 + x23) * 0.5f, y123 = (y12 + y23) * 0.5f;
			const float x234 = (x23 + x34) * 0.5f, y234 = (y23 + y34) * 0.5f;
			const float xm = (x123 + x234) * 0.5f, ym = (y123 + y234) * 0.5f;
			float* r = pend[depth++];
			r[0] = xm; r[1] = ym; r[2] = x234; r[3] = y234; r[4] = x34; r[5] = y34; r[6] = a[6]; r[7] = a[7];
			a[2] = x12; a[3] = y12; a[4] = x123; a[5] = y123; a[6] = xm; a[7] = ym;
		} else {
			memcpy(a, pend[--depth], sizeof(a));
		}
	}
}

/* path.cpp:184-201 */
static void pb_quad(PathBuf* p, float cx, float cy, float x, float y)
{
	const v2 s = pb_last(p);
	pb_cubic(p, s.x + (2.0f / 3.0f) * (cx - s.x), s.y + (2.0f / 3.0f) * (cy - s.y),
	         x + (2.0f / 3.0f) * (cx - x), y + (2.0f / 3.0f) * (cy - y), x, y);
}

/* points of an arc by rotation recurrence (path.cpp:326-335, 617-626, 670-679): no dedupe */
static void pb_spin(PathBuf* p, float cx, float cy, float rx, float ry, float ca, float sa, float cd, float sd, uint32_t count)
{
	for (uint32_t i = 0; i < count; ++i) {
		const float ns = sd * ca + cd * sa;
		const float nc = cd * ca - sd * sa;
		ca = nc;
		sa = ns;
		pb_raw(p, cx + rx * ca, cy + ry * sa);
	}
}

/* path.cpp:275-286 */
static void pb_rect(PathBuf* p, float x, float y, float w, float h)
{
	if (vgb_abs(w) < VGB_EPSILON || vgb_abs(h) < VGB_EPSILON) {
		return;
	}
	pb_move(p, x, y);
	pb_add(p, x, y + h);
	pb_add(p, x + w, y + h);
	pb_add(p, x + w, y);
	pb_close(p);
}

/* path.cpp:599-631 (pathCircle forwards here, 561-565) */
static void pb_ellipse(PathBuf* p, float cx, float cy, float rx, float ry)
{
	const float avg = (rx + ry) * 0.5f;
	const float da = vgb_acos((p->scale * avg) / ((p->scale * avg) + p->tol)) * 2.0f;
	const uint32_t half = vgb_umax(2u, vgb_f2u(vgb_ceil(VGB_PI / da)));
	const uint32_t np = half * 2;
	pb_move(p, cx + rx, cy);
	const float dt = -VGB_PI2 / (float)np;
	pb_spin(p, cx, cy, rx, ry, 1.0f, 0.0f, vgb_cos(dt), vgb_sin(dt), np - 1);
	pb_close(p);
}

/* path.cpp:288-409 */
static void pb_rrect(PathBuf* p, float x, float y, float w, float h, float r)
{
	if (r < 0.1f) {
		pb_rect(p, x, y, w, h);
		return;
	}
	const float mr = (w < h ? w : h) * 0.5f;
	if (w == h && r >= mr - VGB_EPSILON) {
		pb_ellipse(p, x + mr, y + mr, mr, mr);
		return;
	}
	r = r < mr ? r : mr;
	const float da = vgb_acos((p->scale * r) / ((p->scale * r) + p->tol)) * 2.0f;
	const uint32_t half = vgb_umax(2u, vgb_f2u(vgb_ceil(VGB_PI / da)));
	const uint32_t nq = (half >> 1) + 1;
	const float dt = -VGB_PI_HALF / (float)(nq - 1);
	const float cd = vgb_cos(dt), sd = vgb_sin(dt);
	pb_move(p, x, y + r);
	pb_add(p, x, y + h - r);
	pb_spin(p, x + r, y + h - r, r, r, -1.0f, 0.0f, cd, sd, nq - 1);
	pb_add(p, x + w - r, y + h);
	pb_spin(p, x + w - r, y + h - r, r, r, 0.0f, 1.0f, cd, sd, nq - 1);
	pb_add(p, x + w, y + r);
	pb_spin(p, x + w - r, y + r, r, r, 1.0f, 0.0f, cd, sd, nq - 1);
	pb_add(p, x + r, y);
	pb_spin(p, x + r, y + r, r, r, 0.0f, -1.0f, cd, sd, nq - 1);
	pb_close(p);
}

/* one rounded corner of path.cpp:426-556 */
static void pb_corner(PathBuf* p, float cx, float cy, float r, float ca, float sa)
{
	const float hda = vgb_acos((p->scale * r) / ((p->scale * r) + p->tol));
	const uint32_t half = vgb_umax(2u, vgb_f2u(vgb_ceil(VGB_PI_HALF / hda)));
	const uint32_t nq = (half >> 1) + 1;
	const float dt = -VGB_PI_HALF / (float)(nq - 1);
	pb_spin(p, cx, cy, r, r, ca, sa, vgb_cos(dt), vgb_sin(dt), nq - 1);
}

static float min3(float a, float b, float c)
{
	const float m = a < b ? a : b;
	return m < c ? m : c;
}

/* path.cpp:411-559 */
static void pb_rrect_var(PathBuf* p, float x, float y, float w, float h, float tl, float tr, float br, float bl)
{
	if (tl < 0.1f && bl < 0.1f && br < 0.1f && tr < 0.1f) {
		pb_rect(p, x, y, w, h);
		return;
	}
	const float hw = w * 0.5f, hh = h * 0.5f;
	const float rtl = min3(tl, hw, hh), rtr = min3(tr, hw, hh), rbl = min3(bl, hw, hh), rbr = min3(br, hw, hh);
	if (rtl < 0.1f) {
		pb_move(p, x, y);
	} else {
		pb_move(p, x + rtl, y);
		pb_corner(p, x + rtl, y + rtl, rtl, 0.0f, -1.0f);
	}
	if (rbl < 0.1f) {
		pb_add(p, x, y + h);
	} else {
		pb_add(p, x, y + h - rbl);
		pb_corner(p, x + rbl, y + h - rbl, rbl, -1.0f, 0.0f);
	}
	if (rbr < 0.1f) {
		pb_add(p, x + w, y + h);
	} else {
		pb_add(p, x + w - rbr, y + h);
		pb_corner(p, x + w - rbr, y + h - rbr, rbr, 0.0f, 1.0f);
	}
	if (rtr < 0.1f) {
		pb_add(p, x + w, y);
	} else {
		pb_add(p, x + w, y + rtr);
		pb_corner(p, x + w - rtr, y + rtr, rtr, 1.0f, 0.0f);
	}
	pb_close(p);
}

/* path.cpp:633-682 */
static void pb_arc(PathBuf* p, float cx, float cy, float r, float a0, float a1, int cw)
{
	while (a0 > VGB_PI2) a0 -= VGB_PI2;
	while (a1 > VGB_PI2) a1 -= VGB_PI2;
	if (!cw) {
		while (a0 < a1) a0 += VGB_PI2;
	} else {
		while (a1 < a0) a1 += VGB_PI2;
	}
	const float da = vgb_acos((p->scale * r) / ((p->scale * r) + p->tol)) * 2.0f;
	const uint32_t np = vgb_umax(2u, vgb_f2u(vgb_ceil(vgb_abs(a1 - a0) / da)));
	const float dt = (a1 - a0) / (float)np;
	const float ca = vgb_cos(a0), sa = vgb_sin(a0);
	if (p->cur >= 0 && p->sp[p->cur].num_vertices != 0) {
		pb_add(p, cx + r * ca, cy + r * sa);
	} else {
		pb_move(p, cx + r * ca, cy + r * sa);
	}
	pb_spin(p, cx, cy, r, r, ca, sa, vgb_cos(dt), vgb_sin(dt), np);
}

/* path.cpp:203-273 */
static void pb_arc_to(PathBuf* p, float x1, float y1, float x2, float y2, float r)
{
	const v2 s = pb_last(p);
	float ax = s.x - x1, ay = s.y - y1, bx = x2 - x1, by = y2 - y1;
	float l = ax * ax + ay * ay;
	float inv = l < VGB_EPSILON ? 0.0f : vgb_rsqrt(l);
	ax *= inv;
	ay *= inv;
	l = bx * bx + by * by;
	inv = l < VGB_EPSILON ? 0.0f : vgb_rsqrt(l);
	bx *= inv;
	by *= inv;
	const float ang = vgb_acos(ax * bx + ay * by);
	const float d = r / vgb_tan(ang / 2.0f);
	if (d > 10000.0f) {
		pb_add(p, x1, y1);
		return;
	}
	const float cross = bx * ay - ax * by;
	if (cross > 0.0f) {
		pb_arc(p, x1 + ax * d + ay * r, y1 + ay * d - ax * r, r, vgb_atan2(ax, -ay), vgb_atan2(-bx, by), 1);
	} else {
		pb_arc(p, x1 + ax * d - ay * r, y1 + ay * d + ax * r, r, vgb_atan2(-ax, ay), vgb_atan2(bx, -by), 0);
	}
}

/* path.cpp:684-705 */
static void pb_polyline(PathBuf* p, const float* c, uint32_t np)
{
	if (p->sp[p->cur].num_vertices > 0 && np > 0) {
		const v2 l = pb_last(p);
		const float dx = l.x - c[0], dy = l.y - c[1];
		if (dx * dx + dy * dy < VGB_EPSILON) {
			c += 2;
			--np;
		}
	}
	for (uint32_t i = 0; i < np; ++i) {
		pb_raw(p, c[2 * i], c[2 * i + 1]);
	}
}

static int pb_replay(PathBuf* p, const uint8_t* cmd, uint32_t ncmd, const float* d)
{
	for (uint32_t i = 0; i < ncmd; ++i) {
		switch (cmd[i]) {
		case VGB_CMD_MOVE_TO: pb_move(p, d[0], d[1]); d += 2; break;
		case VGB_CMD_LINE_TO: pb_add(p, d[0], d[1]); d += 2; break;
		case VGB_CMD_CUBIC_TO: pb_cubic(p, d[0], d[1], d[2], d[3], d[4], d[5]); d += 6; break;
		case VGB_CMD_QUAD_TO: pb_quad(p, d[0], d[1], d[2], d[3]); d += 4; break;
		case VGB_CMD_ARC_TO: pb_arc_to(p, d[0], d[1], d[2], d[3], d[4]); d += 5; break;
		case VGB_CMD_RECT: pb_rect(p, d[0], d[1], d[2], d[3]); d += 4; break;
		case VGB_CMD_ROUNDED_RECT: pb_rrect(p, d[0], d[1], d[2], d[3], d[4]); d += 5; break;
		case VGB_CMD_ROUNDED_RECT_VARYING: pb_rrect_var(p, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]); d += 8; break;
		case VGB_CMD_CIRCLE: pb_ellipse(p, d[0], d[1], d[2], d[2]); d += 3; break;
		case VGB_CMD_ELLIPSE: pb_ellipse(p, d[0], d[1], d[2], d[3]); d += 4; break;
		case VGB_CMD_ARC: pb_arc(p, d[0], d[1], d[2], d[3], d[4], d[5] != 0.0f); d += 6; break;
		case VGB_CMD_POLYLINE: {
			uint32_t np;
			memcpy(&np, d, 4);
			pb_polyline(p, d + 1, np);
			d += 1 + 2 * (size_t)np;
		} break;
		case VGB_CMD_CLOSE: pb_close(p); break;
		default: return -1;
		}
	}
	return 0;
}

/* =========================================================================================
 * stroker (src/stroker.cpp)
 * ======================================================================================= */
/* stroker.cpp:31-38 */
static v2 dir2(v2 a, v2 b)
{
	const float dx = b.x - a.x, dy = b.y - a.y;
	const float l = dx * dx + dy * dy;
	const float inv = l < VGB_EPSILON ? 0.0f : vgb_rsqrt(l);
	v2 r = { dx * inv, dy * inv };
	return r;
}

/* stroker.cpp:40-53 */
static v2 extrude(v2 d01, v2 d12)
{
	v2 v = { -d01.y, d01.x };
	const float cross = d12.x * d01.y - d01.x * d12.y;
	if (vgb_abs(cross) > (1.0f / 100.0f)) {
		const float s = 1.0f / cross;
		v.x = (d01.x - d12.x) * s;
		v.y = (d01.y - d12.y) * s;
	}
	return v;
}

/*
 * A stroke is a ribbon of cross-sections; a cross-section is W vertex ids ordered from the
 * left edge to the right edge (W = 2 plain stroke, 3 thin AA stroke, 4 AA stroke).  Two
 * consecutive cross-sections are bridged by W-1 quads, each (pa,pb,cb | pa,cb,ca): this one
 * rule generates stroker.cpp:1118-1121 (W=2), 2092-2097 (W=3) and 1557-1564 (W=4).
 */
typedef struct { uint32_t id[4]; } Cross;

static void bridge(MeshBuf* m, int W, const Cross* p, const Cross* c)
{
	for (int k = 0; k + 1 < W; ++k) {
		mb_tri(m, p->id[k], p->id[k + 1], c->id[k + 1]);
		mb_tri(m, p->id[k], c->id[k + 1], c->id[k]);
	}
}

static Cross cross_seq(uint32_t first, int W, int reversed)
{
	Cross c;
	for (int k = 0; k < W; ++k) {
		c.id[k] = reversed ? first + (uint32_t)(W - 1 - k) : first + (uint32_t)k;
	}
	return c;
}

typedef struct {
	float scale, tol, fringe;
} StrokerCfg;

/* ---- AA stroke, stroker.cpp:1390-1991 ---------------------------------------------------- */
static void stroke_aa(MeshBuf* m, const StrokerCfg* cfg, const v2* vtx, uint32_t n, int closed, uint32_t color, float width, uint32_t cap, uint32_t join)
{
	const uint32_t c0 = color & 0x00FFFFFFu;
	const float fr = cfg->fringe;
	const float hsw = (width - fr) * 0.5f;
	const float haa = hsw + fr;
	const float da = vgb_acos((cfg->scale * hsw) / ((cfg->scale * hsw) + cfg->tol)) * 2.0f;
	const uint32_t nhc = vgb_umax(2u, vgb_f2u(vgb_ceil(VGB_PI / da)));
	Cross prev, first;
	int havePrev = 0, haveFirst = 0;
	v2 d01;
	if (closed) {
		cap = VGB_CAP_BUTT; /* dispatch, stroker.cpp:285-303 */
		d01 = dir2(vtx[n - 1], vtx[0]);
	} else {
		const v2 p0 = vtx[0];
		d01 = dir2(p0, vtx[1]);
		const v2 l = { -d01.y, d01.x };
		if (cap == VGB_CAP_ROUND) {
			/* 1475-1515 */
			const float sa0 = vgb_atan2(l.y, l.x);
			for (uint32_t i = 0; i < nhc; ++i) {
				const float a = sa0 + (float)i * VGB_PI / (float)(nhc - 1);
				const float ca = vgb_cos(a), sa = vgb_sin(a);
				mb_v(m, p0.x + ca * hsw, p0.y + sa * hsw, color);
				mb_v(m, p0.x + ca * haa, p0.y + sa * haa, c0);
			}
			for (uint32_t i = 0; i + 2 < nhc; ++i) mb_tri(m, 0, 2 * i + 2, 2 * i + 4);
			for (uint32_t i = 0; i + 1 < nhc; ++i) {
				const uint32_t b = 2 * i;
				mb_tri(m, b, b + 1, b + 3);
				mb_tri(m, b, b + 3, b + 2);
			}
			prev.id[0] = 1; prev.id[1] = 0; prev.id[2] = 2 * (nhc - 1); prev.id[3] = 2 * (nhc - 1) + 1;
		} else {
			/* butt 1422-1447 pulls the fringe back by one fringe width, square 1448-1474 by hsw(_aa) */
			const float ka = cap == VGB_CAP_BUTT ? fr : haa;
			const float ki = cap == VGB_CAP_BUTT ? 0.0f : hsw;
			const v2 li = { l.x * hsw, l.y * hsw }, la = { l.x * haa, l.y * haa };
			const v2 da_ = { d01.x * ka, d01.y * ka }, di = { d01.x * ki, d01.y * ki };
			mb_v(m, p0.x + (la.x - da_.x), p0.y + (la.y - da_.y), c0);
			if (cap == VGB_CAP_BUTT) {
				mb_v(m, p0.x + li.x, p0.y + li.y, color);
				mb_v(m, p0.x - li.x, p0.y - li.y, color);
			} else {
				mb_v(m, p0.x + (li.x - di.x), p0.y + (li.y - di.y), color);
				mb_v(m, p0.x - (li.x + di.x), p0.y - (li.y + di.y), color);
			}
			mb_v(m, p0.x - (la.x + da_.x), p0.y - (la.y + da_.y), c0);
			mb_tri(m, 0, 2, 1);
			mb_tri(m, 0, 3, 2);
			prev = cross_seq(0, 4, 0);
		}
		havePrev = 1;
	}
	const uint32_t iBeg = closed ? 0 : 1, iEnd = closed ? n : n - 1;
	for (uint32_t i = iBeg; i < iEnd; ++i) {
		const v2 p1 = vtx[i];
		const v2 d12 = dir2(p1, vtx[i == n - 1 ? 0 : i + 1]);
		const v2 v = extrude(d01, d12);
		const v2 va = { v.x * haa, v.y * haa }, vi = { v.x * hsw, v.y * hsw };
		const int left = (d12.x * va.x + d12.y * va.y) >= 0.0f; /* 1534-1535 */
		const float sg = left ? 1.0f : -1.0f;
		const uint32_t F = m->nv;
		Cross in = cross_seq(F, 4, !left), out;
		/* inner corner pair, identical for every join type: 1538-1539 / 1695-1696 */
		mb_v(m, p1.x + sg * va.x, p1.y + sg * va.y, c0);
		mb_v(m, p1.x + sg * vi.x, p1.y + sg * vi.y, color);
		if (join == VGB_JOIN_MITER) {
			mb_v(m, p1.x - sg * vi.x, p1.y - sg * vi.y, color);
			mb_v(m, p1.x - sg * va.x, p1.y - sg * va.y, c0);
			out = in;
			if (havePrev) bridge(m, 4, &prev, &in);
		} else {
			/* outer side: left inner -> perpCW, right inner -> perpCCW (1581-1582, 1737-1738) */
			const v2 s01 = { sg * d01.y, -sg * d01.x }, s12 = { sg * d12.y, -sg * d12.x };
			uint32_t na = 1;
			float a01 = 0.0f, step = 0.0f;
			if (join == VGB_JOIN_ROUND) {
				a01 = vgb_atan2(s01.y, s01.x);
				float a12 = vgb_atan2(s12.y, s12.x);
				if (left) {
					if (a12 < a01) a12 += VGB_PI2;
					na = vgb_umax(2u, vgb_f2u((a12 - a01) / da));
				} else {
					if (a12 > a01) a12 -= VGB_PI2;
					na = vgb_umax(2u, vgb_f2u((a01 - a12) / da));
				}
				step = (a12 - a01) / (float)na;
			}
			const float bev = join == VGB_JOIN_BEVEL ? vgb_abs(s01.x * s12.x + s01.y * s12.y) * fr : 0.0f;
			{
				float x = p1.x + s01.x * hsw, y = p1.y + s01.y * hsw;
				if (join == VGB_JOIN_BEVEL) {
					x = x - d01.x * bev;
					y = y - d01.y * bev;
				}
				mb_v(m, x, y, color);
				mb_v(m, p1.x + s01.x * haa, p1.y + s01.y * haa, c0);
			}
			for (uint32_t k = 1; k < na; ++k) {
				const float a = a01 + (float)k * step;
				const float ca = vgb_cos(a), sa = vgb_sin(a);
				mb_v(m, p1.x + ca * hsw, p1.y + sa * hsw, color);
				mb_v(m, p1.x + ca * haa, p1.y + sa * haa, c0);
			}
			{
				float x = p1.x + s12.x * hsw, y = p1.y + s12.y * hsw;
				if (join == VGB_JOIN_BEVEL) {
					x = x + d12.x * bev;
					y = y + d12.y * bev;
				}
				mb_v(m, x, y, color);
				mb_v(m, p1.x + s12.x * haa, p1.y + s12.y * haa, c0);
			}
			if (havePrev) bridge(m, 4, &prev, &in);
			for (uint32_t k = 0; k < na; ++k) {
				const uint32_t A = F + 2 + 2 * k;
				if (left) {
					mb_tri(m, F + 1, A, A + 2);
					mb_tri(m, A, A + 1, A + 3);
					mb_tri(m, A, A + 3, A + 2);
				} else {
					mb_tri(m, F + 1, A + 2, A);
					mb_tri(m, A, A + 3, A + 1);
					mb_tri(m, A, A + 2, A + 3);
				}
			}
			const uint32_t E = F + 2 + 2 * na;
			if (left) {
				out.id[0] = F; out.id[1] = F + 1; out.id[2] = E; out.id[3] = E + 1;
			} else {
				out.id[0] = E + 1; out.id[1] = E; out.id[2] = F + 1; out.id[3] = F;
			}
		}
		if (!havePrev && !haveFirst) {
			first = in;
			haveFirst = 1;
		}
		prev = out;
		havePrev = 1;
		d01 = d12;
	}
	if (closed) {
		bridge(m, 4, &prev, &first); /* 1970-1984 */
		return;
	}
	{
		const v2 p1 = vtx[n - 1];
		const v2 l = { -d01.y, d01.x };
		const uint32_t C = m->nv;
		if (cap == VGB_CAP_ROUND) {
			/* 1917-1968 */
			const float sa0 = vgb_atan2(l.y, l.x);
			for (uint32_t i = 0; i < nhc; ++i) {
				const float a = sa0 - (float)i * VGB_PI / (float)(nhc - 1);
				const float ca = vgb_cos(a), sa = vgb_sin(a);
				mb_v(m, p1.x + ca * hsw, p1.y + sa * hsw, color);
				mb_v(m, p1.x + ca * haa, p1.y + sa * haa, c0);
			}
			Cross c;
			c.id[0] = C + 1; c.id[1] = C; c.id[2] = C + 2 * (nhc - 1); c.id[3] = C + 2 * (nhc - 1) + 1;
			bridge(m, 4, &prev, &c);
			for (uint32_t i = 0; i + 2 < nhc; ++i) mb_tri(m, C, C + 2 * i + 4, C + 2 * i + 2);
			for (uint32_t i = 0; i + 1 < nhc; ++i) {
				const uint32_t b = C + 2 * i;
				mb_tri(m, b, b + 3, b + 1);
				mb_tri(m, b, b + 2, b + 3);
			}
		} else {
			/* 1858-1916 */
			const float ka = cap == VGB_CAP_BUTT ? fr : haa;
			const float ki = cap == VGB_CAP_BUTT ? 0.0f : hsw;
			const v2 li = { l.x * hsw, l.y * hsw }, la = { l.x * haa, l.y * haa };
			const v2 da_ = { d01.x * ka, d01.y * ka }, di = { d01.x * ki, d01.y * ki };
			mb_v(m, p1.x + (la.x + da_.x), p1.y + (la.y + da_.y), c0);
			if (cap == VGB_CAP_BUTT) {
				mb_v(m, p1.x + li.x, p1.y + li.y, color);
				mb_v(m, p1.x - li.x, p1.y - li.y, color);
			} else {
				mb_v(m, p1.x + (li.x + di.x), p1.y + (li.y + di.y), color);
				mb_v(m, p1.x - (li.x - di.x), p1.y - (li.y - di.y), color);
			}
			mb_v(m, p1.x - (la.x - da_.x), p1.y - (la.y - da_.y), c0);
			const Cross c = cross_seq(C, 4, 0);
			bridge(m, 4, &prev, &c);
			mb_tri(m, C, C + 1, C + 2);
			mb_tri(m, C, C + 2, C + 3);
		}
	}
}

/* ---- plain stroke, stroker.cpp:1008-1388 ------------------------------------------------ */
static void stroke_plain(MeshBuf* m, const StrokerCfg* cfg, const v2* vtx, uint32_t n, int closed, float width, uint32_t cap, uint32_t join)
{
	const float hsw = width * 0.5f;
	const float da = vgb_acos((cfg->scale * hsw) / ((cfg->scale * hsw) + cfg->tol)) * 2.0f;
	const uint32_t nhc = vgb_umax(2u, vgb_f2u(vgb_ceil(VGB_PI / da)));
	Cross prev, first;
	int havePrev = 0, haveFirst = 0;
	v2 d01;
	if (closed) {
		cap = VGB_CAP_BUTT;
		d01 = dir2(vtx[n - 1], vtx[0]);
	} else {
		const v2 p0 = vtx[0];
		d01 = dir2(p0, vtx[1]);
		const v2 l = { -d01.y, d01.x };
		if (cap == VGB_CAP_ROUND) {
			/* 1059-1080 */
			const float sa0 = vgb_atan2(l.y, l.x);
			for (uint32_t i = 0; i < nhc; ++i) {
				const float a = sa0 + (float)i * VGB_PI / (float)(nhc - 1);
				mb_v(m, p0.x + vgb_cos(a) * hsw, p0.y + vgb_sin(a) * hsw, 0);
			}
			for (uint32_t i = 0; i + 2 < nhc; ++i) mb_tri(m, 0, i + 1, i + 2);
			prev.id[0] = 0;
			prev.id[1] = nhc - 1;
		} else {
			const v2 li = { l.x * hsw, l.y * hsw };
			if (cap == VGB_CAP_BUTT) { /* 1032-1044 */
				mb_v(m, p0.x + li.x, p0.y + li.y, 0);
				mb_v(m, p0.x - li.x, p0.y - li.y, 0);
			} else { /* 1045-1058 */
				const v2 di = { d01.x * hsw, d01.y * hsw };
				mb_v(m, p0.x + (li.x - di.x), p0.y + (li.y - di.y), 0);
				mb_v(m, p0.x - (li.x + di.x), p0.y - (li.y + di.y), 0);
			}
			prev = cross_seq(0, 2, 0);
		}
		havePrev = 1;
	}
	const uint32_t iBeg = closed ? 0 : 1, iEnd = closed ? n : n - 1;
	for (uint32_t i = iBeg; i < iEnd; ++i) {
		const v2 p1 = vtx[i];
		const v2 d12 = dir2(p1, vtx[i == n - 1 ? 0 : i + 1]);
		const v2 v = extrude(d01, d12);
		const v2 vi = { v.x * hsw, v.y * hsw };
		const int left = (d12.x * vi.x + d12.y * vi.y) >= 0.0f; /* 1099-1100 */
		const float sg = left ? 1.0f : -1.0f;
		const uint32_t F = m->nv;
		Cross in = cross_seq(F, 2, !left), out;
		mb_v(m, p1.x + sg * vi.x, p1.y + sg * vi.y, 0);
		if (join == VGB_JOIN_MITER) {
			mb_v(m, p1.x - sg * vi.x, p1.y - sg * vi.y, 0);
			out = in;
			if (havePrev) bridge(m, 2, &prev, &in);
		} else {
			const v2 s01 = { sg * d01.y, -sg * d01.x }, s12 = { sg * d12.y, -sg * d12.x };
			uint32_t na = 1;
			float a01 = 0.0f, step = 0.0f;
			if (join == VGB_JOIN_ROUND) {
				a01 = vgb_atan2(s01.y, s01.x);
				float a12 = vgb_atan2(s12.y, s12.x);
				if (left) {
					if (a12 < a01) a12 += VGB_PI2;
					na = vgb_umax(2u, vgb_f2u((a12 - a01) / da));
				} else {
					if (a12 > a01) a12 -= VGB_PI2;
					na = vgb_umax(2u, vgb_f2u((a01 - a12) / da));
				}
				step = (a12 - a01) / (float)na;
			}
			mb_v(m, p1.x + s01.x * hsw, p1.y + s01.y * hsw, 0);
			for (uint32_t k = 1; k < na; ++k) {
				const float a = a01 + (float)k * step;
				mb_v(m, p1.x + hsw * vgb_cos(a), p1.y + hsw * vgb_sin(a), 0);
			}
			mb_v(m, p1.x + s12.x * hsw, p1.y + s12.y * hsw, 0);
			if (havePrev) bridge(m, 2, &prev, &in);
			for (uint32_t k = 0; k < na; ++k) {
				if (left) mb_tri(m, F, F + k + 1, F + k + 2);
				else mb_tri(m, F, F + k + 2, F + k + 1);
			}
			if (left) {
				out.id[0] = F; out.id[1] = F + na + 1;
			} else {
				out.id[0] = F + na + 1; out.id[1] = F;
			}
		}
		if (!havePrev && !haveFirst) {
			first = in;
			haveFirst = 1;
		}
		prev = out;
		havePrev = 1;
		d01 = d12;
	}
	if (closed) {
		bridge(m, 2, &prev, &first); /* 1372-1381 */
		return;
	}
	{
		const v2 p1 = vtx[n - 1];
		const v2 l = { -d01.y, d01.x };
		const uint32_t C = m->nv;
		if (cap == VGB_CAP_ROUND) {
			/* 1342-1371 */
			const float sa0 = vgb_atan2(l.y, l.x);
			for (uint32_t i = 0; i < nhc; ++i) {
				const float a = sa0 - (float)i * VGB_PI / (float)(nhc - 1);
				mb_v(m, p1.x + vgb_cos(a) * hsw, p1.y + vgb_sin(a) * hsw, 0);
			}
			Cross c;
			c.id[0] = C; c.id[1] = C + nhc - 1;
			bridge(m, 2, &prev, &c);
			for (uint32_t i = 0; i + 2 < nhc; ++i) mb_tri(m, C, C + i + 2, C + i + 1);
		} else {
			const v2 li = { l.x * hsw, l.y * hsw };
			if (cap == VGB_CAP_BUTT) { /* 1303-1321 */
				mb_v(m, p1.x + li.x, p1.y + li.y, 0);
				mb_v(m, p1.x - li.x, p1.y - li.y, 0);
			} else { /* 1322-1341 */
				const v2 di = { d01.x * hsw, d01.y * hsw };
				mb_v(m, p1.x + (li.x + di.x), p1.y + (li.y + di.y), 0);
				mb_v(m, p1.x - (li.x - di.x), p1.y - (li.y - di.y), 0);
			}
			const Cross c = cross_seq(C, 2, 0);
			bridge(m, 2, &prev, &c);
		}
	}
}

/* ---- thin AA stroke, stroker.cpp:1993-2314 (round cap -> square, round join -> bevel: 318-327) */
static void stroke_thin(MeshBuf* m, const StrokerCfg* cfg, const v2* vtx, uint32_t n, int closed, uint32_t color, uint32_t cap, uint32_t join)
{
	const uint32_t c0 = color & 0x00FFFFFFu;
	const float haa = cfg->fringe;
	if (cap == VGB_CAP_ROUND) cap = VGB_CAP_SQUARE;
	if (join == VGB_JOIN_ROUND) join = VGB_JOIN_BEVEL;
	Cross prev, first;
	int havePrev = 0, haveFirst = 0;
	v2 d01;
	if (closed) {
		d01 = dir2(vtx[n - 1], vtx[0]);
	} else {
		const v2 p0 = vtx[0];
		d01 = dir2(p0, vtx[1]);
		const v2 la = { -d01.y * haa, d01.x * haa };
		if (cap == VGB_CAP_BUTT) { /* 2021-2035 */
			mb_v(m, p0.x + la.x, p0.y + la.y, c0);
			mb_v(m, p0.x, p0.y, color);
			mb_v(m, p0.x - la.x, p0.y - la.y, c0);
		} else { /* 2036-2051 */
			const v2 da_ = { d01.x * haa, d01.y * haa };
			mb_v(m, p0.x + (la.x - da_.x), p0.y + (la.y - da_.y), c0);
			mb_v(m, p0.x, p0.y, color);
			mb_v(m, p0.x - (la.x + da_.x), p0.y - (la.y + da_.y), c0);
		}
		prev = cross_seq(0, 3, 0);
		havePrev = 1;
	}
	const uint32_t iBeg = closed ? 0 : 1, iEnd = closed ? n : n - 1;
	for (uint32_t i = iBeg; i < iEnd; ++i) {
		const v2 p1 = vtx[i];
		const v2 d12 = dir2(p1, vtx[i == n - 1 ? 0 : i + 1]);
		const v2 v = extrude(d01, d12);
		const v2 va = { v.x * haa, v.y * haa };
		const int left = (d12.x * va.x + d12.y * va.y) >= 0.0f; /* 2072-2073 */
		const float sg = left ? 1.0f : -1.0f;
		const uint32_t F = m->nv;
		Cross in = cross_seq(F, 3, !left), out;
		mb_v(m, p1.x + sg * va.x, p1.y + sg * va.y, c0);
		mb_v(m, p1.x, p1.y, color);
		if (join == VGB_JOIN_MITER) {
			mb_v(m, p1.x - sg * va.x, p1.y - sg * va.y, c0);
			out = in;
			if (havePrev) bridge(m, 3, &prev, &in);
		} else {
			const v2 s01 = { sg * d01.y, -sg * d01.x }, s12 = { sg * d12.y, -sg * d12.x };
			mb_v(m, p1.x + s01.x * haa, p1.y + s01.y * haa, c0);
			mb_v(m, p1.x + s12.x * haa, p1.y + s12.y * haa, c0);
			if (havePrev) bridge(m, 3, &prev, &in);
			if (left) {
				mb_tri(m, F + 1, F + 2, F + 3);
				out.id[0] = F; out.id[1] = F + 1; out.id[2] = F + 3;
			} else {
				mb_tri(m, F + 1, F + 3, F + 2);
				out.id[0] = F + 3; out.id[1] = F + 1; out.id[2] = F;
			}
		}
		if (!havePrev && !haveFirst) {
			first = in;
			haveFirst = 1;
		}
		prev = out;
		havePrev = 1;
		d01 = d12;
	}
	if (closed) {
		bridge(m, 3, &prev, &first); /* 2295-2307 */
		return;
	}
	{
		const v2 p1 = vtx[n - 1];
		const v2 la = { -d01.y * haa, d01.x * haa };
		const uint32_t C = m->nv;
		if (cap == VGB_CAP_BUTT) { /* 2247-2268 */
			mb_v(m, p1.x + la.x, p1.y + la.y, c0);
			mb_v(m, p1.x, p1.y, color);
			mb_v(m, p1.x - la.x, p1.y - la.y, c0);
		} else { /* 2269-2291 */
			const v2 da_ = { d01.x * haa, d01.y * haa };
			mb_v(m, p1.x + (la.x + da_.x), p1.y + (la.y + da_.y), c0);
			mb_v(m, p1.x, p1.y, color);
			mb_v(m, p1.x - (la.x - da_.x), p1.y - (la.y - da_.y), c0);
		}
		const Cross c = cross_seq(C, 3, 0);
		bridge(m, 3, &prev, &c);
	}
}

/* ---- convex fills, stroker.cpp:334-365 and 713-806 (scalar variant) -------------------- */
static void fill_fan(MeshBuf* m, const v2* vtx, uint32_t n)
{
	for (uint32_t i = 0; i < n; ++i) mb_v(m, vtx[i].x, vtx[i].y, 0);
	for (uint32_t k = 1; k + 1 < n; ++k) mb_tri(m, 0, k, k + 1);
}

static void fill_aa(MeshBuf* m, const StrokerCfg* cfg, const v2* vtx, uint32_t n, uint32_t color)
{
	const uint32_t c0 = color & 0x00FFFFFFu;
	const v2 a = { vtx[1].x - vtx[0].x, vtx[1].y - vtx[0].y }, b = { vtx[2].x - vtx[0].x, vtx[2].y - vtx[0].y };
	const float aa = cfg->fringe * 0.5f * vgb_sign(a.x * b.y - b.x * a.y);
	v2 d01 = dir2(vtx[n - 1], vtx[0]);
	for (uint32_t i = 0; i < n; ++i) {
		const v2 d12 = dir2(vtx[i], vtx[i == n - 1 ? 0 : i + 1]);
		const v2 v = extrude(d01, d12);
		const v2 e = { v.x * aa, v.y * aa };
		mb_v(m, vtx[i].x + e.x, vtx[i].y + e.y, color);
		mb_v(m, vtx[i].x - e.x, vtx[i].y - e.y, c0);
		d01 = d12;
	}
	for (uint32_t k = 0; k + 2 < n; ++k) mb_tri(m, 0, 2 * k + 2, 2 * k + 4);
	for (uint32_t i = 0; i < n; ++i) {
		const uint32_t f = 2 * i;
		const uint32_t g = (i + 1 == n) ? 0 : f + 2; /* next polygon vertex's inner id */
		mb_tri(m, f, f + 1, g + 1);
		mb_tri(m, f, g + 1, g);
	}
}

static uint32_t min_verts(uint32_t op)
{
	return (op == VGB_OP_CONVEX_FILL_AA || op == VGB_OP_CONVEX_FILL) ? 3u : 2u; /* vg.cpp:3101, 3450 */
}

/* one strokerXXX call; returns 1 when the mesh has per-vertex colours */
static int run_mesh(MeshBuf* m, const StrokerCfg* cfg, uint32_t op, const v2* vtx, uint32_t n, int closed, uint32_t color, float width, uint32_t cap, uint32_t join)
{
	m->nv = 0;
	m->ni = 0;
	switch (op) {
	case VGB_OP_STROKE_AA: stroke_aa(m, cfg, vtx, n, closed, color, width, cap, join); return 1;
	case VGB_OP_STROKE: stroke_plain(m, cfg, vtx, n, closed, width, cap, join); return 0;
	case VGB_OP_STROKE_AA_THIN: stroke_thin(m, cfg, vtx, n, closed, color, cap, join); return 1;
	case VGB_OP_CONVEX_FILL_AA: fill_aa(m, cfg, vtx, n, color); return 1;
	case VGB_OP_CONVEX_FILL: fill_fan(m, vtx, n); return 0;
	default: return 0;
	}
}

/* =========================================================================================
 * C entry points (same contracts as oracle/ref_capi.cpp)
 * ======================================================================================= */
static __thread PathBuf t_path;
static __thread MeshBuf t_mesh;
static __thread v2* t_xf;
static __thread uint32_t t_xfcap;

uint32_t vgo_variant(void) { return 4; /* bit2: restatement */ }

int vgo_flatten(const uint8_t* cmd, uint32_t ncmd, const float* data, float scale, float tol,
                float* pts, uint32_t ptsCap, vgb_subpath* sp, uint32_t spCap, uint32_t* numPts, uint32_t* numSp)
{
	PathBuf* p = &t_path;
	pb_reset(p, scale, tol);
	if (pb_replay(p, cmd, ncmd, data) < 0) return -1;
	*numPts = p->n;
	*numSp = p->nsp;
	if (p->n > ptsCap || p->nsp > spCap) return -5;
	if (p->n) memcpy(pts, p->v, sizeof(v2) * p->n);
	if (p->nsp) memcpy(sp, p->sp, sizeof(vgb_subpath) * p->nsp);
	return 0;
}

int vgo_mesh(uint32_t op, const float* vtx, uint32_t n, int closed, uint32_t color, float width, uint32_t cap, uint32_t join,
             float scale, float tol, float fringe,
             float* pos, uint32_t* col, uint16_t* idx, uint32_t vCap, uint32_t iCap, uint32_t* nv, uint32_t* ni, int* hasColor)
{
	const StrokerCfg cfg = { scale, tol, fringe };
	MeshBuf* m = &t_mesh;
	*hasColor = run_mesh(m, &cfg, op, (const v2*)vtx, n, closed, color, width, cap, join);
	*nv = m->nv;
	*ni = m->ni;
	if (m->nv > vCap || m->ni > iCap) return -5;
	if (m->nv) memcpy(pos, m->pos, sizeof(v2) * m->nv);
	if (*hasColor && col && m->nv) memcpy(col, m->col, sizeof(uint32_t) * m->nv);
	if (m->ni) memcpy(idx, m->idx, sizeof(uint16_t) * m->ni);
	return 0;
}

int vgo_run_batch(uint32_t pathBegin, uint32_t pathEnd,
                  const uint32_t* pathCmdOff, const uint32_t* pathDataOff, const uint8_t* cmd, const float* data,
                  const float* pathScale, const float* pathTol, const float* pathXform, float fringe,
                  const uint32_t* drawBegin, const uint32_t* drawOp, const uint32_t* drawColor, const float* drawWidth,
                  const uint8_t* drawCap, const uint8_t* drawJoin,
                  int store, float* pos, uint32_t* col, uint16_t* idx, uint64_t vCap, uint64_t iCap,
                  uint64_t* meshVtxOff, uint64_t* meshIdxOff, uint32_t meshCap, uint64_t* totals)
{
	PathBuf* p = &t_path;
	MeshBuf* m = &t_mesh;
	uint64_t nv = 0, ni = 0, npts = 0, nmesh = 0, nsub = 0, sum = 0;
	for (uint32_t q = pathBegin; q < pathEnd; ++q) {
		const StrokerCfg cfg = { pathScale[q], pathTol[q], fringe };
		pb_reset(p, pathScale[q], pathTol[q]);
		if (pb_replay(p, cmd + pathCmdOff[q], pathCmdOff[q + 1] - pathCmdOff[q], data + pathDataOff[q]) < 0) return -1;
		const v2* v = p->v;
		if (pathXform) {
			/* vg.cpp:4957-4975 with the scalar transformPos2D (vg_util.h:24-28) */
			const float* t = pathXform + 6 * (size_t)q;
			if (p->n > t_xfcap) {
				t_xfcap = p->n * 2 + 64;
				t_xf = (v2*)realloc(t_xf, sizeof(v2) * t_xfcap);
			}
			for (uint32_t i = 0; i < p->n; ++i) {
				t_xf[i].x = t[0] * p->v[i].x + t[2] * p->v[i].y + t[4];
				t_xf[i].y = t[1] * p->v[i].x + t[3] * p->v[i].y + t[5];
			}
			v = t_xf;
		}
		npts += p->n;
		nsub += p->nsp;
		for (uint32_t d = drawBegin[q]; d < drawBegin[q + 1]; ++d) {
			for (uint32_t s = 0; s < p->nsp; ++s) {
				int hasCol = 0;
				m->nv = 0;
				m->ni = 0;
				if (p->sp[s].num_vertices >= min_verts(drawOp[d])) {
					hasCol = run_mesh(m, &cfg, drawOp[d], v + p->sp[s].first_vertex, p->sp[s].num_vertices, p->sp[s].is_closed,
					                  drawColor[d], drawWidth[d], drawCap[d], drawJoin[d]);
				}
				if (store) {
					if (nmesh >= meshCap || nv + m->nv > vCap || ni + m->ni > iCap) return -5;
					meshVtxOff[nmesh] = nv;
					meshIdxOff[nmesh] = ni;
					if (m->nv) {
						memcpy(pos + 2 * nv, m->pos, sizeof(v2) * m->nv);
						if (hasCol) memcpy(col + nv, m->col, sizeof(uint32_t) * m->nv);
						else for (uint32_t k = 0; k < m->nv; ++k) col[nv + k] = drawColor[d];
					}
					if (m->ni) memcpy(idx + ni, m->idx, sizeof(uint16_t) * m->ni);
				}
				if (m->ni) sum = sum * 1099511628211ull + m->idx[m->ni - 1] + m->nv;
				nv += m->nv;
				ni += m->ni;
				++nmesh;
			}
		}
	}
	if (store) {
		meshVtxOff[nmesh] = nv;
		meshIdxOff[nmesh] = ni;
	}
	totals[0] = nv; totals[1] = ni; totals[2] = npts; totals[3] = nmesh; totals[4] = sum; totals[5] = nsub;
	return 0;
}

/* array evaluation of the shared fp32 math (tests/test_math.py): op 0 sin, 1 cos, 2 acos, 3 atan2(y,x), 4 tan */
void vgo_math_eval(uint32_t op, const float* x, const float* y, float* out, uint32_t n)
{
	for (uint32_t i = 0; i < n; ++i) {
		switch (op) {
		case 0: out[i] = vgb_sin(x[i]); break;
		case 1: out[i] = vgb_cos(x[i]); break;
		case 2: out[i] = vgb_acos(x[i]); break;
		case 3: out[i] = vgb_atan2(y[i], x[i]); break;
		default: out[i] = vgb_tan(x[i]); break;
		}
	}
}

/* ---- final layout (vg.cpp:5207-5244, 5321-5405) -------------------------------
 * Restatement of the merge step for colour meshes: first-fit packing of the meshes into vertex
 * buffers of max_vb_vertices, one draw command per buffer the batch touches, uint16 indices
 * rebased by the vertices that precede the mesh in its command, white-pixel uv for every vertex. */
/* meshScissor (may be NULL): per mesh, scissor id | 1u << 31 when a scissor command ran between the previous draw
 * and the one this mesh opens (m_ForceNewDrawCommand, vg.cpp:3946-4033): the flag stays up until a mesh with
 * vertices reaches allocDrawCommand (vg.cpp:5368, cleared at 5402), which then opens a command in the same buffer */
int vgo_final_layout_sc(const uint64_t* meshVtxOff, const uint64_t* meshIdxOff, uint32_t numMeshes, const uint16_t* idxLocal,
                        const vgb_final_cfg* cfg, const uint32_t* meshScissor, uint16_t* idxOut, int16_t* uvOut, vgb_draw_cmd* cmds,
                        uint32_t cmdCap, uint32_t* numCmds)
{
	uint32_t vb = cfg->vb_id, inVb = cfg->vb_count, inIb = cfg->ib_count, inCmd = cfg->cmd_vertices;
	int open = 0; /* a record for the current command exists */
	int force = 0;
	uint32_t n = 0;
	for (uint32_t m = 0; m < numMeshes; ++m) {
		const uint64_t vo = meshVtxOff[m] - meshVtxOff[0], io = meshIdxOff[m] - meshIdxOff[0];
		const uint32_t nv = (uint32_t)(meshVtxOff[m + 1] - meshVtxOff[m]), ni = (uint32_t)(meshIdxOff[m + 1] - meshIdxOff[m]);
		if (meshScissor != NULL && (meshScissor[m] >> 31) != 0) {
			force = 1;
		}
		if (nv == 0) {
			continue; /* never submitted: vg.cpp:3101, 3450 */
		}
		if (nv >= cfg->max_vb_vertices) {
			return -5; /* vg.cpp:5323 */
		}
		if (inVb + nv > cfg->max_vb_vertices) { /* vg.cpp:5327 */
			vb += 1;
			inVb = 0;
			inCmd = 0;
			open = 0;
		}
		if (force) { /* vg.cpp:5368: no merging with the previous command */
			inCmd = 0;
			open = 0;
			force = 0;
		}
		if (!open) {
			if (n == cmdCap) {
				return -5;
			}
			vgb_draw_cmd* c = &cmds[n++];
			memset(c, 0, sizeof(*c));
			c->vb_id = vb;
			c->first_vertex = inVb;
			c->first_index = inIb;
			c->first_mesh = m;
			c->scissor_id = meshScissor != NULL ? (meshScissor[m] & 0x7fffffffu) : 0u; /* vg.cpp:5392-5395 */
			c->stream_vertex = vo;
			c->stream_index = io;
			open = 1;
		}
		for (uint32_t i = 0; i < nv; ++i) { /* vg.cpp:5220-5227 */
			uvOut[2 * (vo + i)] = cfg->uv[0];
			uvOut[2 * (vo + i) + 1] = cfg->uv[1];
		}
		for (uint32_t i = 0; i < ni; ++i) { /* vg.cpp:5240, vg_util.cpp:447-520 */
			idxOut[io + i] = (uint16_t)(idxLocal[io + i] + (uint16_t)inCmd);
		}
		inVb += nv;
		inIb += ni;
		inCmd += nv;
		cmds[n - 1].num_vertices += nv;
		cmds[n - 1].num_indices += ni;
		cmds[n - 1].num_meshes = m + 1 - cmds[n - 1].first_mesh;
	}
	*numCmds = n;
	return 0;
}

int vgo_final_layout(const uint64_t* meshVtxOff, const uint64_t* meshIdxOff, uint32_t numMeshes, const uint16_t* idxLocal,
                     const vgb_final_cfg* cfg, uint16_t* idxOut, int16_t* uvOut, vgb_draw_cmd* cmds, uint32_t cmdCap, uint32_t* numCmds)
{
	return vgo_final_layout_sc(meshVtxOff, meshIdxOff, numMeshes, idxLocal, cfg, NULL, idxOut, uvOut, cmds, cmdCap, numCmds);
}

/* vgutil::transformPos2D per vertex (vg_util.h:24-28, scalar association; -ffp-contract=off) */
void vgo_transform_positions(const float* pos, uint32_t n, const float* mtx, float* out)
{
	for (uint32_t i = 0; i < n; ++i) {
		const float x = pos[2 * i], y = pos[2 * i + 1];
		out[2 * i] = mtx[0] * x + mtx[2] * y + mtx[4];
		out[2 * i + 1] = mtx[1] * x + mtx[3] * y + mtx[5];
	}
}
