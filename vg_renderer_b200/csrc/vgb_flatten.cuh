/*
 * vgb_flatten.cuh -- device-side path flattening: one thread walks one SUB-PATH's commands.
 *
 * Why a thread per sub-path: every sequential dependence of the reference's Path object
 * (P1 of a cubic = last KEPT vertex, greedy epsilon-dedupe against the last kept vertex,
 * pathClose dropping a coincident end point; src/path.cpp:93-97, 761-784, 707-726) is confined
 * to one sub-path, so walking a sub-path in order inside one thread reproduces the reference
 * sequence exactly, while the millions of sub-paths of a batch run in parallel.  The walk runs
 * twice from the same code (EMIT=false: count -> exclusive scan -> EMIT=true: store), so
 * every sub-path's polyline lands at its scanned offset with no atomics and no gaps.
 *
 * Arithmetic: every product/sum is individually rounded (file is compiled with -fmad=false
 * and uses vgb_math.h for transcendental functions), so points are bit-identical to a
 * non-FMA x86-64 build of the reference.
 */
#ifndef VGB_FLATTEN_CUH
#define VGB_FLATTEN_CUH

#include "vgb_math.h"
#include "vgb.h"

#ifndef VGB_FLAT_CONVERGE
#define VGB_FLAT_CONVERGE 1
#endif

namespace vgb
{
/* number of floats each command consumes (POLYLINE: 1 + 2n, handled separately) */
__host__ __device__ __forceinline__ uint32_t cmdArity(uint32_t c)
{
	/* 2,2,6,4,5,4,5,8,3,4,6,(1),0 packed 4 bits each */
	return (uint32_t)((0x0164385454622ull >> (4 * c)) & 0xF);
}

template<bool EMIT>
struct SubPathWalker
{
	float lx, ly;   /* last vertex in the array (local space)        */
	float px, py;   /* the one before it (needed when close drops the last one) */
	float fx, fy;   /* first vertex                                   */
	uint32_t n;
	bool closed;
	float scale, tol;
	const float* m; /* 2x3 transform or nullptr                       */
	float2* out;
	uint32_t limit; /* EMIT: this sub-path's final vertex count (from the count pass). A vertex
	                 * that pathClose later drops sits at index == limit, i.e. in the NEXT
	                 * sub-path's first slot, and must never be stored. */

	__device__ __forceinline__ void push(float x, float y)
	{
		if (EMIT && n < limit) {
			float2 t;
			if (m) {
				/* vgutil::transformPos2D, src/vg_util.h:24-28 (scalar association) */
				t.x = m[0] * x + m[2] * y + m[4];
				t.y = m[1] * x + m[3] * y + m[5];
			} else {
				t.x = x;
				t.y = y;
			}
			out[n] = t;
		}
		if (n == 0) {
			fx = x;
			fy = y;
		}
		px = lx;
		py = ly;
		lx = x;
		ly = y;
		++n;
	}

	/* pathAddVertex, src/path.cpp:761-784 */
	__device__ __forceinline__ void add(float x, float y)
	{
		if (n != 0) {
			const float dx = lx - x;
			const float dy = ly - y;
			const float distSqr = dx * dx + dy * dy;
			if (distSqr < VGB_EPSILON) {
				return;
			}
		}
		push(x, y);
	}

	/* pathCubicTo, src/path.cpp:86-182: left-first subdivision, at most 10 pending right
	 * siblings; a non-flat node met with a full stack is dropped without a vertex.
	 * The pending siblings live in SHARED memory, laid out [level][component][thread] so that a
	 * warp's accesses are conflict-free (a per-thread local array would spill 320 B x 1M threads
	 * through L1 into L2).  Consecutive pending siblings share an end point (P4 of entry k is P1
	 * of entry k-1, the deepest entry ends at the curve's end point), so an entry stores 6 floats. */
	float* stk;      /* this thread's column: stk[(level * 6 + c) * STRIDE] */
	uint32_t stride;

	__device__ void cubic(float c1x, float c1y, float c2x, float c2y, float x, float y)
	{
		int sp = 0;
		float x1 = lx, y1 = ly, x2 = c1x, y2 = c1y, x3 = c2x, y3 = c2y, x4 = x, y4 = y;
		const float tessTol = tol / (scale * scale);
		const uint32_t S = stride;
#if VGB_FLAT_CONVERGE
		/* The lanes that enter together stay in the loop until all of them are done and meet at a vote once
		 * per node: without it the two sides of the flat / split branch loop back on their own and the
		 * flatness test itself runs twice per iteration with half the lanes each (ncu source page: 12 of 32
		 * lanes at the loop head). */
		const unsigned mask = __activemask();
		bool done = false;
#endif
		for (;;) {
#if VGB_FLAT_CONVERGE
			if (!done) {
#endif
			const float dx = x4 - x1;
			const float dy = y4 - y1;
			const float d2 = fabsf((x2 - x4) * dy - (y2 - y4) * dx);
			const float d3 = fabsf((x3 - x4) * dy - (y3 - y4) * dx);
			const float d23 = d2 + d3;
			const bool flat = d23 * d23 <= tessTol * (dx * dx + dy * dy);
			if (flat) {
				add(x4, y4);
			}
			if (!flat && sp < 10) {
				const float x12 = (x1 + x2) * 0.5f, y12 = (y1 + y2) * 0.5f;
				const float x23 = (x2 + x3) * 0.5f, y23 = (y2 + y3) * 0.5f;
				const float x34 = (x3 + x4) * 0.5f, y34 = (y3 + y4) * 0.5f;
				const float x123 = (x12 + x23) * 0.5f, y123 = (y12 + y23) * 0.5f;
				const float x234 = (x23 + x34) * 0.5f, y234 = (y23 + y34) * 0.5f;
				const float x1234 = (x123 + x234) * 0.5f, y1234 = (y123 + y234) * 0.5f;
				/* pending right half (P1234, P234, P34, [P4 = current P4]) */
				float* e = stk + (uint32_t)sp * 6u * S;
				e[0] = x1234; e[S] = y1234;
				e[2 * S] = x234; e[3 * S] = y234;
				e[4 * S] = x34;  e[5 * S] = y34;
				++sp;
				x2 = x12; y2 = y12; x3 = x123; y3 = y123; x4 = x1234; y4 = y1234;
			} else if (sp == 0) {
				/* only reachable from a flat node: a non-flat root always has sp<10 */
#if VGB_FLAT_CONVERGE
				done = true;
#else
				break;
#endif
			} else {
				--sp;
				const float* e = stk + (uint32_t)sp * 6u * S;
				/* the popped sibling ends where the entry below it starts, or at the curve's end */
				if (sp > 0) {
					x4 = e[-6 * (int)S];
					y4 = e[-5 * (int)S];
				} else {
					x4 = x;
					y4 = y;
				}
				x1 = e[0]; y1 = e[S];
				x2 = e[2 * S]; y2 = e[3 * S];
				x3 = e[4 * S]; y3 = e[5 * S];
			}
#if VGB_FLAT_CONVERGE
			}
			if (__all_sync(mask, done)) {
				break;
			}
#endif
		}
	}

	/* pathClose, src/path.cpp:707-726 */
	__device__ __forceinline__ void close()
	{
		if (closed || n <= 2) {
			return;
		}
		closed = true;
		const float dx = lx - fx;
		const float dy = ly - fy;
		const float distSqr = dx * dx + dy * dy;
		if (distSqr < VGB_EPSILON) {
			--n;
			lx = px;
			ly = py;
		}
	}

	/* quarter circle by rotation recurrence, src/path.cpp:319-337 etc.; points bypass dedupe */
	__device__ __forceinline__ void quarter(float cx, float cy, float r, float ca, float sa, float cosd, float sind, uint32_t nQ)
	{
		for (uint32_t i = 1; i < nQ; ++i) {
			const float ns = sind * ca + cosd * sa;
			const float nc = cosd * ca - sind * sa;
			ca = nc;
			sa = ns;
			push(cx + r * ca, cy + r * sa);
		}
	}

	/* pathRect, src/path.cpp:275-286 */
	__device__ void rect(float x, float y, float w, float h)
	{
		if (fabsf(w) < VGB_EPSILON || fabsf(h) < VGB_EPSILON) {
			return;
		}
		/* a RECT command always starts its own sub-path record, so n == 0 here */
		add(x, y);
		add(x, y + h);
		add(x + w, y + h);
		add(x + w, y);
		close();
	}

	/* pathEllipse / pathCircle, src/path.cpp:561-631 */
	__device__ void ellipse(float cx, float cy, float rx, float ry)
	{
		const float avgR = (rx + ry) * 0.5f;
		const float da = vgb_acos((scale * avgR) / ((scale * avgR) + tol)) * 2.0f;
		const uint32_t nHalf = vgb_umax(2u, vgb_f2u(vgb_ceil(VGB_PI / da)));
		const uint32_t numPoints = nHalf * 2;
		add(cx + rx, cy);
		const float dtheta = -VGB_PI2 / (float)numPoints;
		const float cosd = vgb_cos(dtheta);
		const float sind = vgb_sin(dtheta);
		float ca = 1.0f, sa = 0.0f;
		for (uint32_t i = 1; i < numPoints; ++i) {
			const float ns = sind * ca + cosd * sa;
			const float nc = cosd * ca - sind * sa;
			ca = nc;
			sa = ns;
			push(cx + rx * ca, cy + ry * sa);
		}
		close();
	}

	/* pathRoundedRect, src/path.cpp:288-409 */
	__device__ void roundedRect(float x, float y, float w, float h, float r)
	{
		if (r < 0.1f) {
			rect(x, y, w, h);
			return;
		}
		const float max_r = fminf(w, h) * 0.5f;
		if (w == h && r >= max_r - VGB_EPSILON) {
			ellipse(x + max_r, y + max_r, max_r, max_r);
			return;
		}
		r = fminf(r, max_r);
		const float da = vgb_acos((scale * r) / ((scale * r) + tol)) * 2.0f;
		const uint32_t nHalf = vgb_umax(2u, vgb_f2u(vgb_ceil(VGB_PI / da)));
		const uint32_t nQ = (nHalf >> 1) + 1;
		const float dtheta = -VGB_PI_HALF / (float)(nQ - 1);
		const float cosd = vgb_cos(dtheta);
		const float sind = vgb_sin(dtheta);
		add(x, y + r);
		add(x, y + h - r);
		quarter(x + r, y + h - r, r, -1.0f, 0.0f, cosd, sind, nQ);
		add(x + w - r, y + h);
		quarter(x + w - r, y + h - r, r, 0.0f, 1.0f, cosd, sind, nQ);
		add(x + w, y + r);
		quarter(x + w - r, y + r, r, 1.0f, 0.0f, cosd, sind, nQ);
		add(x + r, y);
		quarter(x + r, y + r, r, 0.0f, -1.0f, cosd, sind, nQ);
		close();
	}

	/* one corner of pathRoundedRectVarying, src/path.cpp:426-556 */
	__device__ __forceinline__ void varyingCorner(float cx, float cy, float r, float ca, float sa)
	{
		const float halfDa = vgb_acos((scale * r) / ((scale * r) + tol));
		const uint32_t nHalf = vgb_umax(2u, vgb_f2u(vgb_ceil(VGB_PI_HALF / halfDa)));
		const uint32_t nQ = (nHalf >> 1) + 1;
		const float dtheta = -VGB_PI_HALF / (float)(nQ - 1);
		quarter(cx, cy, r, ca, sa, vgb_cos(dtheta), vgb_sin(dtheta), nQ);
	}

	/* pathRoundedRectVarying, src/path.cpp:411-559 */
	__device__ void roundedRectVarying(float x, float y, float w, float h, float rTL, float rTR, float rBR, float rBL)
	{
		if (rTL < 0.1f && rBL < 0.1f && rBR < 0.1f && rTR < 0.1f) {
			rect(x, y, w, h);
			return;
		}
		const float halfw = w * 0.5f;
		const float halfh = h * 0.5f;
		const float rtl = fminf(fminf(rTL, halfw), halfh);
		const float rtr = fminf(fminf(rTR, halfw), halfh);
		const float rbl = fminf(fminf(rBL, halfw), halfh);
		const float rbr = fminf(fminf(rBR, halfw), halfh);
		if (rtl < 0.1f) {
			add(x, y);
		} else {
			add(x + rtl, y);
			varyingCorner(x + rtl, y + rtl, rtl, 0.0f, -1.0f);
		}
		if (rbl < 0.1f) {
			add(x, y + h);
		} else {
			add(x, y + h - rbl);
			varyingCorner(x + rbl, y + h - rbl, rbl, -1.0f, 0.0f);
		}
		if (rbr < 0.1f) {
			add(x + w, y + h);
		} else {
			add(x + w - rbr, y + h);
			varyingCorner(x + w - rbr, y + h - rbr, rbr, 0.0f, 1.0f);
		}
		if (rtr < 0.1f) {
			add(x + w, y);
		} else {
			add(x + w, y + rtr);
			varyingCorner(x + w - rtr, y + rtr, rtr, 1.0f, 0.0f);
		}
		close();
	}

	/* pathArc, src/path.cpp:633-682 */
	__device__ void arc(float cx, float cy, float r, float a0, float a1, uint32_t dir)
	{
		while (a0 > VGB_PI2) a0 -= VGB_PI2;
		while (a1 > VGB_PI2) a1 -= VGB_PI2;
		if (dir == VGB_WINDING_CCW) {
			while (a0 < a1) a0 += VGB_PI2;
		} else {
			while (a1 < a0) a1 += VGB_PI2;
		}
		const float da = vgb_acos((scale * r) / ((scale * r) + tol)) * 2.0f;
		const uint32_t numPoints = vgb_umax(2u, vgb_f2u(vgb_ceil(fabsf(a1 - a0) / da)));
		const float dtheta = (a1 - a0) / (float)numPoints;
		const float cosd = vgb_cos(dtheta);
		const float sind = vgb_sin(dtheta);
		float ca = vgb_cos(a0);
		float sa = vgb_sin(a0);
		add(cx + r * ca, cy + r * sa); /* lineTo or moveTo: both end in pathAddVertex */
		for (uint32_t i = 0; i < numPoints; ++i) {
			const float ns = sind * ca + cosd * sa;
			const float nc = cosd * ca - sind * sa;
			ca = nc;
			sa = ns;
			push(cx + r * ca, cy + r * sa);
		}
	}

	/* pathArcTo, src/path.cpp:203-273 */
	__device__ void arcTo(float x1, float y1, float x2, float y2, float r)
	{
		const float x0 = lx, y0 = ly;
		float dx0 = x0 - x1, dy0 = y0 - y1, dx1 = x2 - x1, dy1 = y2 - y1;
		{
			const float lenSqr = dx0 * dx0 + dy0 * dy0;
			const float invLen = lenSqr < VGB_EPSILON ? 0.0f : vgb_rsqrt(lenSqr);
			dx0 *= invLen;
			dy0 *= invLen;
		}
		{
			const float lenSqr = dx1 * dx1 + dy1 * dy1;
			const float invLen = lenSqr < VGB_EPSILON ? 0.0f : vgb_rsqrt(lenSqr);
			dx1 *= invLen;
			dy1 *= invLen;
		}
		const float a = vgb_acos(dx0 * dx1 + dy0 * dy1);
		const float d = r / vgb_tan(a / 2.0f);
		if (d > 10000.0f) {
			add(x1, y1);
			return;
		}
		float cx, cy, a0, a1;
		uint32_t dir;
		const float cross = dx1 * dy0 - dx0 * dy1;
		if (cross > 0.0f) {
			cx = x1 + dx0 * d + dy0 * r;
			cy = y1 + dy0 * d - dx0 * r;
			a0 = vgb_atan2(dx0, -dy0);
			a1 = vgb_atan2(-dx1, dy1);
			dir = VGB_WINDING_CW;
		} else {
			cx = x1 + dx0 * d - dy0 * r;
			cy = y1 + dy0 * d + dx0 * r;
			a0 = vgb_atan2(-dx0, dy0);
			a1 = vgb_atan2(dx1, -dy1);
			dir = VGB_WINDING_CCW;
		}
		arc(cx, cy, r, a0, a1, dir);
	}

	/* pathPolyline, src/path.cpp:684-705: only the first point is deduped */
	__device__ void polyline(const float* c, uint32_t np)
	{
		if (n > 0 && np > 0) {
			const float dx = lx - c[0];
			const float dy = ly - c[1];
			if (dx * dx + dy * dy < VGB_EPSILON) {
				c += 2;
				--np;
			}
		}
		for (uint32_t i = 0; i < np; ++i) {
			push(c[2 * i], c[2 * i + 1]);
		}
	}

	__device__ void run(const uint8_t* __restrict__ cmd, uint32_t c0, uint32_t c1, const float* __restrict__ d)
	{
		for (uint32_t c = c0; c < c1; ++c) {
			const uint32_t t = cmd[c];
			switch (t) {
			case VGB_CMD_MOVE_TO:
			case VGB_CMD_LINE_TO: add(d[0], d[1]); break;
			case VGB_CMD_CUBIC_TO: cubic(d[0], d[1], d[2], d[3], d[4], d[5]); break;
			case VGB_CMD_QUAD_TO: {
				/* pathQuadraticTo, src/path.cpp:184-201 */
				const float x0 = lx, y0 = ly, cx = d[0], cy = d[1], x = d[2], y = d[3];
				const float c1x = x0 + (2.0f / 3.0f) * (cx - x0);
				const float c1y = y0 + (2.0f / 3.0f) * (cy - y0);
				const float c2x = x + (2.0f / 3.0f) * (cx - x);
				const float c2y = y + (2.0f / 3.0f) * (cy - y);
				cubic(c1x, c1y, c2x, c2y, x, y);
			} break;
			case VGB_CMD_ARC_TO: arcTo(d[0], d[1], d[2], d[3], d[4]); break;
			case VGB_CMD_RECT: rect(d[0], d[1], d[2], d[3]); break;
			case VGB_CMD_ROUNDED_RECT: roundedRect(d[0], d[1], d[2], d[3], d[4]); break;
			case VGB_CMD_ROUNDED_RECT_VARYING: roundedRectVarying(d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]); break;
			case VGB_CMD_CIRCLE: ellipse(d[0], d[1], d[2], d[2]); break;
			case VGB_CMD_ELLIPSE: ellipse(d[0], d[1], d[2], d[3]); break;
			case VGB_CMD_ARC: arc(d[0], d[1], d[2], d[3], d[4], d[5] != 0.0f ? VGB_WINDING_CW : VGB_WINDING_CCW); break;
			case VGB_CMD_POLYLINE: {
				const uint32_t np = __float_as_uint(d[0]);
				polyline(d + 1, np);
				d += 2 * (size_t)np;
			} break;
			case VGB_CMD_CLOSE: close(); break;
			default: break;
			}
			d += (t == VGB_CMD_POLYLINE) ? 1u : cmdArity(t);
		}
	}
};
} // namespace vgb

#endif
