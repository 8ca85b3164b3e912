/*
 * vgb_stroke.cuh -- device-side mesh generation, expressed per PIECE.
 *
 * A mesh (one strokerXXX call of the reference) is cut into pieces whose outputs are
 * contiguous, in piece order, in the mesh's vertex AND index streams:
 *   AA stroke of an N-point polyline (stroker.cpp:1390-1991): piece i = polyline vertex i
 *       open:   0 = start cap, 1..N-2 = joins, N-1 = end cap
 *       closed: 0..N-1 = joins (wrap-around); the last one also emits the closing connect
 *   convex AA fill of an N-gon (stroker.cpp:713-806, scalar variant = golden):
 *       pieces 0..N-3 = fan triangles (3 indices, no vertex), then N vertex pieces
 *       (2 vertices + one fringe quad) -- exactly the reference's [fan][quads] index order.
 * Piece sizes depend on data (round joins: n = max(2, uint(dAngle/da))), so sizes are counted
 * (k_group_count), scanned, then emitted (k_group_emit) -- both call classify() below, which
 * is therefore the single place where every count decision lives.
 *
 * All float expressions follow the reference's evaluation order operation by operation and the
 * file is compiled with -fmad=false, so positions are bit-identical to a non-FMA x86-64 build.
 */
#ifndef VGB_STROKE_CUH
#define VGB_STROKE_CUH

#include "vgb_internal.h"
#include "vgb_math.h"

namespace vgb
{
enum { PK_NONE = 0, PK_CAP_START, PK_JOIN, PK_CAP_END, PK_FILL_FAN, PK_FILL_VTX, PK_FAN_TRI, PK_FAN_VTX };
enum { PF_LEFT = 1, PF_HAS_PREV = 2, PF_CLOSING = 4 };

/* one job = one mesh: (draw, sub-path) */
struct JobCtx
{
	uint32_t first, N, closed;
	uint32_t nfix; /* round joins of this job always have n == 2 when the turn is unambiguous (3*da > pi) */
	uint32_t sse;  /* convex AA fills in the SIMD build's index order (VGB_RUN_FILL_ORDER_SSE) */
	uint32_t op, cap, join;
	float hsw, hsw_aa, da, fringe;
	uint32_t nHC, color, c0;
};

struct Piece
{
	uint32_t kind, vcount, icount, n, flags, li;
	float2 p, d01, d12, v;
	float a01, arcDa;
};

__device__ __forceinline__ float2 f2(float x, float y) { return make_float2(x, y); }

/* vec2Dir, stroker.cpp:31-38 */
__device__ __forceinline__ float2 vdir(float2 a, float2 b)
{
	const float dx = b.x - a.x;
	const float dy = b.y - a.y;
	const float lenSqr = dx * dx + dy * dy;
	const float invLen = lenSqr < VGB_EPSILON ? 0.0f : vgb_rsqrt(lenSqr);
	return f2(dx * invLen, dy * invLen);
}

/* calcExtrusionVector, stroker.cpp:40-53 */
__device__ __forceinline__ float2 extrusion(float2 d01, float2 d12)
{
	float2 v = f2(-d01.y, d01.x);
	const float cross = d12.x * d01.y - d01.x * d12.y;
	if (fabsf(cross) > (1.0f / 100.0f)) {
		const float s = 1.0f / cross;
		v = f2((d01.x - d12.x) * s, (d01.y - d12.y) * s);
	}
	return v;
}

/* number of pieces of a job (0 = the caller skips this sub-path: vg.cpp:3101, 3450) */
__host__ __device__ __forceinline__ uint32_t pieceCount(uint32_t op, uint32_t N, bool sseFill = false)
{
	switch (op) {
	case VGB_OP_STROKE_AA:
	case VGB_OP_STROKE:
	case VGB_OP_STROKE_AA_THIN: return N >= 2 ? N : 0;
	case VGB_OP_CONVEX_FILL_AA: return N >= 3 ? (sseFill ? N : 2 * N - 2) : 0; /* SIMD order: vertex i carries F(i-1) and Q(i) */
	case VGB_OP_CONVEX_FILL: return N >= 3 ? 2 * N - 2 : 0;
	default: return 0;
	}
}

/* side test + arc subdivision of one join (stroker.cpp:1528-1535, 1581-1596, 1737-1752) */
/* cap = true: a ROUND CAP lane riding along (round joins only) for the one thing it shares with a
 * join, its start angle atan2(l.y, l.x), l = perpCCW(d) -- the right-inner form of a01 below; the
 * caller overwrites every other field afterwards */
template<int JT>
__device__ __forceinline__ void classifyJoinAA(const JobCtx& J, Piece& P, bool cap = false)
{
	const uint32_t jt = JT < 3 ? (uint32_t)JT : J.join; /* compile-time join type when the whole batch shares it */
	P.v = extrusion(P.d01, P.d12);
	const float vx = P.v.x * J.hsw_aa;
	const float vy = P.v.y * J.hsw_aa;
	const float proj = P.d12.x * vx + P.d12.y * vy;
	const bool left = !cap && proj >= 0.0f;
	if (left) {
		P.flags |= PF_LEFT;
	}
	const uint32_t conn = (P.flags & PF_HAS_PREV) ? 18u : 0u;
	if (jt == VGB_JOIN_MITER) {
		P.n = 0;
		P.vcount = 4;
		P.icount = conn;
	} else {
		uint32_t n = 1;
		if (jt == VGB_JOIN_ROUND) {
			/* left inner: r = perpCW(d), right inner: l = perpCCW(d) */
			const float2 s01 = left ? f2(P.d01.y, -P.d01.x) : f2(-P.d01.y, P.d01.x);
			const float2 s12 = left ? f2(P.d12.y, -P.d12.x) : f2(-P.d12.y, P.d12.x);
			const float a01 = vgb_atan2(s01.y, s01.x);
			float a12 = vgb_atan2(s12.y, s12.x);
			if (left) {
				if (a12 < a01) {
					a12 += VGB_PI2;
				}
				n = vgb_umax(2u, vgb_f2u((a12 - a01) / J.da));
			} else {
				if (a12 > a01) {
					a12 -= VGB_PI2;
				}
				n = vgb_umax(2u, vgb_f2u((a01 - a12) / J.da));
			}
			P.a01 = a01;
			P.arcDa = (a12 - a01) / (float)n;
		}
		P.n = n;
		P.vcount = 2 * n + 4;
		P.icount = conn + 9 * n;
	}
	if (P.flags & PF_CLOSING) {
		P.icount += 18;
	}
}

/* ---- plain (W=2) and thin AA (W=3) strokes: the same ribbon walk as the AA stroke with fewer ids
 * per cross-section (stroker.cpp:1008-1388 and 1993-2314).  JobCtx carries the op-specific side-test
 * scale in hsw_aa (plain: w/2, thin: fringe) and, for thin strokes, the already mapped cap/join
 * (round cap -> square, round join -> bevel: stroker.cpp:318-327). */
__device__ __forceinline__ void classifyRibbon(const JobCtx& J, uint32_t q, const float2* __restrict__ v, Piece& P)
{
	const uint32_t N = J.N;
	const bool thin = J.op == VGB_OP_STROKE_AA_THIN;
	const uint32_t W = thin ? 3u : 2u;
	const uint32_t conn = 6 * (W - 1);
	const uint32_t fanI = (J.cap == VGB_CAP_ROUND && !thin) ? 3 * (J.nHC - 2) : 0u;
	if (!J.closed && (q == 0 || q == N - 1)) {
		const bool end = q != 0;
		P.kind = end ? PK_CAP_END : PK_CAP_START;
		P.p = v[q];
		P.d01 = end ? vdir(v[N - 2], P.p) : vdir(P.p, v[1]);
		P.n = J.nHC;
		P.vcount = thin ? 3u : (J.cap == VGB_CAP_ROUND ? J.nHC : 2u);
		P.icount = end ? conn + fanI : fanI;
		if (end) {
			P.flags = PF_HAS_PREV;
		}
		return;
	}
	P.kind = PK_JOIN;
	P.p = v[q];
	P.d01 = vdir(v[q == 0 ? N - 1 : q - 1], P.p);
	P.d12 = vdir(P.p, v[q == N - 1 ? 0 : q + 1]);
	if (!(J.closed && q == 0)) {
		P.flags |= PF_HAS_PREV;
	}
	if (J.closed && q == N - 1) {
		P.flags |= PF_CLOSING;
	}
	P.v = extrusion(P.d01, P.d12);
	const float proj = P.d12.x * (P.v.x * J.hsw_aa) + P.d12.y * (P.v.y * J.hsw_aa); /* 1099 / 2072 */
	const bool left = proj >= 0.0f;
	if (left) {
		P.flags |= PF_LEFT;
	}
	const uint32_t c = (P.flags & PF_HAS_PREV) ? conn : 0u;
	if (J.join == VGB_JOIN_MITER) {
		P.vcount = W;
		P.icount = c;
	} else if (thin) {
		P.n = 1;
		P.vcount = 4;
		P.icount = c + 3;
	} else {
		uint32_t n = 1;
		if (J.join == VGB_JOIN_ROUND) {
			const float2 s01 = left ? f2(P.d01.y, -P.d01.x) : f2(-P.d01.y, P.d01.x);
			const float2 s12 = left ? f2(P.d12.y, -P.d12.x) : f2(-P.d12.y, P.d12.x);
			const float a01 = vgb_atan2(s01.y, s01.x);
			float a12 = vgb_atan2(s12.y, s12.x);
			if (left) {
				if (a12 < a01) a12 += VGB_PI2;
				n = vgb_umax(2u, vgb_f2u((a12 - a01) / J.da));
			} else {
				if (a12 > a01) a12 -= VGB_PI2;
				n = vgb_umax(2u, vgb_f2u((a01 - a12) / J.da));
			}
			P.a01 = a01;
			P.arcDa = (a12 - a01) / (float)n;
		}
		P.n = n;
		P.vcount = n + 2;
		P.icount = c + 3 * n;
	}
	if (P.flags & PF_CLOSING) {
		P.icount += conn;
	}
}

/* ops present in a batch (kernel template parameter: code for absent ops is compiled out) */
enum { OPS_STROKE_AA = 1, OPS_FILL_AA = 2, OPS_OTHER = 4, OPS_ALL = 7 };

/*
 * Count-only shortcut for AA strokes (k_group_count).  Sizes of caps, miter and bevel joins do not
 * depend on geometry; a round join has n = max(2, uint(theta / da)) with theta the turning angle
 * <= pi, so when 3*da > pi (J.nfix, decided in k_draw_prep: strokes thinner than ~4.2 px at the
 * default tolerance) n == 2 -- PROVIDED the reference's side test and its atan2 ordering agree,
 * which is certain when the turn is unambiguous: |sin(theta)| > 1e-3 and both segments longer than
 * the epsilon below which vec2Dir returns zero.  Then the side is the sign of the raw cross
 * product as well.  Anything else (near-collinear, hairpin, degenerate) returns false and takes
 * the full classify(), so counts stay bit-exact.  needSide: the caller needs PF_LEFT (the warp's
 * last lane publishes the hand-over word).  pm, p, pn: the piece's polyline point and its
 * wrapped neighbours, loaded by the caller (so that it can keep several groups' loads in flight).
 */
template<int JT, int CT>
__device__ __forceinline__ bool quickCountAA(const JobCtx& J, uint32_t q, float2 pm, float2 p, float2 pn, bool needSide, Piece& P)
{
	const uint32_t jt = JT < 3 ? (uint32_t)JT : J.join;
	const uint32_t ct = CT < 3 ? (uint32_t)CT : J.cap;
	if (jt == VGB_JOIN_ROUND && !J.nfix) {
		return false;
	}
	const uint32_t N = J.N;
	const bool open = !J.closed;
	const bool isStart = open && q == 0, isEnd = open && q == N - 1;
	P.kind = PK_JOIN;
	P.flags = 0;
	P.li = q;
	if (isStart || isEnd) {
		P.kind = isStart ? PK_CAP_START : PK_CAP_END;
		P.flags = isEnd ? PF_HAS_PREV : 0u;
		if (ct == VGB_CAP_ROUND) {
			P.n = J.nHC;
			P.vcount = 2 * J.nHC;
			P.icount = isStart ? 9 * J.nHC - 12 : 9 * J.nHC + 6;
		} else {
			P.n = 0;
			P.vcount = 4;
			P.icount = isStart ? 6u : 24u;
		}
		return true;
	}
	if (!(J.closed && q == 0)) {
		P.flags |= PF_HAS_PREV;
	}
	if (J.closed && q == N - 1) {
		P.flags |= PF_CLOSING;
	}
	if (jt == VGB_JOIN_ROUND || needSide) {
		const float ax = p.x - pm.x, ay = p.y - pm.y, bx = pn.x - p.x, by = pn.y - p.y;
		const float la = ax * ax + ay * ay, lb = bx * bx + by * by;
		const float cr = ax * by - ay * bx;
		if (!(la > 2e-5f && lb > 2e-5f && cr * cr > 1e-6f * (la * lb))) {
			return false;
		}
		if (cr > 0.0f) {
			P.flags |= PF_LEFT;
		}
	}
	const uint32_t conn = (P.flags & PF_HAS_PREV) ? 18u : 0u;
	const uint32_t n = jt == VGB_JOIN_MITER ? 0u : (jt == VGB_JOIN_BEVEL ? 1u : 2u);
	P.n = n;
	P.vcount = jt == VGB_JOIN_MITER ? 4u : 2 * n + 4;
	P.icount = conn + 9 * n + ((P.flags & PF_CLOSING) ? 18u : 0u);
	return true;
}

/* AA-stroke piece q of job J from its point and the two segment directions around it (dA: incoming,
 * dB: outgoing; a start cap ignores dA, an end cap dB).  One code path whatever the piece is. */
template<int JT, int CT>
__device__ __forceinline__ Piece classifyStrokeAA(const JobCtx& J, uint32_t q, float2 p, float2 dA, float2 dB)
{
	Piece P;
	P.vcount = 0;
	P.icount = 0;
	P.n = 0;
	P.flags = 0;
	P.li = q;
	P.a01 = 0.0f;
	P.arcDa = 0.0f;
	P.d12 = P.v = f2(0.0f, 0.0f);
	P.p = p;
	const uint32_t N = J.N;
	const uint32_t ct = CT < 3 ? (uint32_t)CT : J.cap;
	const uint32_t jt = JT < 3 ? (uint32_t)JT : J.join;
	const bool open = !J.closed;
	const bool isStart = open && q == 0, isEnd = open && q == N - 1;
	const bool isCap = isStart || isEnd;
	/* with round joins a round cap takes its start angle from the joins' atan2 (same expression) */
	const bool capRides = isCap && ct == VGB_CAP_ROUND && jt == VGB_JOIN_ROUND;
	if (isCap) {
		P.kind = isStart ? PK_CAP_START : PK_CAP_END;
		P.d01 = isStart ? dB : dA;
	} else {
		P.kind = PK_JOIN;
		P.d01 = dA;
		P.d12 = dB;
		if (!(J.closed && q == 0)) {
			P.flags |= PF_HAS_PREV;
		}
		if (J.closed && q == N - 1) {
			P.flags |= PF_CLOSING;
		}
	}
	if (!isCap || capRides) {
		classifyJoinAA<JT>(J, P, isCap);
	}
	if (isCap) {
		P.flags = isEnd ? PF_HAS_PREV : 0u;
		P.d12 = f2(0.0f, 0.0f);
		P.v = f2(0.0f, 0.0f);
		P.arcDa = 0.0f;
		if (ct == VGB_CAP_ROUND) {
			P.n = J.nHC;
			P.vcount = 2 * J.nHC;
			P.icount = isStart ? 9 * J.nHC - 12 : 9 * J.nHC + 6;
			if (!capRides) {
				P.a01 = vgb_atan2(P.d01.x, -P.d01.y); /* startAngle = atan2(l.y, l.x), l = perpCCW(d) (1476, 1919) */
			}
		} else {
			P.n = 0;
			P.a01 = 0.0f;
			P.vcount = 4;
			P.icount = isStart ? 6u : 24u;
		}
	}
	return P;
}

/* the single source of every count decision: piece q of job J */
template<uint32_t OPS, int JT = 3, int CT = 3>
__device__ __forceinline__ Piece classify(const JobCtx& J, uint32_t q, const float2* __restrict__ pts)
{
	const float2* v = pts + J.first;
	const uint32_t N = J.N;
	if ((OPS & OPS_STROKE_AA) && (!(OPS & ~OPS_STROKE_AA) || J.op == VGB_OP_STROKE_AA)) {
		/* a start cap uses the outgoing direction, an end cap the incoming one (its other neighbour
		 * index is clamped onto itself and gives the zero vector) */
		const bool open = !J.closed;
		const float2 pm = v[q > 0 ? q - 1 : (open ? 0u : N - 1)];
		const float2 pn = v[q < N - 1 ? q + 1 : (open ? q : 0u)];
		const float2 p = v[q];
		return classifyStrokeAA<JT, CT>(J, q, p, vdir(pm, p), vdir(p, pn));
	}
	Piece P;
	P.kind = PK_NONE;
	P.vcount = 0;
	P.icount = 0;
	P.n = 0;
	P.flags = 0;
	P.li = q;
	P.a01 = 0.0f;
	P.arcDa = 0.0f;
	P.p = P.d01 = P.d12 = P.v = f2(0.0f, 0.0f);
	if (false) {
	} else if ((OPS & OPS_FILL_AA) && (!(OPS & ~OPS_FILL_AA) || J.op == VGB_OP_CONVEX_FILL_AA)) {
		if (!J.sse && q < N - 2) {
			P.kind = PK_FILL_FAN;
			P.icount = 3;
		} else {
			const uint32_t i = J.sse ? q : q - (N - 2);
			P.kind = PK_FILL_VTX;
			P.li = i;
			P.vcount = 2;
			P.icount = (J.sse && i > 0 && i + 1 < N) ? 9u : 6u;
			const float2 pm = v[i == 0 ? N - 1 : i - 1];
			const float2 pn = v[i == N - 1 ? 0 : i + 1];
			P.p = v[i];
			P.d01 = vdir(pm, P.p);
			P.d12 = vdir(P.p, pn);
		}
	} else if ((OPS & OPS_OTHER) && (J.op == VGB_OP_STROKE || J.op == VGB_OP_STROKE_AA_THIN)) {
		classifyRibbon(J, q, v, P);
	} else if ((OPS & OPS_OTHER) && J.op == VGB_OP_CONVEX_FILL) {
		/* strokerConvexFill, stroker.cpp:334-365: indices (0,k,k+1); positions are the polygon itself */
		if (q < N - 2) {
			P.kind = PK_FAN_TRI;
			P.icount = 3;
		} else {
			P.kind = PK_FAN_VTX;
			P.li = q - (N - 2);
			P.vcount = 1;
			P.p = v[P.li];
		}
	}
	return P;
}

/* A piece's hand-over to its successor packed in one word: kind | left << 3 | n << 4 | vcount << 16.
 * The successor rebuilds the four ids (LeftAA, Left, Right, RightAA) from it and its own F
 * (F_prev = F - vcount_prev). */
__device__ __forceinline__ uint32_t packTail(const Piece& P)
{
	return P.kind | ((P.flags & PF_LEFT) ? 8u : 0u) | ((P.n & 0xFFFu) << 4) | (P.vcount << 16);
}

__device__ __forceinline__ uint4 out4FromTail(uint32_t capType, uint32_t joinType, uint32_t tail, uint32_t F)
{
	const uint32_t kind = tail & 7u, n = (tail >> 4) & 0xFFFu;
	const uint32_t Fp = F - (tail >> 16);
	if (kind == PK_CAP_START) {
		if (capType == VGB_CAP_ROUND) {
			const uint32_t e = 2 * (n - 1);
			return make_uint4(Fp + 1, Fp, Fp + e, Fp + e + 1);
		}
		return make_uint4(Fp, Fp + 1, Fp + 2, Fp + 3);
	}
	const uint32_t n2 = (joinType == VGB_JOIN_MITER) ? 0u : 2 * n;
	if (tail & 8u) {
		return make_uint4(Fp, Fp + 1, Fp + 2 + n2, Fp + 3 + n2);
	}
	return make_uint4(Fp + 3 + n2, Fp + 2 + n2, Fp + 1, Fp);
}

/* the 4 ids of a join's entry cross-section (what the previous piece connects to) */
__device__ __forceinline__ uint4 joinIn4(bool left, uint32_t F)
{
	return left ? make_uint4(F, F + 1, F + 2, F + 3) : make_uint4(F + 3, F + 2, F + 1, F);
}

/* ---- sinks: where a piece's vertices / colours / indices go -------------------------------------
 * Two of them with one interface, so that the address space of every store is known at compile time
 * (a sink whose pointers are chosen at run time between shared and global memory makes every store a
 * generic ST with 64-bit address arithmetic -- round-1 profile: 89 generic ST.E, no STS):
 *   Sink   scalar stores through plain pointers: the rare warp whose pieces exceed the staging image
 *          writes straight to HBM through it;
 *   SinkS  the warp-private staging image in SHARED memory (STS with immediate offsets); uint16 indices
 *          leave as packed 32-bit words: an index sequence is formed as "logical" words (i0,i1) (i2,i3) ..
 *          and a lane whose sequence starts at an odd uint16 position stores the funnel-shifted words
 *          (i1,i2) (i3,i4) .. plus the two odd halves, through the same code (the PRMT selector is the
 *          only thing that differs between lanes), so parity never diverges a warp. */
template<class IDX, bool ALIGNED = true>
struct Sink
{
	float2* pos;
	uint32_t* col;
	IDX* idx;
	uint32_t base; /* added to every index value (0 for per-mesh uint16) */
	uint32_t par;  /* unused here (SinkS: parity of idx) */

	__device__ __forceinline__ void V(uint32_t k, float x, float y, uint32_t c) const
	{
		pos[k] = make_float2(x, y);
		col[k] = c;
	}
	/* two consecutive vertices starting at an EVEN k: one 128-bit and one 64-bit store.  Every AA
	 * piece has an even vertex count, so every piece starts at an even vertex of its mesh and the
	 * staging image keeps that parity. */
	__device__ __forceinline__ void VP(uint32_t k, float x0, float y0, uint32_t c0_, float x1, float y1, uint32_t c1_) const
	{
		if (ALIGNED) {
			*reinterpret_cast<float4*>(pos + k) = make_float4(x0, y0, x1, y1);
			*reinterpret_cast<uint2*>(col + k) = make_uint2(c0_, c1_);
		} else {
			/* batches mixing odd-vertex-count ops (plain / thin strokes, fan fills) lose the parity */
			pos[k] = make_float2(x0, y0);
			pos[k + 1] = make_float2(x1, y1);
			col[k] = c0_;
			col[k + 1] = c1_;
		}
	}
	/* W-1 quads between two cross-sections of W ids (W=2: stroker.cpp:1118-1121, W=3: 2092-2097) */
	template<int W>
	__device__ __forceinline__ void bridge(uint32_t k, uint4 p, uint4 c) const
	{
		const uint32_t pa[4] = {p.x, p.y, p.z, p.w}, ca[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
		for (int i = 0; i + 1 < W; ++i) {
			tri2(k + 6 * i, pa[i], pa[i + 1], ca[i + 1], pa[i], ca[i + 1], ca[i]);
		}
	}
	__device__ __forceinline__ void I(uint32_t k, uint32_t v) const { idx[k] = (IDX)(v + base); }
	__device__ __forceinline__ void tri(uint32_t k, uint32_t a, uint32_t b, uint32_t c) const
	{
		I(k, a);
		I(k + 1, b);
		I(k + 2, c);
	}
	/* two consecutive triangles */
	__device__ __forceinline__ void tri2(uint32_t k, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f) const
	{
		tri(k, a, b, c);
		tri(k + 3, d, e, f);
	}
	/* the 18-index strip between two cross-sections (stroker.cpp:1557-1564 and siblings) */
	__device__ __forceinline__ void connect(uint32_t k, uint4 p, uint4 c) const
	{
		tri2(k + 0, p.x, p.y, c.y, p.x, c.y, c.x);
		tri2(k + 6, p.y, p.z, c.z, p.y, c.z, c.y);
		tri2(k + 12, p.z, p.w, c.w, p.z, c.w, c.z);
	}
	/* ... at the start of the piece's index sequence */
	__device__ __forceinline__ void connect0(uint4 p, uint4 c) const { connect(0, p, c); }
};

__device__ __forceinline__ uint32_t pk16(uint32_t lo, uint32_t hi) { return __byte_perm(lo, hi, 0x5410); }

template<class IDX, bool ALIGNED = true>
struct SinkS
{
	float2* pos;   /* all three point into the warp's staging image in shared memory */
	uint32_t* col;
	IDX* idx;
	uint32_t base;
	uint32_t par;  /* uint16 only: idx sits at an odd uint16 position (of its 32-bit word) */

	__device__ __forceinline__ void V(uint32_t k, float x, float y, uint32_t c) const
	{
		pos[k] = make_float2(x, y);
		col[k] = c;
	}
	__device__ __forceinline__ void VP(uint32_t k, float x0, float y0, uint32_t c0_, float x1, float y1, uint32_t c1_) const
	{
		if (ALIGNED) {
			*reinterpret_cast<float4*>(pos + k) = make_float4(x0, y0, x1, y1);
			*reinterpret_cast<uint2*>(col + k) = make_uint2(c0_, c1_);
		} else {
			pos[k] = make_float2(x0, y0);
			pos[k + 1] = make_float2(x1, y1);
			col[k] = c0_;
			col[k + 1] = c1_;
		}
	}
	template<int W>
	__device__ __forceinline__ void bridge(uint32_t k, uint4 p, uint4 c) const
	{
		const uint32_t pa[4] = {p.x, p.y, p.z, p.w}, ca[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
		for (int i = 0; i + 1 < W; ++i) {
			tri2(k + 6 * i, pa[i], pa[i + 1], ca[i + 1], pa[i], ca[i + 1], ca[i]);
		}
	}
	__device__ __forceinline__ void I(uint32_t k, uint32_t v) const { idx[k] = (IDX)(v + base); }

	/* NW logical words (2*NW - HALF indices, HALF = the last word carries one index only) at uint16 position k */
	template<int NW, bool HALF>
	__device__ __forceinline__ void words(uint32_t k, const uint32_t (&w)[NW]) const
	{
		const uint32_t odd = (par + k) & 1u;
		const uint32_t sel = odd ? 0x5432u : 0x3210u;
		uint16_t* q = reinterpret_cast<uint16_t*>(idx) + k;
		uint32_t* wq = reinterpret_cast<uint32_t*>(q + odd);
#pragma unroll
		for (int j = 0; j + 1 < NW; ++j) {
			wq[j] = __byte_perm(w[j], w[j + 1], sel);
		}
		if (HALF) {
			/* even: the lone index follows the words; odd: the first index precedes them (the lone one
			 * went out as the high half of the last funnel-shifted word) */
			const uint32_t h = odd ? w[0] : w[NW - 1];
			q[odd ? 0 : 2 * (NW - 1)] = (uint16_t)h;
		} else if (odd) {
			q[0] = (uint16_t)w[0];
			q[2 * NW - 1] = (uint16_t)(w[NW - 1] >> 16);
		} else {
			wq[NW - 1] = w[NW - 1];
		}
	}
	__device__ __forceinline__ void tri(uint32_t k, uint32_t a, uint32_t b, uint32_t c) const
	{
		if (sizeof(IDX) == 2) {
			a += base; b += base; c += base;
			const uint32_t w[2] = {pk16(a, b), c};
			words<2, true>(k, w);
		} else {
			I(k, a);
			I(k + 1, b);
			I(k + 2, c);
		}
	}
	__device__ __forceinline__ void tri2(uint32_t k, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f) const
	{
		if (sizeof(IDX) == 2) {
			a += base; b += base; c += base; d += base; e += base; f += base;
			const uint32_t w[3] = {pk16(a, b), pk16(c, d), pk16(e, f)};
			words<3, false>(k, w);
		} else {
			I(k, a);
			I(k + 1, b);
			I(k + 2, c);
			I(k + 3, d);
			I(k + 4, e);
			I(k + 5, f);
		}
	}
	/* 18 indices: p.x,p.y,c.y | p.x,c.y,c.x | p.y,p.z,c.z | p.y,c.z,c.y | p.z,p.w,c.w | p.z,c.w,c.z */
	__device__ __forceinline__ void connect(uint32_t k, uint4 p, uint4 c) const
	{
		if (sizeof(IDX) == 2) {
			p.x += base; p.y += base; p.z += base; p.w += base;
			c.x += base; c.y += base; c.z += base; c.w += base;
			const uint32_t w[9] = {pk16(p.x, p.y), pk16(c.y, p.x), pk16(c.y, c.x), pk16(p.y, p.z), pk16(c.z, p.y),
			                       pk16(c.z, c.y), pk16(p.z, p.w), pk16(c.w, p.z), pk16(c.w, c.z)};
			words<9, false>(k, w);
		} else {
			tri2(k + 0, p.x, p.y, c.y, p.x, c.y, c.x);
			tri2(k + 6, p.y, p.z, c.z, p.y, c.z, c.y);
			tri2(k + 12, p.z, p.w, c.w, p.z, c.w, c.z);
		}
	}
	/* the connect strip at the START of the piece's sequence (where joins and end caps have it): parity is the
	 * lane's own (par), every store at an immediate offset */
	__device__ __forceinline__ void connect0(uint4 p, uint4 c) const
	{
		if (sizeof(IDX) == 2) {
			p.x += base; p.y += base; p.z += base; p.w += base;
			c.x += base; c.y += base; c.z += base; c.w += base;
			const uint32_t sel = par ? 0x5432u : 0x3210u;
			uint16_t* q = reinterpret_cast<uint16_t*>(idx);
			uint32_t* wq = reinterpret_cast<uint32_t*>(q + par);
			const uint32_t w0 = pk16(p.x, p.y), w1 = pk16(c.y, p.x), w2 = pk16(c.y, c.x), w3 = pk16(p.y, p.z), w4 = pk16(c.z, p.y),
			               w5 = pk16(c.z, c.y), w6 = pk16(p.z, p.w), w7 = pk16(c.w, p.z), w8 = pk16(c.w, c.z);
			wq[0] = __byte_perm(w0, w1, sel);
			wq[1] = __byte_perm(w1, w2, sel);
			wq[2] = __byte_perm(w2, w3, sel);
			wq[3] = __byte_perm(w3, w4, sel);
			wq[4] = __byte_perm(w4, w5, sel);
			wq[5] = __byte_perm(w5, w6, sel);
			wq[6] = __byte_perm(w6, w7, sel);
			wq[7] = __byte_perm(w7, w8, sel);
			if (par) {
				q[0] = (uint16_t)w0;
				q[17] = (uint16_t)(w8 >> 16);
			} else {
				wq[8] = w8;
			}
		} else {
			connect(0, p, c);
		}
	}
};

/* ---- index sequences of the AA stroke ----------------------------------------------------------- */
/* round cap after its optional connect (start: stroker.cpp:1491-1510, end: 1947-1968):
 * [fan (F, b+2, b+4) x (n-2)][quads x (n-1)], b = F+2i; an end cap swaps the 2nd and 3rd index of every triangle.
 * Only the one or two cap lanes of a warp run this: plain 16/32-bit stores at two cursors, nothing clever. */
template<class SINK>
__device__ __forceinline__ void roundCapSteps(const SINK& S, uint32_t k, bool isEnd, uint32_t n, uint32_t F)
{
	const uint32_t o2 = isEnd ? 4u : 2u, o3 = isEnd ? 2u : 4u;
	const uint32_t q2 = isEnd ? 3u : 1u, q3 = isEnd ? 1u : 3u, r2 = isEnd ? 2u : 3u, r3 = isEnd ? 3u : 2u;
	uint32_t kf = k, kq = k + 3 * (n - 2);
	for (uint32_t i = 0; i + 1 < n; ++i, kq += 6) {
		const uint32_t b = F + 2 * i;
		if (i + 2 < n) {
			S.I(kf, F);
			S.I(kf + 1, b + o2);
			S.I(kf + 2, b + o3);
			kf += 3;
		}
		S.I(kq, b);
		S.I(kq + 1, b + q2);
		S.I(kq + 2, b + q3);
		S.I(kq + 3, b);
		S.I(kq + 4, b + r2);
		S.I(kq + 5, b + r3);
	}
}

/* the 9 indices of join slice i (stroker.cpp:1673-1685 / 1828-1840): hub = F+1, A = F+2+2i;
 * left inner: (hub, A, A+2) (A, A+1, A+3) (A, A+3, A+2), right inner: 2nd and 3rd index of each swapped */
struct Slice9
{
	uint32_t v[9];
};
__device__ __forceinline__ Slice9 sliceIds(bool left, uint32_t hub, uint32_t A)
{
	const uint32_t A1 = A + 1, A2 = A + 2, A3 = A + 3;
	Slice9 r;
	r.v[0] = hub;
	r.v[1] = left ? A : A2;
	r.v[2] = left ? A2 : A;
	r.v[3] = A;
	r.v[4] = left ? A1 : A3;
	r.v[5] = left ? A3 : A1;
	r.v[6] = A;
	r.v[7] = left ? A3 : A2;
	r.v[8] = left ? A2 : A3;
	return r;
}

/* a bevel / round join's n slices after its optional connect (k = 18 or 0), through any sink */
template<class SINK>
__device__ __forceinline__ void joinSlices(const SINK& S, uint32_t k, bool left, uint32_t n, uint32_t F)
{
	for (uint32_t i = 0; i < n; ++i, k += 9) {
		const Slice9 s = sliceIds(left, F + 1, F + 2 + 2 * i);
		S.tri(k, s.v[0], s.v[1], s.v[2]);
		S.tri2(k + 3, s.v[3], s.v[4], s.v[5], s.v[6], s.v[7], s.v[8]);
	}
}

/* ... and through the staging sink with uint16 indices: straight-line packed words.  k is 0 or 18, so the slices
 * start with the lane's own parity (S.par), and every PAIR of slices (18 indices) keeps it: one aligned word
 * pointer and one PRMT selector serve the whole sequence, all stores at immediate offsets; an odd last slice
 * (9 indices) closes it. */
template<bool AL>
__device__ __forceinline__ void joinSlices(const SinkS<uint16_t, AL>& S, uint32_t k, bool left, uint32_t n, uint32_t F)
{
	const uint32_t odd = S.par;
	const uint32_t sel = odd ? 0x5432u : 0x3210u;
	uint16_t* q = S.idx + k;
	uint32_t* wq = reinterpret_cast<uint32_t*>(q + odd);
	const uint32_t hub = F + 1 + S.base;
	uint32_t A = F + 2 + S.base;
	uint32_t i = 0;
	for (; i + 2 <= n; i += 2, A += 4, q += 18, wq += 9) {
		const Slice9 a = sliceIds(left, hub, A), b = sliceIds(left, hub, A + 2);
		const uint32_t w0 = pk16(a.v[0], a.v[1]), w1 = pk16(a.v[2], a.v[3]), w2 = pk16(a.v[4], a.v[5]), w3 = pk16(a.v[6], a.v[7]),
		               w4 = pk16(a.v[8], b.v[0]), w5 = pk16(b.v[1], b.v[2]), w6 = pk16(b.v[3], b.v[4]), w7 = pk16(b.v[5], b.v[6]),
		               w8 = pk16(b.v[7], b.v[8]);
		wq[0] = __byte_perm(w0, w1, sel);
		wq[1] = __byte_perm(w1, w2, sel);
		wq[2] = __byte_perm(w2, w3, sel);
		wq[3] = __byte_perm(w3, w4, sel);
		wq[4] = __byte_perm(w4, w5, sel);
		wq[5] = __byte_perm(w5, w6, sel);
		wq[6] = __byte_perm(w6, w7, sel);
		wq[7] = __byte_perm(w7, w8, sel);
		if (odd) {
			q[0] = (uint16_t)w0;
			q[17] = (uint16_t)(w8 >> 16);
		} else {
			wq[8] = w8;
		}
	}
	if (i < n) {
		const Slice9 a = sliceIds(left, hub, A);
		const uint32_t w0 = pk16(a.v[0], a.v[1]), w1 = pk16(a.v[2], a.v[3]), w2 = pk16(a.v[4], a.v[5]), w3 = pk16(a.v[6], a.v[7]), w4 = a.v[8];
		wq[0] = __byte_perm(w0, w1, sel);
		wq[1] = __byte_perm(w1, w2, sel);
		wq[2] = __byte_perm(w2, w3, sel);
		wq[3] = __byte_perm(w3, w4, sel);
		q[odd ? 0 : 8] = (uint16_t)(odd ? w0 : w4);
	}
}

/* ---- AA stroke pieces ---------------------------------------------------------------- */
/* everything a butt / square cap needs, gathered from its lane (round caps go through emitStrokeAA) */
struct CapRec
{
	float2 p, d;      /* end point and direction of the adjacent segment        */
	float hsw, hsw_aa, fringe;
	uint32_t color, c0;
	uint32_t n;       /* nHC for round caps                                    */
	uint32_t F;       /* mesh-local id of the cap's first vertex               */
	uint4 prev;       /* cross-section before an end cap                       */
	uint32_t vl, il;  /* block-local offsets of its vertices / indices         */
	uint32_t base;    /* index base                                            */
	uint32_t info;    /* owner thread | cap type << 16 | isEnd << 24          */
};

/* start cap: stroker.cpp:1422-1515; end cap: 1858-1968.  An end cap is the start cap mirrored:
 * the along-segment offsets change sign (negation is exact) and every triangle swaps its 2nd
 * and 3rd index. */
template<int CT = 3, class SINK>
__device__ __forceinline__ void emitCapAA(const CapRec& R, const SINK& S)
{
	const uint32_t capType = CT < 3 ? (uint32_t)CT : ((R.info >> 16) & 0xFF);
	const bool isEnd = (R.info >> 24) != 0;
	const float2 p0 = R.p, d = R.d;
	const float2 l = f2(-d.y, d.x);
	const float hsw = R.hsw, hsw_aa = R.hsw_aa;
	const uint32_t F = R.F;
	uint32_t k = 0;
	if (capType == VGB_CAP_ROUND) {
		const uint32_t n = R.n;
		const float startAngle = vgb_atan2(l.y, l.x);
		for (uint32_t i = 0; i < n; ++i) {
			const float t = (float)i * VGB_PI / (float)(n - 1);
			const float a = startAngle + (isEnd ? -t : t);
			float sa, ca;
			vgb_sincos(a, &sa, &ca);
			S.VP(2 * i, p0.x + ca * hsw, p0.y + sa * hsw, R.color, p0.x + ca * hsw_aa, p0.y + sa * hsw_aa, R.c0);
		}
		if (isEnd) {
			const uint32_t e = F + 2 * (n - 1);
			S.connect0(R.prev, make_uint4(F + 1, F, e, e + 1));
			k = 18;
		}
		roundCapSteps(S, k, isEnd, n, F);
		return;
	}
	const float2 ds = isEnd ? f2(-d.x, -d.y) : d; /* d-terms of the start-cap formulas */
	const float lx = l.x * hsw, ly = l.y * hsw;
	const float lax = l.x * hsw_aa, lay = l.y * hsw_aa;
	if (capType == VGB_CAP_BUTT) {
		const float dax = ds.x * R.fringe, day = ds.y * R.fringe;
		S.VP(0, p0.x + (lax - dax), p0.y + (lay - day), R.c0, p0.x + lx, p0.y + ly, R.color);
		S.VP(2, p0.x - lx, p0.y - ly, R.color, p0.x - (lax + dax), p0.y - (lay + day), R.c0);
	} else {
		const float dx = ds.x * hsw, dy = ds.y * hsw;
		const float dax = ds.x * hsw_aa, day = ds.y * hsw_aa;
		S.VP(0, p0.x + (lax - dax), p0.y + (lay - day), R.c0, p0.x + (lx - dx), p0.y + (ly - dy), R.color);
		S.VP(2, p0.x - (lx + dx), p0.y - (ly + dy), R.color, p0.x - (lax + dax), p0.y - (lay + day), R.c0);
	}
	if (isEnd) {
		S.connect0(R.prev, make_uint4(F, F + 1, F + 2, F + 3));
		k = 18;
	}
	/* stroker.cpp:1443-1444 (start: 0,2,1, 0,3,2) / 1889-1890 (end: 0,1,2, 0,2,3) */
	S.I(k, F);
	S.I(k + 1, isEnd ? F + 1 : F + 2);
	S.I(k + 2, isEnd ? F + 2 : F + 1);
	S.I(k + 3, F);
	S.I(k + 4, isEnd ? F + 2 : F + 3);
	S.I(k + 5, isEnd ? F + 3 : F + 2);
}

/* stroker.cpp:1536-1847. firstLeft: side of the mesh's first join (closing connect only) */
template<int JT = 3, class SINK>
__device__ __forceinline__ void emitJoinAA(const JobCtx& J, const Piece& P, uint32_t F, uint4 prev, bool firstLeft, const SINK& S)
{
	const uint32_t jt = JT < 3 ? (uint32_t)JT : J.join;
	const bool left = (P.flags & PF_LEFT) != 0;
	const float hsw = J.hsw, hsw_aa = J.hsw_aa;
	const float2 p1 = P.p;
	/* right inner: every "+v" of the left case becomes "-v" (negation is exact) */
	const float2 vs = left ? P.v : f2(-P.v.x, -P.v.y);
	const float vax = vs.x * hsw_aa, vay = vs.y * hsw_aa;
	const float vx = vs.x * hsw, vy = vs.y * hsw;
	uint32_t k = 0;
	if (jt == VGB_JOIN_MITER) {
		S.VP(0, p1.x + vax, p1.y + vay, J.c0, p1.x + vx, p1.y + vy, J.color);
		S.VP(2, p1.x - vx, p1.y - vy, J.color, p1.x - vax, p1.y - vay, J.c0);
		if (P.flags & PF_HAS_PREV) {
			S.connect0(prev, joinIn4(left, F));
			k = 18;
		}
	} else {
		const uint32_t n = P.n;
		const float2 d01 = P.d01, d12 = P.d12;
		const float2 s01 = left ? f2(d01.y, -d01.x) : f2(-d01.y, d01.x);
		const float2 s12 = left ? f2(d12.y, -d12.x) : f2(-d12.y, d12.x);
		S.VP(0, p1.x + vax, p1.y + vay, J.c0, p1.x + vx, p1.y + vy, J.color);
		/* first arc vertex */
		{
			float ax = p1.x + s01.x * hsw, ay = p1.y + s01.y * hsw;
			if (jt == VGB_JOIN_BEVEL) {
				const float cosAngle = fabsf(s01.x * s12.x + s01.y * s12.y);
				const float t = cosAngle * J.fringe;
				ax = ax - d01.x * t;
				ay = ay - d01.y * t;
			}
			S.VP(2, ax, ay, J.color, p1.x + s01.x * hsw_aa, p1.y + s01.y * hsw_aa, J.c0);
		}
		/* middle arc vertices */
		for (uint32_t i = 1; i < n; ++i) {
			const float a = P.a01 + (float)i * P.arcDa;
			float sa, ca;
			vgb_sincos(a, &sa, &ca);
			S.VP(2 + 2 * i, p1.x + ca * hsw, p1.y + sa * hsw, J.color, p1.x + ca * hsw_aa, p1.y + sa * hsw_aa, J.c0);
		}
		/* last arc vertex */
		{
			float ax = p1.x + s12.x * hsw, ay = p1.y + s12.y * hsw;
			if (jt == VGB_JOIN_BEVEL) {
				const float cosAngle = fabsf(s01.x * s12.x + s01.y * s12.y);
				const float t = cosAngle * J.fringe;
				ax = ax + d12.x * t;
				ay = ay + d12.y * t;
			}
			S.VP(2 + 2 * n, ax, ay, J.color, p1.x + s12.x * hsw_aa, p1.y + s12.y * hsw_aa, J.c0);
		}
		if (P.flags & PF_HAS_PREV) {
			S.connect0(prev, joinIn4(left, F));
			k = 18;
		}
		joinSlices(S, k, left, n, F);
		k += 9 * n;
	}
	if (P.flags & PF_CLOSING) {
		/* stroker.cpp:1970-1984: previous = this piece, first = piece 0 (its F is 0) */
		S.connect(k, out4FromTail(0, jt, packTail(P), F + P.vcount), joinIn4(firstLeft, 0));
	}
}

/*
 * Joins AND round caps of an AA stroke through one code path (used by the warp-autonomous kernel,
 * where the few cap lanes of a warp would otherwise run a whole separate function while the join
 * lanes wait).  A round cap is emitted as "a join without inner pair whose every arc point is a
 * middle point": both kinds share the sin/cos loop and the paired stores; only the short index
 * patterns differ.  Butt/square caps (4 vertices) and miter joins keep their own small bodies.
 * Values are exactly those of emitJoinAA / emitCapAA (same expressions, same order).
 */
template<int JT = 3, int CT = 3, class SINK>
__device__ __forceinline__ void emitStrokeAA(const JobCtx& J, const Piece& P, uint32_t F, uint4 prev, bool firstLeft, const SINK& S)
{
	const uint32_t jt = JT < 3 ? (uint32_t)JT : J.join;
	const uint32_t ct = CT < 3 ? (uint32_t)CT : J.cap;
	const bool isCap = P.kind != PK_JOIN;
	const bool isEnd = P.kind == PK_CAP_END;
	if ((isCap && ct != VGB_CAP_ROUND) || (!isCap && jt == VGB_JOIN_MITER)) {
		if (isCap) {
			CapRec R;
			R.p = P.p; R.d = P.d01; R.hsw = J.hsw; R.hsw_aa = J.hsw_aa; R.fringe = J.fringe; R.color = J.color; R.c0 = J.c0;
			R.n = P.n; R.F = F; R.prev = prev; R.vl = 0; R.il = 0; R.base = S.base; R.info = (ct << 16) | ((isEnd ? 1u : 0u) << 24);
			emitCapAA<CT>(R, S);
		} else {
			emitJoinAA<JT>(J, P, F, prev, firstLeft, S);
		}
		return;
	}
	const bool left = (P.flags & PF_LEFT) != 0;
	const float hsw = J.hsw, hsw_aa = J.hsw_aa;
	const float2 p1 = P.p;
	const uint32_t n = P.n; /* join: arc segments; cap: nHC */
	const float2 d01 = P.d01, d12 = P.d12;
	const float2 s01 = left ? f2(d01.y, -d01.x) : f2(-d01.y, d01.x);
	const float2 s12 = left ? f2(d12.y, -d12.x) : f2(-d12.y, d12.x);
	if (!isCap) {
		const float2 vs = left ? P.v : f2(-P.v.x, -P.v.y);
		S.VP(0, p1.x + vs.x * hsw_aa, p1.y + vs.y * hsw_aa, J.c0, p1.x + vs.x * hsw, p1.y + vs.y * hsw, J.color);
	}
	/* One loop over the arc points of both kinds, so that the cap lanes' iterations are shared:
	 *   join: k = 0..n at vertex 2+2k; the end points k = 0 and k = n take their direction from the
	 *         segment normals (s01, s12), the middle ones from the angle a01 + k*arcDa;
	 *   cap:  k = 0..nHC-1 at vertex 2k, all from the angle startAngle +- k*pi/(nHC-1).
	 * sin/cos is evaluated for every lane and iteration and simply not used where a normal applies. */
	const uint32_t vb = isCap ? 0u : 2u;
	const uint32_t cnt = isCap ? n : n + 1;
	const float capDen = (float)(n - 1);
	/* (i*pi)/capDen: when capDen is a power of two (nHC = 2, 3, 5, 9 ...) the division is an exact scaling
	 * and multiplying by the exact reciprocal 2^-e gives the same bits; other values keep the division */
	const uint32_t denBits = __float_as_uint(capDen);
	const bool denPow2 = (denBits & 0x007FFFFFu) == 0u;
	const float capRcp = __uint_as_float(0x7F000000u - denBits);
	const float bev = (!isCap && jt == VGB_JOIN_BEVEL) ? fabsf(s01.x * s12.x + s01.y * s12.y) * J.fringe : 0.0f;
	for (uint32_t i = 0; i < cnt; ++i) {
		float t = (float)i * (isCap ? VGB_PI : P.arcDa);
		if (isCap) {
			t = denPow2 ? t * capRcp : t / capDen;
		}
		const float a = P.a01 + (isEnd ? -t : t);
		float dy, dx;
		vgb_sincos(a, &dy, &dx);
		const bool first = !isCap && i == 0, last = !isCap && i == n;
		if (first) {
			dx = s01.x;
			dy = s01.y;
		}
		if (last) {
			dx = s12.x;
			dy = s12.y;
		}
		float ix = p1.x + dx * hsw, iy = p1.y + dy * hsw;
		if (jt == VGB_JOIN_BEVEL) {
			/* stroker.cpp:1614-1617, 1643-1646 (bev is 0 for caps: x - d*0 would still be exact, but skip it) */
			if (first) {
				ix = ix - d01.x * bev;
				iy = iy - d01.y * bev;
			}
			if (last) {
				ix = ix + d12.x * bev;
				iy = iy + d12.y * bev;
			}
		}
		S.VP(vb + 2 * i, ix, iy, J.color, p1.x + dx * hsw_aa, p1.y + dy * hsw_aa, J.c0);
	}
	/* indices.  The connect strip of joins AND end caps is one straight-line block (packed words in the staging
	 * sink); then the joins' slices, straight-line too, while the one or two cap lanes of the warp run their
	 * fan / quad steps on their own. */
	const bool hasPrev = (P.flags & PF_HAS_PREV) != 0;
	if (hasPrev) {
		const uint32_t e = F + 2 * (n - 1);
		S.connect0(prev, isCap ? make_uint4(F + 1, F, e, e + 1) : joinIn4(left, F));
	}
	const uint32_t k = hasPrev ? 18u : 0u;
	if (isCap) {
		roundCapSteps(S, k, isEnd, n, F);
	} else {
		joinSlices(S, k, left, n, F);
		if (P.flags & PF_CLOSING) {
			S.connect(k + 9 * n, out4FromTail(0, jt, packTail(P), F + P.vcount), joinIn4(firstLeft, 0));
		}
	}
}

/* side of join 0 of a closed polyline (needed by the closing connect) */
__device__ __forceinline__ bool firstJoinLeft(const JobCtx& J, const float2* __restrict__ pts)
{
	const float2* v = pts + J.first;
	const float2 d01 = vdir(v[J.N - 1], v[0]);
	const float2 d12 = vdir(v[0], v[J.N == 1 ? 0 : 1]);
	const float2 e = extrusion(d01, d12);
	const float proj = d12.x * (e.x * J.hsw_aa) + d12.y * (e.y * J.hsw_aa);
	return proj >= 0.0f;
}

/* ---- convex AA fill pieces (stroker.cpp:713-806) ---------------------------------------- */
template<class SINK>
__device__ __forceinline__ void emitFillFan(const Piece& P, const SINK& S)
{
	S.tri(0, 0, 2 * P.li + 2, 2 * P.li + 4);
}

template<class SINK>
__device__ __forceinline__ void emitFillVtx(const JobCtx& J, const Piece& P, const float2* __restrict__ pts, const SINK& S)
{
	const float2* v = pts + J.first;
	const float2 a = f2(v[1].x - v[0].x, v[1].y - v[0].y);
	const float2 b = f2(v[2].x - v[0].x, v[2].y - v[0].y);
	const float cross = a.x * b.y - b.x * a.y;
	const float aa = J.fringe * 0.5f * vgb_sign(cross);
	const float2 e = extrusion(P.d01, P.d12);
	const float ex = e.x * aa, ey = e.y * aa;
	S.VP(0, P.p.x + ex, P.p.y + ey, J.color, P.p.x - ex, P.p.y - ey, J.c0);
	const uint32_t f = 2 * P.li;
	if (J.sse && P.li > 0 && P.li + 1 < J.N) {
		/* SIMD order (stroker.cpp:640-690): fan triangle F(i-1) = (0, 2i, 2i+2) right before this vertex's quad */
		S.tri(0, 0, f, f + 2);
		S.tri2(3, f, f + 1, f + 3, f, f + 3, f + 2);
		return;
	}
	if (P.li + 1 < J.N) {
		S.tri2(0, f, f + 1, f + 3, f, f + 3, f + 2);
	} else {
		S.tri2(0, f, f + 1, 1, f, 1, 0);
	}
}
/* ---- plain / thin ribbons: hand-over and emission ------------------------------------------ */
/* ids (left .. right) a plain/thin piece hands to its successor, rebuilt from its tail word */
__device__ __forceinline__ uint4 ribbonOutFromTail(const JobCtx& J, uint32_t tail, uint32_t F)
{
	const uint32_t kind = tail & 7u, n = (tail >> 4) & 0xFFFu;
	const uint32_t Fp = F - (tail >> 16);
	const bool left = (tail & 8u) != 0;
	if (J.op == VGB_OP_STROKE_AA_THIN) {
		if (kind == PK_CAP_START || J.join == VGB_JOIN_MITER) {
			return (kind == PK_CAP_START || left) ? make_uint4(Fp, Fp + 1, Fp + 2, 0) : make_uint4(Fp + 2, Fp + 1, Fp, 0);
		}
		return left ? make_uint4(Fp, Fp + 1, Fp + 3, 0) : make_uint4(Fp + 3, Fp + 1, Fp, 0); /* 2152-2154 / 2232-2234 */
	}
	if (kind == PK_CAP_START) {
		return make_uint4(Fp, J.cap == VGB_CAP_ROUND ? Fp + n - 1 : Fp + 1, 0, 0); /* 1043-1044, 1079-1080 */
	}
	const uint32_t e = (J.join == VGB_JOIN_MITER) ? 1u : n + 1;
	return left ? make_uint4(Fp, Fp + e, 0, 0) : make_uint4(Fp + e, Fp, 0, 0); /* 1130-1131, 1195-1196, 1228-1229, 1289-1290 */
}

template<class SINK>
__device__ __forceinline__ void emitRibbonCap(const JobCtx& J, const Piece& P, uint32_t F, uint4 prev, const SINK& S)
{
	const bool end = P.kind == PK_CAP_END;
	const float2 p0 = P.p, d = P.d01;
	const float2 ds = end ? f2(-d.x, -d.y) : d; /* an end cap mirrors the along-segment offsets */
	const float2 l = f2(-d.y, d.x);
	if (J.op == VGB_OP_STROKE_AA_THIN) {
		/* 2021-2051, 2247-2291 */
		const float h = J.hsw_aa;
		const float lax = l.x * h, lay = l.y * h;
		if (J.cap == VGB_CAP_BUTT) {
			S.V(0, p0.x + lax, p0.y + lay, J.c0);
			S.V(1, p0.x, p0.y, J.color);
			S.V(2, p0.x - lax, p0.y - lay, J.c0);
		} else {
			const float dax = ds.x * h, day = ds.y * h;
			S.V(0, p0.x + (lax - dax), p0.y + (lay - day), J.c0);
			S.V(1, p0.x, p0.y, J.color);
			S.V(2, p0.x - (lax + dax), p0.y - (lay + day), J.c0);
		}
		if (end) {
			S.template bridge<3>(0, prev, make_uint4(F, F + 1, F + 2, 0));
		}
		return;
	}
	const float hsw = J.hsw;
	uint32_t k = 0;
	if (J.cap == VGB_CAP_ROUND) {
		/* 1059-1080, 1342-1371 */
		const uint32_t n = P.n;
		const float startAngle = vgb_atan2(l.y, l.x);
		for (uint32_t i = 0; i < n; ++i) {
			const float t = (float)i * VGB_PI / (float)(n - 1);
			const float a = startAngle + (end ? -t : t);
			S.V(i, p0.x + vgb_cos(a) * hsw, p0.y + vgb_sin(a) * hsw, J.color);
		}
		if (end) {
			S.template bridge<2>(0, prev, make_uint4(F, F + n - 1, 0, 0));
			k = 6;
		}
		for (uint32_t i = 0; i + 2 < n; ++i, k += 3) {
			if (end) S.tri(k, F, F + i + 2, F + i + 1);
			else S.tri(k, F, F + i + 1, F + i + 2);
		}
		return;
	}
	const float lx = l.x * hsw, ly = l.y * hsw;
	if (J.cap == VGB_CAP_BUTT) {
		S.V(0, p0.x + lx, p0.y + ly, J.color);
		S.V(1, p0.x - lx, p0.y - ly, J.color);
	} else {
		const float dx = ds.x * hsw, dy = ds.y * hsw;
		S.V(0, p0.x + (lx - dx), p0.y + (ly - dy), J.color);
		S.V(1, p0.x - (lx + dx), p0.y - (ly + dy), J.color);
	}
	if (end) {
		S.template bridge<2>(0, prev, make_uint4(F, F + 1, 0, 0));
	}
}

template<class SINK>
__device__ __forceinline__ void emitRibbonJoin(const JobCtx& J, const Piece& P, uint32_t F, uint4 prev, bool firstLeft, const SINK& S)
{
	const bool left = (P.flags & PF_LEFT) != 0;
	const bool thin = J.op == VGB_OP_STROKE_AA_THIN;
	const float2 p1 = P.p;
	const float2 vs = left ? P.v : f2(-P.v.x, -P.v.y);
	const float2 d01 = P.d01, d12 = P.d12;
	const float2 s01 = left ? f2(d01.y, -d01.x) : f2(-d01.y, d01.x);
	const float2 s12 = left ? f2(d12.y, -d12.x) : f2(-d12.y, d12.x);
	uint32_t k = 0;
	uint4 out;
	if (thin) {
		const float h = J.hsw_aa;
		const float vax = vs.x * h, vay = vs.y * h;
		const uint4 in = left ? make_uint4(F, F + 1, F + 2, 0) : make_uint4(F + 2, F + 1, F, 0);
		S.V(0, p1.x + vax, p1.y + vay, J.c0);
		S.V(1, p1.x, p1.y, J.color);
		if (J.join == VGB_JOIN_MITER) {
			S.V(2, p1.x - vax, p1.y - vay, J.c0);
			out = in;
		} else {
			S.V(2, p1.x + s01.x * h, p1.y + s01.y * h, J.c0);
			S.V(3, p1.x + s12.x * h, p1.y + s12.y * h, J.c0);
			out = left ? make_uint4(F, F + 1, F + 3, 0) : make_uint4(F + 3, F + 1, F, 0);
		}
		if (P.flags & PF_HAS_PREV) {
			S.template bridge<3>(0, prev, in);
			k = 12;
		}
		if (J.join != VGB_JOIN_MITER) {
			if (left) S.tri(k, F + 1, F + 2, F + 3);
			else S.tri(k, F + 1, F + 3, F + 2);
			k += 3;
		}
		if (P.flags & PF_CLOSING) {
			S.template bridge<3>(k, out, firstLeft ? make_uint4(0, 1, 2, 0) : make_uint4(2, 1, 0, 0));
		}
		return;
	}
	const float hsw = J.hsw;
	const float vx = vs.x * hsw, vy = vs.y * hsw;
	const uint4 in = left ? make_uint4(F, F + 1, 0, 0) : make_uint4(F + 1, F, 0, 0);
	S.V(0, p1.x + vx, p1.y + vy, J.color);
	if (J.join == VGB_JOIN_MITER) {
		S.V(1, p1.x - vx, p1.y - vy, J.color);
		out = in;
		if (P.flags & PF_HAS_PREV) {
			S.template bridge<2>(0, prev, in);
			k = 6;
		}
	} else {
		const uint32_t n = P.n;
		S.V(1, p1.x + s01.x * hsw, p1.y + s01.y * hsw, J.color);
		for (uint32_t i = 1; i < n; ++i) {
			const float a = P.a01 + (float)i * P.arcDa;
			S.V(1 + i, p1.x + hsw * vgb_cos(a), p1.y + hsw * vgb_sin(a), J.color);
		}
		S.V(1 + n, p1.x + s12.x * hsw, p1.y + s12.y * hsw, J.color);
		if (P.flags & PF_HAS_PREV) {
			S.template bridge<2>(0, prev, in);
			k = 6;
		}
		for (uint32_t i = 0; i < n; ++i, k += 3) {
			if (left) S.tri(k, F, F + i + 1, F + i + 2);
			else S.tri(k, F, F + i + 2, F + i + 1);
		}
		out = left ? make_uint4(F, F + n + 1, 0, 0) : make_uint4(F + n + 1, F, 0, 0);
	}
	if (P.flags & PF_CLOSING) {
		S.template bridge<2>(k, out, firstLeft ? make_uint4(0, 1, 0, 0) : make_uint4(1, 0, 0, 0));
	}
}

} // namespace vgb

#endif
