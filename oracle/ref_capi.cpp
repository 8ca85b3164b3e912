/*
 * ref_capi.cpp -- C wrapper around the UNMODIFIED reference hot path (vg::pathXXX /
 * vg::strokerXXX, compiled from /root/reference/src/{path,stroker,vg_util}.cpp by
 * oracle/Makefile against oracle/bx_shim).  TEST INFRASTRUCTURE ONLY: loaded by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
 * product path.  No reference source is copied here; this file only CALLS the reference API
 * (include/vg/path.h:19-38, include/vg/stroker.h:11-70) the way src/vg.cpp does
 * (ctxBeginPath 2969-2981, ctxStrokePathColor 3401-3492, ctxFillPathColor 3061-3179,
 * transformPath 4957-4975).  The bx headers it is compiled against are this repo's shim (bx is neither
 * vendored nor pinned by the reference): with respect to bx's math, "parity unpinned" (DESIGN.md section 2).
 */
#include <vg/path.h>
#include <vg/stroker.h>
#include <bx/allocator.h>
#include "vg_util.h" /* /root/reference/src, via -I */
#include "../include/vgb.h"
#include <string.h>
#include <vector>

namespace
{
bx::DefaultAllocator s_alloc;

struct Ref
{
	vg::Path* path;
	vg::Stroker* stroker;
	std::vector<float> xf;
	Ref() : path(vg::createPath(&s_alloc)), stroker(vg::createStroker(&s_alloc)) {}
	~Ref() { vg::destroyStroker(stroker); vg::destroyPath(path); }
};

Ref& ref()
{
	static thread_local Ref r;
	return r;
}

/* replay one path's command stream through vg::pathXXX; returns floats consumed or -1 */
int replay(vg::Path* p, const uint8_t* cmd, uint32_t ncmd, const float* d)
{
	const float* d0 = d;
	for (uint32_t i = 0; i < ncmd; ++i) {
		switch (cmd[i]) {
		case VGB_CMD_MOVE_TO: vg::pathMoveTo(p, d[0], d[1]); d += 2; break;
		case VGB_CMD_LINE_TO: vg::pathLineTo(p, d[0], d[1]); d += 2; break;
		case VGB_CMD_CUBIC_TO: vg::pathCubicTo(p, d[0], d[1], d[2], d[3], d[4], d[5]); d += 6; break;
		case VGB_CMD_QUAD_TO: vg::pathQuadraticTo(p, d[0], d[1], d[2], d[3]); d += 4; break;
		case VGB_CMD_ARC_TO: vg::pathArcTo(p, d[0], d[1], d[2], d[3], d[4]); d += 5; break;
		case VGB_CMD_RECT: vg::pathRect(p, d[0], d[1], d[2], d[3]); d += 4; break;
		case VGB_CMD_ROUNDED_RECT: vg::pathRoundedRect(p, d[0], d[1], d[2], d[3], d[4]); d += 5; break;
		case VGB_CMD_ROUNDED_RECT_VARYING: vg::pathRoundedRectVarying(p, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]); d += 8; break;
		case VGB_CMD_CIRCLE: vg::pathCircle(p, d[0], d[1], d[2]); d += 3; break;
		case VGB_CMD_ELLIPSE: vg::pathEllipse(p, d[0], d[1], d[2], d[3]); d += 4; break;
		case VGB_CMD_ARC: vg::pathArc(p, d[0], d[1], d[2], d[3], d[4], d[5] != 0.0f ? vg::Winding::CW : vg::Winding::CCW); d += 6; break;
		case VGB_CMD_POLYLINE: {
			uint32_t n;
			memcpy(&n, d, 4);
			vg::pathPolyline(p, d + 1, n);
			d += 1 + 2 * (size_t)n;
		} break;
		case VGB_CMD_CLOSE: vg::pathClose(p); break;
		default: return -1;
		}
	}
	return (int)(d - d0);
}

void runMesh(vg::Stroker* s, vg::Mesh* m, uint32_t op, const float* vtx, uint32_t n, bool closed, uint32_t color, float width, uint32_t cap, uint32_t join)
{
	memset(m, 0, sizeof(*m));
	switch (op) {
	case VGB_OP_STROKE_AA: vg::strokerPolylineStrokeAA(s, m, vtx, n, closed, color, width, (vg::LineCap::Enum)cap, (vg::LineJoin::Enum)join); break;
	case VGB_OP_STROKE: vg::strokerPolylineStroke(s, m, vtx, n, closed, width, (vg::LineCap::Enum)cap, (vg::LineJoin::Enum)join); break;
	case VGB_OP_STROKE_AA_THIN: vg::strokerPolylineStrokeAAThin(s, m, vtx, n, closed, color, (vg::LineCap::Enum)cap, (vg::LineJoin::Enum)join); break;
	case VGB_OP_CONVEX_FILL_AA: vg::strokerConvexFillAA(s, m, vtx, n, color); break;
	case VGB_OP_CONVEX_FILL: vg::strokerConvexFill(s, m, vtx, n); break;
	default: break;
	}
}

inline uint32_t minVerts(uint32_t op)
{
	return (op == VGB_OP_CONVEX_FILL_AA || op == VGB_OP_CONVEX_FILL) ? 3u : 2u; /* vg.cpp:3101, 3450 */
}
}

extern "C" {

/* which build is this: bit0 = VG_CONFIG_ENABLE_SIMD, bit1 = libm-backed bx shim */
uint32_t vgref_variant(void)
{
	uint32_t v = 0;
#if VG_CONFIG_ENABLE_SIMD
	v |= 1;
#endif
#if defined(BX_SHIM_LIBM) && BX_SHIM_LIBM
	v |= 2;
#endif
#if defined(VGREF_DROPIN) && VGREF_DROPIN
	v |= 4; /* this file linked against libvgb200.so's vg:: symbols instead of path.cpp + stroker.cpp */
#endif
	return v;
}

/* Calls EVERY symbol of include/vg/path.h:19-38 and include/vg/stroker.h:11-85 once, on a fresh Path /
 * Stroker pair created through a bx allocator (the link-level drop-in test executes this in the build
 * that is linked against libvgb200.so).  out[0..7] = points, sub-paths, and V/I of a plain stroke, an AA
 * stroke, a thin stroke; out[8..11] = V/I of a fan fill and an AA fill; out[12] = concave result bits
 * (Begin | End << 1 | EndAA << 2), out[13..14] = V/I of the concave AA mesh. */
int vgref_touch_all(uint32_t* out)
{
	vg::Path* p = vg::createPath(&s_alloc);
	vg::Stroker* s = vg::createStroker(&s_alloc);
	vg::pathReset(p, 1.0f, 0.25f);
	vg::strokerReset(s, 1.0f, 0.25f, 1.0f);
	vg::pathMoveTo(p, 10.0f, 10.0f);
	vg::pathLineTo(p, 60.0f, 15.0f);
	vg::pathCubicTo(p, 80.0f, 0.0f, 120.0f, 90.0f, 150.0f, 40.0f);
	vg::pathQuadraticTo(p, 170.0f, 10.0f, 200.0f, 60.0f);
	vg::pathArcTo(p, 230.0f, 90.0f, 200.0f, 140.0f, 12.0f);
	vg::pathClose(p);
	vg::pathRect(p, 300.0f, 10.0f, 40.0f, 30.0f);
	vg::pathRoundedRect(p, 300.0f, 60.0f, 80.0f, 50.0f, 9.0f);
	vg::pathRoundedRectVarying(p, 300.0f, 130.0f, 80.0f, 50.0f, 4.0f, 8.0f, 12.0f, 0.0f);
	vg::pathCircle(p, 450.0f, 40.0f, 25.0f);
	vg::pathEllipse(p, 450.0f, 120.0f, 35.0f, 20.0f);
	vg::pathMoveTo(p, 520.0f, 20.0f);
	vg::pathArc(p, 560.0f, 60.0f, 30.0f, 0.3f, 2.8f, vg::Winding::CW);
	const float poly[8] = {600.0f, 10.0f, 640.0f, 30.0f, 620.0f, 70.0f, 590.0f, 50.0f};
	vg::pathMoveTo(p, 600.0f, 10.0f);
	vg::pathPolyline(p, poly, 4);
	const float* v = vg::pathGetVertices(p);
	const uint32_t nv = vg::pathGetNumVertices(p);
	const vg::SubPath* sp = vg::pathGetSubPaths(p);
	const uint32_t ns = vg::pathGetNumSubPaths(p);
	out[0] = nv;
	out[1] = ns;
	vg::Mesh m;
	const float* v0 = v + 2 * (size_t)sp[0].m_FirstVertexID;
	const uint32_t n0 = sp[0].m_NumVertices;
	vg::strokerPolylineStroke(s, &m, v0, n0, sp[0].m_IsClosed, 5.0f, vg::LineCap::Round, vg::LineJoin::Round);
	out[2] = m.m_NumVertices; out[3] = m.m_NumIndices;
	vg::strokerPolylineStrokeAA(s, &m, v0, n0, sp[0].m_IsClosed, 0xFF2040FFu, 5.0f, vg::LineCap::Square, vg::LineJoin::Bevel);
	out[4] = m.m_NumVertices; out[5] = m.m_NumIndices;
	vg::strokerPolylineStrokeAAThin(s, &m, v0, n0, sp[0].m_IsClosed, 0xFF2040FFu, vg::LineCap::Butt, vg::LineJoin::Miter);
	out[6] = m.m_NumVertices; out[7] = m.m_NumIndices;
	const float* v2 = v + 2 * (size_t)sp[2].m_FirstVertexID; /* the rounded rect: convex */
	vg::strokerConvexFill(s, &m, v2, sp[2].m_NumVertices);
	out[8] = m.m_NumVertices; out[9] = m.m_NumIndices;
	vg::strokerConvexFillAA(s, &m, v2, sp[2].m_NumVertices, 0xFF00FF00u);
	out[10] = m.m_NumVertices; out[11] = m.m_NumIndices;
	uint32_t bits = vg::strokerConcaveFillBegin(s) ? 1u : 0u;
	vg::strokerConcaveFillAddContour(s, v0, n0);
	vg::Mesh c;
	memset(&c, 0, sizeof(c));
	bits |= vg::strokerConcaveFillEnd(s, &c, vg::FillRule::NonZero) ? 2u : 0u;
	vg::strokerConcaveFillBegin(s);
	vg::strokerConcaveFillAddContour(s, v0, n0);
	memset(&c, 0, sizeof(c));
	bits |= vg::strokerConcaveFillEndAA(s, &c, 0xFFFFFFFFu, vg::FillRule::EvenOdd) ? 4u : 0u;
	out[12] = bits;
	out[13] = c.m_NumVertices; out[14] = c.m_NumIndices;
	vg::destroyStroker(s);
	vg::destroyPath(p);
	return 0;
}

/* flatten ONE path. returns 0, or -1 on bad command / -5 if a capacity is too small */
int vgref_flatten(const uint8_t* cmd, uint32_t ncmd, const float* data, float scale, float tol,
                  float* pts, uint32_t ptsCap, vgb_subpath* sp, uint32_t spCap, uint32_t* numPts, uint32_t* numSp)
{
	Ref& r = ref();
	vg::pathReset(r.path, scale, tol);
	if (replay(r.path, cmd, ncmd, data) < 0) return -1;
	const uint32_t nv = vg::pathGetNumVertices(r.path);
	const uint32_t ns = vg::pathGetNumSubPaths(r.path);
	*numPts = nv;
	*numSp = ns;
	if (nv > ptsCap || ns > spCap) return -5;
	if (nv) memcpy(pts, vg::pathGetVertices(r.path), sizeof(float) * 2 * nv);
	const vg::SubPath* s = vg::pathGetSubPaths(r.path);
	for (uint32_t i = 0; i < ns; ++i) {
		sp[i].first_vertex = s[i].m_FirstVertexID;
		sp[i].num_vertices = s[i].m_NumVertices;
		sp[i].is_closed = s[i].m_IsClosed ? 1 : 0;
	}
	return 0;
}

/* ONE strokerXXX call on a caller polyline */
int vgref_mesh(uint32_t op, const float* vtx, uint32_t n, int closed, uint32_t color, float width, uint32_t cap, uint32_t join,
               float scale, float tol, float fringe,
               float* pos, uint32_t* col, uint16_t* idx, uint32_t vCap, uint32_t iCap, uint32_t* nv, uint32_t* ni, int* hasColor)
{
	Ref& r = ref();
	vg::strokerReset(r.stroker, scale, tol, fringe);
	vg::Mesh m;
	runMesh(r.stroker, &m, op, vtx, n, closed != 0, color, width, cap, join);
	*nv = m.m_NumVertices;
	*ni = m.m_NumIndices;
	*hasColor = m.m_ColorBuffer != nullptr;
	if (m.m_NumVertices > vCap || m.m_NumIndices > iCap) return -5;
	if (m.m_NumVertices) memcpy(pos, m.m_PosBuffer, sizeof(float) * 2 * m.m_NumVertices);
	if (m.m_ColorBuffer && col) memcpy(col, m.m_ColorBuffer, sizeof(uint32_t) * m.m_NumVertices);
	if (m.m_NumIndices) memcpy(idx, m.m_IndexBuffer, sizeof(uint16_t) * m.m_NumIndices);
	return 0;
}

/*
 * Whole batch, paths [pathBegin, pathEnd), the way vg.cpp drives the objects: per path
 * pathReset+strokerReset, build, transform, then for every draw of that path and every
 * sub-path with enough vertices one strokerXXX call.  Draws must be sorted by path
 * (drawBegin[p] .. drawBegin[p+1]).  With store=0 nothing is copied out (timing mode); totals
 * and an order-sensitive checksum are always produced.
 * Outputs (store=1): pos/col/idx concatenated, per-mesh exclusive offsets (meshCap+1 entries).
 */
int vgref_run_batch(uint32_t pathBegin, uint32_t pathEnd,
                    const uint32_t* pathCmdOff, const uint32_t* pathDataOff, const uint8_t* cmd, const float* data,
                    const float* pathScale, const float* pathTol, const float* pathXform, float fringe,
                    const uint32_t* drawBegin, const uint32_t* drawOp, const uint32_t* drawColor, const float* drawWidth,
                    const uint8_t* drawCap, const uint8_t* drawJoin,
                    int store, float* pos, uint32_t* col, uint16_t* idx, uint64_t vCap, uint64_t iCap,
                    uint64_t* meshVtxOff, uint64_t* meshIdxOff, uint32_t meshCap,
                    uint64_t* totals /* [0]=verts [1]=indices [2]=points [3]=meshes [4]=checksum [5]=subpaths */)
{
	Ref& r = ref();
	uint64_t nv = 0, ni = 0, npts = 0, nmesh = 0, nsub = 0, sum = 0;
	for (uint32_t p = pathBegin; p < pathEnd; ++p) {
		const float scale = pathScale[p], tol = pathTol[p];
		vg::pathReset(r.path, scale, tol);
		vg::strokerReset(r.stroker, scale, tol, fringe);
		if (replay(r.path, cmd + pathCmdOff[p], pathCmdOff[p + 1] - pathCmdOff[p], data + pathDataOff[p]) < 0) return -1;
		const uint32_t n = vg::pathGetNumVertices(r.path);
		const float* v = vg::pathGetVertices(r.path);
		if (pathXform) {
			r.xf.resize(2 * (size_t)n + 4);
			vgutil::batchTransformPositions(v, n, r.xf.data(), pathXform + 6 * (size_t)p);
			v = r.xf.data();
		}
		npts += n;
		const uint32_t ns = vg::pathGetNumSubPaths(r.path);
		const vg::SubPath* sp = vg::pathGetSubPaths(r.path);
		nsub += ns;
		for (uint32_t d = drawBegin[p]; d < drawBegin[p + 1]; ++d) {
			for (uint32_t s = 0; s < ns; ++s) {
				vg::Mesh m;
				memset(&m, 0, sizeof(m));
				if (sp[s].m_NumVertices >= minVerts(drawOp[d])) {
					runMesh(r.stroker, &m, drawOp[d], v + 2 * (size_t)sp[s].m_FirstVertexID, sp[s].m_NumVertices, sp[s].m_IsClosed,
					        drawColor[d], drawWidth[d], drawCap[d], drawJoin[d]);
				}
				if (store) {
					if (nmesh >= meshCap || nv + m.m_NumVertices > vCap || ni + m.m_NumIndices > iCap) return -5;
					meshVtxOff[nmesh] = nv;
					meshIdxOff[nmesh] = ni;
					if (m.m_NumVertices) {
						memcpy(pos + 2 * nv, m.m_PosBuffer, sizeof(float) * 2 * m.m_NumVertices);
						if (m.m_ColorBuffer) memcpy(col + nv, m.m_ColorBuffer, sizeof(uint32_t) * m.m_NumVertices);
						else for (uint32_t k = 0; k < m.m_NumVertices; ++k) col[nv + k] = drawColor[d];
					}
					if (m.m_NumIndices) memcpy(idx + ni, m.m_IndexBuffer, sizeof(uint16_t) * m.m_NumIndices);
				}
				if (m.m_NumIndices) sum = sum * 1099511628211ull + m.m_IndexBuffer[m.m_NumIndices - 1] + m.m_NumVertices;
				nv += m.m_NumVertices;
				ni += m.m_NumIndices;
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

/* ---- final layout ---------------------------------------------------------
 * src/vg.cpp itself cannot be compiled here (bgfx, bx handle allocators), so the BOOK-KEEPING of
 * its merge step is restated below with the reference statement cited on each line; the two data
 * transforms of that step are the reference's own compiled functions: vgutil::memset32 for the uv
 * splat (vg.cpp:5223) and vgutil::batchTransformDrawIndices for the index rebase (vg.cpp:5240).
 * Input: the meshes of a batch in submission order (local uint16 indices), as vgref_run_batch
 * stores them.  Empty meshes are skipped like vg.cpp:3101 / 3450 skip them.  One record per draw
 * command the batch touches; a record describes the batch's share of that command. */
/* meshScissor (may be null): per mesh, scissor id | 1u << 31 on the first mesh of a draw recorded after a scissor
 * command (ctxSetScissor / ctxIntersectScissor / ctxResetScissor / ctxPopState raise m_ForceNewDrawCommand,
 * vg.cpp:3946-4033; pinned against the real context in tests/test_vgctx.py) */
int vgref_final_layout_sc(const uint64_t* meshVtxOff, const uint64_t* meshIdxOff, uint32_t numMeshes, const uint16_t* idxLocal,
                          const vgb_final_cfg* cfg, const uint32_t* meshScissor, uint16_t* idxOut, int16_t* uvOut, vgb_draw_cmd* cmds,
                          uint32_t cmdCap, uint32_t* numCmds)
{
	uint32_t vbID = cfg->vb_id, vbCount = cfg->vb_count, ibCount = cfg->ib_count; /* Context state on entry */
	bool forceNew = cfg->cmd_vertices == 0;      /* m_ForceNewDrawCommand: no open command to continue  */
	uint32_t cmdNumVertices = cfg->cmd_vertices; /* m_NumVertices of the command being appended to     */
	vgb_draw_cmd* cur = nullptr;
	uint32_t n = 0;
	for (uint32_t m = 0; m < numMeshes; ++m) {
		const uint64_t vo = meshVtxOff[m] - meshVtxOff[0], io = meshIdxOff[m] - meshIdxOff[0];
		const uint32_t nv = (uint32_t)(meshVtxOff[m + 1] - meshVtxOff[m]), ni = (uint32_t)(meshIdxOff[m + 1] - meshIdxOff[m]);
		if (meshScissor && (meshScissor[m] >> 31)) forceNew = true; /* stays up until allocDrawCommand runs, vg.cpp:5402 */
		if (nv == 0) continue;
		if (nv >= cfg->max_vb_vertices) return -5;  /* VG_CHECK, vg.cpp:5323 */
		if (vbCount + nv > cfg->max_vb_vertices) {  /* allocVertices: the buffer cannot hold it, vg.cpp:5327-5335 */
			++vbID;
			vbCount = 0;
			forceNew = true;
		}
		const uint32_t firstVertexID = vbCount;     /* vg.cpp:5339-5340 */
		vbCount += nv;
		const uint32_t firstIndexID = ibCount;      /* allocIndices, vg.cpp:5353-5355 */
		ibCount += ni;
		if (forceNew || !cur) {                     /* allocDrawCommand, vg.cpp:5368-5402 (type and handle never change here) */
			if (n == cmdCap) return -5;
			cur = &cmds[n++];
			memset(cur, 0, sizeof(*cur));
			cur->vb_id = vbID;
			cur->first_vertex = firstVertexID;
			cur->first_index = firstIndexID;
			cur->first_mesh = m;
			cur->scissor_id = meshScissor ? (meshScissor[m] & 0x7fffffffu) : 0u; /* m_ScissorRect, vg.cpp:5392-5395 */
			cur->stream_vertex = vo;
			cur->stream_index = io;
			if (forceNew) cmdNumVertices = 0;       /* vg.cpp:5391 */
			forceNew = false;                       /* vg.cpp:5402 */
		}
		/* createDrawCommand_VertexColor, vg.cpp:5214-5243 (pos and colour are plain copies) */
		vgutil::memset32(uvOut + 2 * vo, nv, cfg->uv);
		vgutil::batchTransformDrawIndices(idxLocal + io, ni, idxOut + io, (uint16_t)cmdNumVertices);
		cmdNumVertices += nv;
		cur->num_vertices += nv;
		cur->num_indices += ni;
		cur->num_meshes = m + 1 - cur->first_mesh;
	}
	*numCmds = n;
	return 0;
}

int vgref_final_layout(const uint64_t* meshVtxOff, const uint64_t* meshIdxOff, uint32_t numMeshes, const uint16_t* idxLocal,
                       const vgb_final_cfg* cfg, uint16_t* idxOut, int16_t* uvOut, vgb_draw_cmd* cmds, uint32_t cmdCap, uint32_t* numCmds)
{
	return vgref_final_layout_sc(meshVtxOff, meshIdxOff, numMeshes, idxLocal, cfg, nullptr, idxOut, uvOut, cmds, cmdCap, numCmds);
}

/* vgutil::batchTransformPositions (vg_util.cpp:136-227; this build: the scalar branch), the
 * function submitCachedMesh (vg.cpp:6137-6160) and transformPath (vg.cpp:4957-4975) call */
void vgref_transform_positions(const float* pos, uint32_t n, const float* mtx, float* out)
{
	vgutil::batchTransformPositions(pos, n, out, mtx);
}

} /* extern "C" */
