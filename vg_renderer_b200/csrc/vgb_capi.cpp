/*
 * vgb_capi.cpp -- host side of libvgb200: the batch recorder (mirrors how src/vg.cpp drives
 * Path + Stroker: ctxBeginPath 2969-2981, ctxMoveTo.. 2983-3059, ctxStrokePathColor 3401-3492,
 * ctxFillPathColor 3061-3179) and the per-object API (1:1 with include/vg/path.h and
 * include/vg/stroker.h).  Host code only records and moves bytes; every vertex, colour and
 * index is computed by the CUDA pipeline (vgb_device.cu).  No CPU fallback exists: without a
 * GPU every entry point that would compute fails with VGB_ERR_NOGPU.
 */
#include "vgb.h"
#include "vgb_internal.h"
#include "vgb_math.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

using namespace vgb;

static thread_local std::string t_err;

static int fail(int code, const char* msg)
{
	t_err = msg;
	return code;
}

#define NEED_PIPE(b) \
	do { \
		if (!(b)) return fail(VGB_ERR_INVALID, "null batch"); \
		if (!(b)->pipe) return fail(VGB_ERR_NOGPU, "record-only batch (device -1): no GPU pipeline, and there is no CPU fallback"); \
	} while (0)

static inline bool finite_all(const float* p, size_t n)
{
	/* a NaN/Inf coordinate makes the reference's pathCubicTo loop forever */
	uint32_t bad = 0;
	for (size_t i = 0; i < n; ++i) {
		uint32_t u;
		memcpy(&u, p + i, 4);
		bad |= ((u & 0x7F800000u) == 0x7F800000u);
	}
	return bad == 0;
}

struct vgb_batch
{
	Pipeline* pipe = nullptr;
	int device = 0;
	float fringe = 1.0f;
	/* command stream */
	std::vector<uint8_t> cmd;
	std::vector<float> data;
	std::vector<SubRec> subs;
	std::vector<uint32_t> slotFirst; /* built at upload: point-slot prefix per sub-path */
	std::vector<float> pathScale, pathTol, pathXform;
	bool anyXform = false;
	std::vector<DrawRec> draws;
	std::vector<uint32_t> jobSub, jobDraw;
	/* scissor state (vgb_batch_set_scissor): table of distinct uint16 rects, id of the current one, pending force flag */
	std::vector<uint16_t> scissors = std::vector<uint16_t>(4, 0);
	std::vector<uint32_t> drawScissor;
	std::unordered_map<uint64_t, uint32_t> scissorIds;
	uint32_t curScissor = 0, scissorBreaks = 0;
	bool forceNewCmd = false, anyScissor = false;
	/* current path */
	bool inPath = false;
	bool subOpen = false; /* the current path has a current sub-path */
	bool drawn = false;   /* a draw was issued for the current path: no more commands */
	uint32_t curPath = 0, curFirstSub = 0;
	/* raw polylines */
	bool rawMode = false;
	std::vector<float> rawPts;
	std::vector<uint32_t> rawFirst, rawNum;
	std::vector<uint8_t> rawClosed;
	bool dirty = true;
	/* borrowed bulk input (vgb_batch_set_cubic_strokes) */
	const float* bulkP8 = nullptr;
	uint32_t bulkN = 0;
	float bulkScale = 1.0f, bulkTol = 0.25f;

	void clear(float fr)
	{
		bulkP8 = nullptr;
		bulkN = 0;
		fringe = fr;
		cmd.clear(); data.clear(); subs.clear();
		pathScale.clear(); pathTol.clear(); pathXform.clear();
		anyXform = false;
		draws.clear(); jobSub.clear(); jobDraw.clear();
		scissors.assign(4, 0); drawScissor.clear(); scissorIds.clear();
		curScissor = 0; scissorBreaks = 0; forceNewCmd = false; anyScissor = false;
		inPath = subOpen = drawn = false;
		rawMode = false;
		rawPts.clear(); rawFirst.clear(); rawNum.clear(); rawClosed.clear();
		dirty = true;
	}

	/* per-path transforms are materialised lazily: identity rows are only written once some path
	 * actually has a transform (or at upload, if any does) */
	void xformIdentityUpTo(size_t numPaths)
	{
		if (!anyXform) return;
		static const float ident[6] = {1, 0, 0, 1, 0, 0};
		while (pathXform.size() < 6 * numPaths) pathXform.insert(pathXform.end(), ident, ident + 6);
	}

	void closeSub()
	{
		if (!subs.empty() && subs.back().pathId == curPath && inPath) {
			subs.back().cmdEnd = (uint32_t)cmd.size();
		}
	}

	void newSub()
	{
		closeSub();
		SubRec r;
		r.cmdBegin = (uint32_t)cmd.size();
		r.cmdEnd = r.cmdBegin;
		r.dataBegin = (uint32_t)data.size();
		r.pathId = curPath;
		subs.push_back(r);
		subOpen = true;
	}

	int push(uint8_t c, const float* d, uint32_t n, bool startsSub)
	{
		if (rawMode) return fail(VGB_ERR_STATE, "path commands after vgb_batch_set_polylines");
		if (bulkP8) return fail(VGB_ERR_STATE, "path commands after vgb_batch_set_cubic_strokes (reset first)");
		if (!inPath) return fail(VGB_ERR_INVALID, "path command outside begin_path");
		if (drawn) return fail(VGB_ERR_INVALID, "path command after a draw: begin a new path");
		if (!finite_all(d, n)) return fail(VGB_ERR_NONFINITE, "non-finite coordinate");
		if (startsSub) {
			newSub();
		} else if (!subOpen) {
			return fail(VGB_ERR_INVALID, "moveTo() must precede this command (path.cpp:82)");
		}
		cmd.push_back(c);
		data.insert(data.end(), d, d + n);
		subs.back().cmdEnd = (uint32_t)cmd.size();
		dirty = true;
		return 0;
	}
};

static inline bool degenerateRect(float w, float h)
{
	return fabsf(w) < VGB_EPSILON || fabsf(h) < VGB_EPSILON; /* path.cpp:277 */
}

extern "C" {

const char* vgb_last_error(void) { return t_err.c_str(); }
int vgb_device_count(void) { return deviceCount(); }
const char* vgb_version(void) { return "vgb200 0.1 (sm_100a)"; }

vgb_batch* vgb_batch_create(int device)
{
	std::string err;
	Pipeline* p = nullptr;
	if (device != -1) {
		p = Pipeline::create(device, &err);
		if (!p) {
			t_err = err;
			return nullptr;
		}
	}
	vgb_batch* b = new (std::nothrow) vgb_batch();
	if (!b) {
		delete p;
		return nullptr;
	}
	b->pipe = p;
	b->device = device;
	return b;
}

void vgb_batch_destroy(vgb_batch* b)
{
	if (!b) return;
	delete b->pipe;
	delete b;
}

int vgb_batch_reset(vgb_batch* b, float fringe)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	b->clear(fringe);
	return 0;
}

int vgb_batch_begin_path(vgb_batch* b, float scale, float tol, const float* xform6)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	if (b->rawMode) return fail(VGB_ERR_STATE, "begin_path after vgb_batch_set_polylines");
	if (b->bulkP8) return fail(VGB_ERR_STATE, "begin_path after vgb_batch_set_cubic_strokes (reset first)");
	if (!(scale > 0.0f) || !(tol > 0.0f) || !finite_all(&scale, 1) || !finite_all(&tol, 1)) return fail(VGB_ERR_INVALID, "scale and tolerance must be positive and finite");
	if (xform6 && !finite_all(xform6, 6)) return fail(VGB_ERR_NONFINITE, "non-finite transform");
	b->closeSub();
	b->curPath = (uint32_t)b->pathScale.size();
	b->curFirstSub = (uint32_t)b->subs.size();
	b->pathScale.push_back(scale);
	b->pathTol.push_back(tol);
	if (xform6) {
		b->anyXform = true;
		b->xformIdentityUpTo(b->pathScale.size() - 1); /* rows of the earlier, untransformed paths */
		b->pathXform.insert(b->pathXform.end(), xform6, xform6 + 6);
	} else {
		b->xformIdentityUpTo(b->pathScale.size());
	}
	b->inPath = true;
	b->subOpen = false;
	b->drawn = false;
	b->dirty = true;
	return 0;
}

extern "C" void vgb_set_error_message(const char* msg) { t_err = msg ? msg : ""; } /* for the other host translation units */

float vgb_batch_fringe(vgb_batch* b) { return b ? b->fringe : 0.0f; }

int vgb_batch_set_path_xform(vgb_batch* b, const float* xform6)
{
	if (!b || !b->inPath || !xform6) return fail(VGB_ERR_INVALID, "set_path_xform outside a path");
	if (!finite_all(xform6, 6)) return fail(VGB_ERR_NONFINITE, "non-finite transform");
	if (b->drawn) return 0; /* vg.cpp transforms a path once, at its first fill/stroke (transformPath, vg.cpp:4957-4975) */
	b->anyXform = true;
	b->xformIdentityUpTo(b->pathScale.size());
	memcpy(&b->pathXform[6 * (size_t)b->curPath], xform6, sizeof(float) * 6);
	b->dirty = true;
	return 0;
}

int vgb_batch_move_to(vgb_batch* b, float x, float y)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	const float d[2] = {x, y};
	return b->push(VGB_CMD_MOVE_TO, d, 2, true);
}
int vgb_batch_line_to(vgb_batch* b, float x, float y)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	const float d[2] = {x, y};
	return b->push(VGB_CMD_LINE_TO, d, 2, false);
}
int vgb_batch_cubic_to(vgb_batch* b, float c1x, float c1y, float c2x, float c2y, float x, float y)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	const float d[6] = {c1x, c1y, c2x, c2y, x, y};
	return b->push(VGB_CMD_CUBIC_TO, d, 6, false);
}
int vgb_batch_quadratic_to(vgb_batch* b, float cx, float cy, float x, float y)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	const float d[4] = {cx, cy, x, y};
	return b->push(VGB_CMD_QUAD_TO, d, 4, false);
}
int vgb_batch_arc_to(vgb_batch* b, float x1, float y1, float x2, float y2, float r)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	const float d[5] = {x1, y1, x2, y2, r};
	return b->push(VGB_CMD_ARC_TO, d, 5, false);
}
int vgb_batch_rect(vgb_batch* b, float x, float y, float w, float h)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	if (degenerateRect(w, h)) return 0; /* pathRect returns without touching the path */
	const float d[4] = {x, y, w, h};
	return b->push(VGB_CMD_RECT, d, 4, true);
}
int vgb_batch_rounded_rect(vgb_batch* b, float x, float y, float w, float h, float r)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	if (r < 0.1f && degenerateRect(w, h)) return 0; /* path.cpp:290-293 */
	const float d[5] = {x, y, w, h, r};
	return b->push(VGB_CMD_ROUNDED_RECT, d, 5, true);
}
int vgb_batch_rounded_rect_varying(vgb_batch* b, float x, float y, float w, float h, float rtl, float rtr, float rbr, float rbl)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	if (rtl < 0.1f && rbl < 0.1f && rbr < 0.1f && rtr < 0.1f && degenerateRect(w, h)) return 0; /* path.cpp:413-416 */
	const float d[8] = {x, y, w, h, rtl, rtr, rbr, rbl};
	return b->push(VGB_CMD_ROUNDED_RECT_VARYING, d, 8, true);
}
int vgb_batch_circle(vgb_batch* b, float x, float y, float r)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	const float d[3] = {x, y, r};
	return b->push(VGB_CMD_CIRCLE, d, 3, true);
}
int vgb_batch_ellipse(vgb_batch* b, float x, float y, float rx, float ry)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	const float d[4] = {x, y, rx, ry};
	return b->push(VGB_CMD_ELLIPSE, d, 4, true);
}
int vgb_batch_arc(vgb_batch* b, float x, float y, float r, float a0, float a1, uint32_t dir)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	/* pathArc normalises its angles with `while (a > 2pi) a -= 2pi` (path.cpp:635-645): beyond ~1e8 the
	 * subtraction no longer changes a float and the reference never returns; a device thread would spin
	 * the same way.  Angles that large are refused at record time like non-finite coordinates are. */
	if (b && (fabsf(a0) > 1.0e6f || fabsf(a1) > 1.0e6f)) return fail(VGB_ERR_INVALID, "arc angle magnitude above 1e6 rad");
	const float d[6] = {x, y, r, a0, a1, dir == VGB_WINDING_CW ? 1.0f : 0.0f};
	/* pathArc: lineTo when the path has a current sub-path, moveTo otherwise (path.cpp:663-667) */
	return b->push(VGB_CMD_ARC, d, 6, !b->subOpen);
}
int vgb_batch_polyline(vgb_batch* b, const float* coords, uint32_t n)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	if (!coords && n) return fail(VGB_ERR_INVALID, "null coords");
	if (!finite_all(coords, 2 * (size_t)n)) return fail(VGB_ERR_NONFINITE, "non-finite coordinate");
	float hdr;
	memcpy(&hdr, &n, 4);
	/* the count word is not a coordinate: bypass the finite check by pushing in two steps */
	if (b->rawMode) return fail(VGB_ERR_STATE, "path commands after vgb_batch_set_polylines");
	if (!b->inPath || b->drawn) return fail(VGB_ERR_INVALID, "polyline outside an open path");
	if (!b->subOpen) return fail(VGB_ERR_INVALID, "moveTo() must precede polyline() (path.cpp:686)");
	b->cmd.push_back(VGB_CMD_POLYLINE);
	b->data.push_back(hdr);
	b->data.insert(b->data.end(), coords, coords + 2 * (size_t)n);
	b->subs.back().cmdEnd = (uint32_t)b->cmd.size();
	b->dirty = true;
	return 0;
}
int vgb_batch_close(vgb_batch* b)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	return b->push(VGB_CMD_CLOSE, nullptr, 0, false);
}

static int addDraw(vgb_batch* b, uint32_t op, uint32_t color, float width, uint32_t cap, uint32_t join, uint32_t pathId, uint32_t sub0, uint32_t sub1)
{
	if (op > VGB_OP_CONVEX_FILL) return fail(VGB_ERR_INVALID, "unknown op");
	if (cap > 2 || join > 2) return fail(VGB_ERR_INVALID, "invalid cap/join");
	if (!finite_all(&width, 1)) return fail(VGB_ERR_NONFINITE, "non-finite width");
	DrawRec r;
	r.ocj = op | (cap << 8) | (join << 16);
	r.color = color;
	r.width = width;
	r.pathId = pathId;
	const uint32_t d = (uint32_t)b->draws.size();
	b->draws.push_back(r);
	const size_t j0 = b->jobSub.size(), nj = sub1 > sub0 ? sub1 - sub0 : 0;
	/* the force flag waits for a draw that has meshes (allocDrawCommand is what clears it, vg.cpp:5402); before the
	 * first draw it says nothing new: describe that case with vgb_final_cfg::cmd_vertices = 0 */
	const bool brk = b->forceNewCmd && d > 0 && nj > 0;
	b->drawScissor.push_back(b->curScissor | (brk ? 0x80000000u : 0u));
	if (brk) ++b->scissorBreaks;
	if (nj > 0) b->forceNewCmd = false;
	b->jobSub.resize(j0 + nj);
	b->jobDraw.resize(j0 + nj);
	for (size_t k = 0; k < nj; ++k) {
		b->jobSub[j0 + k] = sub0 + (uint32_t)k;
		b->jobDraw[j0 + k] = d;
	}
	b->dirty = true;
	return 0;
}

int vgb_batch_set_scissor(vgb_batch* b, const float rect[4], int forceNewCommand)
{
	if (!b || !rect) return fail(VGB_ERR_INVALID, "null argument");
	if (!finite_all(rect, 4)) return fail(VGB_ERR_NONFINITE, "non-finite scissor rect");
	/* (uint16_t)scissor[k], vg.cpp:5392-5395.  Negative sizes happen (ctxSetScissor under a mirrored transform); the
	 * reference's x86 build stores them modulo 2^16: spelled out so that it is defined behaviour and the same everywhere */
	uint16_t r[4];
	for (int k = 0; k < 4; ++k) r[k] = (rect[k] > -2147483520.0f && rect[k] < 2147483520.0f) ? (uint16_t)(int32_t)rect[k] : (uint16_t)0;
	/* one table entry per distinct rect; id 0 is the "no scissor set" entry and is never handed out */
	const uint64_t key = (uint64_t)r[0] | (uint64_t)r[1] << 16 | (uint64_t)r[2] << 32 | (uint64_t)r[3] << 48;
	const auto it = b->scissorIds.find(key);
	uint32_t id;
	if (it != b->scissorIds.end()) {
		id = it->second;
	} else {
		id = (uint32_t)(b->scissors.size() / 4);
		b->scissors.insert(b->scissors.end(), r, r + 4);
		b->scissorIds.emplace(key, id);
	}
	b->curScissor = id;
	b->forceNewCmd = b->forceNewCmd || forceNewCommand != 0;
	b->anyScissor = true;
	b->dirty = true;
	return 0;
}

int vgb_batch_scissor_rects(vgb_batch* b, uint16_t* rects4, uint32_t cap, uint32_t* n)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	const uint32_t cnt = (uint32_t)(b->scissors.size() / 4);
	if (n) *n = cnt;
	if (rects4) {
		if (cap < cnt) return fail(VGB_ERR_CAPACITY, "scissor table capacity too small");
		memcpy(rects4, b->scissors.data(), b->scissors.size() * sizeof(uint16_t));
	}
	return 0;
}

int vgb_batch_recorded_draw_scissor(vgb_batch* b, uint32_t* out)
{
	if (!b || !out) return fail(VGB_ERR_INVALID, "null argument");
	if (!b->drawScissor.empty()) memcpy(out, b->drawScissor.data(), b->drawScissor.size() * sizeof(uint32_t));
	return 0;
}

int vgb_batch_stroke(vgb_batch* b, uint32_t op, uint32_t color, float width, uint32_t cap, uint32_t join)
{
	if (!b || !b->inPath) return fail(VGB_ERR_INVALID, "stroke outside a path");
	if (op != VGB_OP_STROKE_AA && op != VGB_OP_STROKE && op != VGB_OP_STROKE_AA_THIN) return fail(VGB_ERR_INVALID, "not a stroke op");
	b->drawn = true;
	return addDraw(b, op, color, width, cap, join, b->curPath, b->curFirstSub, (uint32_t)b->subs.size());
}

int vgb_batch_convex_fill(vgb_batch* b, uint32_t color, int aa)
{
	if (!b || !b->inPath) return fail(VGB_ERR_INVALID, "fill outside a path");
	b->drawn = true;
	return addDraw(b, aa ? VGB_OP_CONVEX_FILL_AA : VGB_OP_CONVEX_FILL, color, 0.0f, 0, 0, b->curPath, b->curFirstSub, (uint32_t)b->subs.size());
}

int vgb_batch_add_cubic_strokes(vgb_batch* b, const float* p8, uint32_t n, float scale, float tol,
                                uint32_t op, uint32_t color, float width, uint32_t cap, uint32_t join)
{
	if (!b || (!p8 && n)) return fail(VGB_ERR_INVALID, "null argument");
	if (b->rawMode || b->bulkP8) return fail(VGB_ERR_STATE, "recording after vgb_batch_set_polylines / set_cubic_strokes");
	if (!(scale > 0.0f) || !(tol > 0.0f)) return fail(VGB_ERR_INVALID, "scale and tolerance must be positive");
	if (!finite_all(p8, 8 * (size_t)n)) return fail(VGB_ERR_NONFINITE, "non-finite coordinate");
	if (n == 0) return 0;
	b->closeSub();
	b->inPath = false;
	const uint32_t path0 = (uint32_t)b->pathScale.size();
	const uint32_t sub0 = (uint32_t)b->subs.size();
	const uint32_t cmd0 = (uint32_t)b->cmd.size();
	const uint32_t data0 = (uint32_t)b->data.size();
	b->cmd.resize(cmd0 + 2 * (size_t)n);
	for (uint32_t i = 0; i < n; ++i) {
		b->cmd[cmd0 + 2 * i] = VGB_CMD_MOVE_TO;
		b->cmd[cmd0 + 2 * i + 1] = VGB_CMD_CUBIC_TO;
	}
	b->data.insert(b->data.end(), p8, p8 + 8 * (size_t)n);
	b->subs.resize(sub0 + (size_t)n);
	for (uint32_t i = 0; i < n; ++i) {
		SubRec& r = b->subs[sub0 + i];
		r.cmdBegin = cmd0 + 2 * i;
		r.cmdEnd = r.cmdBegin + 2;
		r.dataBegin = data0 + 8 * i;
		r.pathId = path0 + i;
	}
	b->pathScale.insert(b->pathScale.end(), n, scale);
	b->pathTol.insert(b->pathTol.end(), n, tol);
	b->xformIdentityUpTo(b->pathScale.size());
	/* scale/tol are uniform over these paths, so one draw record serves all of them */
	return addDraw(b, op, color, width, cap, join, path0, sub0, sub0 + n);
}

int vgb_batch_add_paths(vgb_batch* b, uint32_t numPaths, const uint32_t* cmdOff, const uint32_t* dataOff, const uint8_t* cmd, const float* data,
                        const float* scale, const float* tol, const float* xform, const uint32_t* drawBegin, const uint32_t* drawOp,
                        const uint32_t* drawColor, const float* drawWidth, const uint8_t* drawCap, const uint8_t* drawJoin)
{
	if (!b || (numPaths && (!cmdOff || !dataOff || !scale || !tol || !drawBegin))) return fail(VGB_ERR_INVALID, "null argument");
	for (uint32_t p = 0; p < numPaths; ++p) {
		int rc = vgb_batch_begin_path(b, scale[p], tol[p], xform ? xform + 6 * (size_t)p : nullptr);
		const float* d = data + dataOff[p];
		for (uint32_t c = cmdOff[p]; c < cmdOff[p + 1] && !rc; ++c) {
			switch (cmd[c]) {
			case VGB_CMD_MOVE_TO: rc = vgb_batch_move_to(b, d[0], d[1]); d += 2; break;
			case VGB_CMD_LINE_TO: rc = vgb_batch_line_to(b, d[0], d[1]); d += 2; break;
			case VGB_CMD_CUBIC_TO: rc = vgb_batch_cubic_to(b, d[0], d[1], d[2], d[3], d[4], d[5]); d += 6; break;
			case VGB_CMD_QUAD_TO: rc = vgb_batch_quadratic_to(b, d[0], d[1], d[2], d[3]); d += 4; break;
			case VGB_CMD_ARC_TO: rc = vgb_batch_arc_to(b, d[0], d[1], d[2], d[3], d[4]); d += 5; break;
			case VGB_CMD_RECT: rc = vgb_batch_rect(b, d[0], d[1], d[2], d[3]); d += 4; break;
			case VGB_CMD_ROUNDED_RECT: rc = vgb_batch_rounded_rect(b, d[0], d[1], d[2], d[3], d[4]); d += 5; break;
			case VGB_CMD_ROUNDED_RECT_VARYING: rc = vgb_batch_rounded_rect_varying(b, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]); d += 8; break;
			case VGB_CMD_CIRCLE: rc = vgb_batch_circle(b, d[0], d[1], d[2]); d += 3; break;
			case VGB_CMD_ELLIPSE: rc = vgb_batch_ellipse(b, d[0], d[1], d[2], d[3]); d += 4; break;
			case VGB_CMD_ARC: rc = vgb_batch_arc(b, d[0], d[1], d[2], d[3], d[4], d[5] != 0.0f ? VGB_WINDING_CW : VGB_WINDING_CCW); d += 6; break;
			case VGB_CMD_POLYLINE: {
				uint32_t np;
				memcpy(&np, d, 4);
				rc = vgb_batch_polyline(b, d + 1, np);
				d += 1 + 2 * (size_t)np;
			} break;
			case VGB_CMD_CLOSE: rc = vgb_batch_close(b); break;
			default: rc = fail(VGB_ERR_INVALID, "unknown command");
			}
		}
		for (uint32_t k = drawBegin[p]; k < drawBegin[p + 1] && !rc; ++k) {
			if (drawOp[k] == VGB_OP_CONVEX_FILL_AA || drawOp[k] == VGB_OP_CONVEX_FILL) rc = vgb_batch_convex_fill(b, drawColor[k], drawOp[k] == VGB_OP_CONVEX_FILL_AA);
			else rc = vgb_batch_stroke(b, drawOp[k], drawColor[k], drawWidth[k], drawCap[k], drawJoin[k]);
		}
		if (rc) return rc;
	}
	return 0;
}

int vgb_batch_set_cubic_strokes(vgb_batch* b, const float* p8, uint32_t n, float scale, float tol,
                                uint32_t op, uint32_t color, float width, uint32_t cap, uint32_t join)
{
	if (!b || (!p8 && n)) return fail(VGB_ERR_INVALID, "null argument");
	if (!(scale > 0.0f) || !(tol > 0.0f)) return fail(VGB_ERR_INVALID, "scale and tolerance must be positive");
	if (op != VGB_OP_STROKE_AA && op != VGB_OP_STROKE && op != VGB_OP_STROKE_AA_THIN) return fail(VGB_ERR_INVALID, "not a stroke op");
	const float fr = b->fringe;
	b->clear(fr);
	b->bulkP8 = p8;
	b->bulkN = n;
	b->bulkScale = scale;
	b->bulkTol = tol;
	b->pathScale.assign(1, scale); /* path 0 carries the (uniform) flatten parameters of every bulk path */
	b->pathTol.assign(1, tol);
	return addDraw(b, op, color, width, cap, join, 0, 0, 0); /* one draw record, jobs are implicit (job i = sub-path i) */
}

int vgb_batch_add_rounded_rect_fills(vgb_batch* b, const float* q, uint32_t n, float scale, float tol, uint32_t color, int aa)
{
	if (!b || (!q && n)) return fail(VGB_ERR_INVALID, "null argument");
	if (b->rawMode) return fail(VGB_ERR_STATE, "recording after vgb_batch_set_polylines");
	if (!(scale > 0.0f) || !(tol > 0.0f)) return fail(VGB_ERR_INVALID, "scale and tolerance must be positive");
	if (!finite_all(q, 5 * (size_t)n)) return fail(VGB_ERR_NONFINITE, "non-finite coordinate");
	if (n == 0) return 0;
	b->closeSub();
	b->inPath = false;
	const uint32_t path0 = (uint32_t)b->pathScale.size();
	const uint32_t sub0 = (uint32_t)b->subs.size();
	static const float ident[6] = {1, 0, 0, 1, 0, 0};
	uint32_t made = 0;
	for (uint32_t i = 0; i < n; ++i) {
		const float* r = q + 5 * (size_t)i;
		b->pathScale.push_back(scale);
		b->pathTol.push_back(tol);
		if (r[4] < 0.1f && degenerateRect(r[2], r[3])) continue; /* empty path: no sub-path, no mesh */
		SubRec s;
		s.cmdBegin = (uint32_t)b->cmd.size();
		s.cmdEnd = s.cmdBegin + 1;
		s.dataBegin = (uint32_t)b->data.size();
		s.pathId = path0 + i;
		b->cmd.push_back(VGB_CMD_ROUNDED_RECT);
		b->data.insert(b->data.end(), r, r + 5);
		b->subs.push_back(s);
		++made;
	}
	b->xformIdentityUpTo(b->pathScale.size());
	return addDraw(b, aa ? VGB_OP_CONVEX_FILL_AA : VGB_OP_CONVEX_FILL, color, 0.0f, 0, 0, path0, sub0, sub0 + made);
}

int vgb_batch_set_polylines(vgb_batch* b, const float* pts, uint32_t numPts, const uint32_t* first, const uint32_t* num,
                            const uint8_t* closed, uint32_t numSub, float scale, float tol)
{
	if (!b || (!pts && numPts) || (numSub && (!first || !num || !closed))) return fail(VGB_ERR_INVALID, "null argument");
	if (!(scale > 0.0f) || !(tol > 0.0f)) return fail(VGB_ERR_INVALID, "scale and tolerance must be positive");
	if (!finite_all(pts, 2 * (size_t)numPts)) return fail(VGB_ERR_NONFINITE, "non-finite coordinate");
	for (uint32_t s = 0; s < numSub; ++s) {
		if ((uint64_t)first[s] + num[s] > numPts) return fail(VGB_ERR_INVALID, "sub-path exceeds the point array");
	}
	const float fr = b->fringe;
	b->clear(fr);
	b->rawMode = true;
	b->rawPts.assign(pts, pts + 2 * (size_t)numPts);
	b->rawFirst.assign(first, first + numSub);
	b->rawNum.assign(num, num + numSub);
	b->rawClosed.assign(closed, closed + numSub);
	b->pathScale.assign(1, scale);
	b->pathTol.assign(1, tol);
	return 0;
}

int vgb_batch_draw_polyline(vgb_batch* b, uint32_t sub, uint32_t op, uint32_t color, float width, uint32_t cap, uint32_t join)
{
	if (!b || !b->rawMode) return fail(VGB_ERR_STATE, "vgb_batch_set_polylines first");
	if (sub >= b->rawFirst.size()) return fail(VGB_ERR_INVALID, "sub-path index out of range");
	return addDraw(b, op, color, width, cap, join, 0, sub, sub + 1);
}

int vgb_batch_draw_all(vgb_batch* b, uint32_t op, uint32_t color, float width, uint32_t cap, uint32_t join)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	if (b->rawMode) return addDraw(b, op, color, width, cap, join, 0, 0, (uint32_t)b->rawFirst.size());
	/* recorded paths: only valid when every path shares scale/tol with path 0 (one draw record) */
	for (size_t i = 1; i < b->pathScale.size(); ++i) {
		if (b->pathScale[i] != b->pathScale[0] || b->pathTol[i] != b->pathTol[0]) return fail(VGB_ERR_INVALID, "draw_all needs uniform scale/tolerance");
	}
	b->closeSub();
	b->inPath = false;
	return addDraw(b, op, color, width, cap, join, 0, 0, (uint32_t)b->subs.size());
}

int vgb_batch_recorded_counts(vgb_batch* b, uint32_t* out6)
{
	if (!b || !out6) return fail(VGB_ERR_INVALID, "null argument");
	b->closeSub();
	if (b->bulkP8) {
		out6[0] = b->bulkN; out6[1] = b->bulkN; out6[2] = 2 * b->bulkN; out6[3] = 8 * b->bulkN; out6[4] = 1; out6[5] = b->bulkN;
		return 0;
	}
	out6[0] = (uint32_t)b->pathScale.size();
	out6[1] = (uint32_t)(b->rawMode ? b->rawFirst.size() : b->subs.size());
	out6[2] = (uint32_t)b->cmd.size();
	out6[3] = (uint32_t)b->data.size();
	out6[4] = (uint32_t)b->draws.size();
	out6[5] = (uint32_t)b->jobSub.size();
	return 0;
}

int vgb_batch_recorded_subpaths(vgb_batch* b, uint32_t* cb, uint32_t* ce, uint32_t* db, uint32_t* pid)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	b->closeSub();
	for (size_t i = 0; i < b->subs.size(); ++i) {
		if (cb) cb[i] = b->subs[i].cmdBegin;
		if (ce) ce[i] = b->subs[i].cmdEnd;
		if (db) db[i] = b->subs[i].dataBegin;
		if (pid) pid[i] = b->subs[i].pathId;
	}
	return 0;
}

int vgb_batch_recorded_jobs(vgb_batch* b, uint32_t* js, uint32_t* jd)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	for (size_t i = 0; i < b->jobSub.size(); ++i) {
		if (js) js[i] = b->jobSub[i];
		if (jd) jd[i] = b->jobDraw[i];
	}
	return 0;
}

int vgb_batch_recorded_streams(vgb_batch* b, uint8_t* cmd, float* data, float* scale, float* tol, float* xform6, uint32_t* drawOp, uint32_t* drawColor,
                               float* drawWidth, uint8_t* drawCap, uint8_t* drawJoin, uint32_t* drawPath)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	if (b->bulkP8 || b->rawMode) return fail(VGB_ERR_STATE, "recorded_streams: bulk / polyline batches have no command streams");
	b->closeSub();
	if (cmd && !b->cmd.empty()) memcpy(cmd, b->cmd.data(), b->cmd.size());
	if (data && !b->data.empty()) memcpy(data, b->data.data(), b->data.size() * sizeof(float));
	const size_t np = b->pathScale.size();
	for (size_t i = 0; i < np; ++i) {
		if (scale) scale[i] = b->pathScale[i];
		if (tol) tol[i] = b->pathTol[i];
		if (xform6) {
			static const float ident[6] = {1, 0, 0, 1, 0, 0};
			const float* m = (b->anyXform && b->pathXform.size() >= 6 * (i + 1)) ? &b->pathXform[6 * i] : ident;
			memcpy(xform6 + 6 * i, m, sizeof(float) * 6);
		}
	}
	for (size_t i = 0; i < b->draws.size(); ++i) {
		const DrawRec& d = b->draws[i];
		if (drawOp) drawOp[i] = d.ocj & 0xFF;
		if (drawCap) drawCap[i] = (uint8_t)((d.ocj >> 8) & 0xFF);
		if (drawJoin) drawJoin[i] = (uint8_t)((d.ocj >> 16) & 0xFF);
		if (drawColor) drawColor[i] = d.color;
		if (drawWidth) drawWidth[i] = d.width;
		if (drawPath) drawPath[i] = d.pathId;
	}
	return 0;
}

int vgb_batch_upload(vgb_batch* b)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	if (!b->pipe) return fail(VGB_ERR_NOGPU, "record-only batch (device -1): no GPU pipeline, and there is no CPU fallback");
	b->closeSub();
	HostTables t;
	t.fringe = b->fringe;
	t.pathScale = b->pathScale.data();
	t.pathTol = b->pathTol.data();
	t.numPaths = (uint32_t)b->pathScale.size();
	if (b->bulkP8) {
		t.bulkP8 = b->bulkP8;
		t.bulkN = b->bulkN;
		t.bulkScale = b->bulkScale;
		t.bulkTol = b->bulkTol;
		t.numSub = b->bulkN;
		t.numJobs = b->bulkN;
	} else if (b->rawMode) {
		static const float none[2] = {0.0f, 0.0f};
		t.rawPts = b->rawPts.empty() ? none : b->rawPts.data();
		t.numRawPts = (uint32_t)(b->rawPts.size() / 2);
		t.rawFirst = b->rawFirst.data();
		t.rawNum = b->rawNum.data();
		t.rawClosed = b->rawClosed.data();
		t.numSub = (uint32_t)b->rawFirst.size();
	} else {
		/* point slots for the single-pass flatten: a generous per-command estimate; a sub-path that
		 * outgrows its slot is re-flattened into an overflow arena on the device, so the estimate
		 * only affects memory and speed, never results */
		static const uint32_t kSlot[VGB_CMD__COUNT] = {1, 1, 48, 48, 80, 4, 200, 200, 160, 160, 80, 0, 0};
		const size_t ns = b->subs.size();
		b->slotFirst.resize(ns + 1);
		uint64_t acc = 0;
		for (size_t i = 0; i < ns; ++i) {
			b->slotFirst[i] = (uint32_t)acc;
			const SubRec& r = b->subs[i];
			uint32_t cap = 2;
			const float* d = b->data.data() + r.dataBegin;
			for (uint32_t c = r.cmdBegin; c < r.cmdEnd; ++c) {
				const uint8_t ty = b->cmd[c];
				if (ty == VGB_CMD_POLYLINE) {
					uint32_t np;
					memcpy(&np, d, 4);
					cap += np;
					d += 1 + 2 * (size_t)np;
				} else {
					cap += kSlot[ty];
					static const uint8_t kArity[VGB_CMD__COUNT] = {2, 2, 6, 4, 5, 4, 5, 8, 3, 4, 6, 0, 0};
					d += kArity[ty];
				}
			}
			acc += cap;
		}
		b->slotFirst[ns] = (uint32_t)acc;
		t.subSlotFirst = acc < 0xE0000000ull ? b->slotFirst.data() : nullptr;
		t.subRec = b->subs.data();
		t.numSub = (uint32_t)b->subs.size();
		t.cmd = b->cmd.data();
		t.numCmd = (uint32_t)b->cmd.size();
		t.data = b->data.data();
		t.numData = (uint32_t)b->data.size();
		b->xformIdentityUpTo(b->pathScale.size());
		t.pathXform = b->anyXform ? b->pathXform.data() : nullptr;
	}
	t.draws = b->draws.data();
	t.numDraws = (uint32_t)b->draws.size();
	if (b->anyScissor && b->drawScissor.size() == b->draws.size()) {
		t.drawScissor = b->drawScissor.data();
		t.numScissorBreaks = b->scissorBreaks;
	}
	if (!b->bulkP8) {
		t.jobSub = b->jobSub.data();
		t.jobDraw = b->jobDraw.data();
		t.numJobs = (uint32_t)b->jobSub.size();
	}
	std::string err;
	const int rc = b->pipe->upload(t, &err);
	if (rc) return fail(rc, err.c_str());
	b->dirty = false;
	return 0;
}

int vgb_batch_run(vgb_batch* b, uint32_t flags, uint64_t vertexBase)
{
	NEED_PIPE(b);
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	if (b->dirty) {
		const int rc = vgb_batch_upload(b);
		if (rc) return rc;
	}
	std::string err;
	const int rc = b->pipe->run(flags, vertexBase, &err);
	if (rc) return fail(rc, err.c_str());
	return 0;
}

int vgb_group_unique_id(void* id128)
{
	std::string err;
	const int rc = groupUniqueId(id128, &err);
	return rc ? fail(rc, err.c_str()) : 0;
}

int vgb_group_create(vgb_group** out, int device, int rank, int nranks, const void* id128)
{
	std::string err;
	const int rc = groupCreate(out, device, rank, nranks, id128, &err);
	return rc ? fail(rc, err.c_str()) : 0;
}

void vgb_group_destroy(vgb_group* g) { groupDestroy(g); }

int vgb_batch_run_merged(vgb_batch* b, vgb_group* g, uint32_t flags)
{
	NEED_PIPE(b);
	if (!g) return fail(VGB_ERR_INVALID, "null group");
	if (!(flags & VGB_RUN_INDEX_U32)) return fail(VGB_ERR_INVALID, "a merged index space needs uint32 indices: pass VGB_RUN_INDEX_U32");
	if (b->dirty) {
		const int rc = vgb_batch_upload(b);
		if (rc) return rc;
	}
	std::string err;
	const int rc = b->pipe->run(flags, 0, &err, groupHook(g));
	return rc ? fail(rc, err.c_str()) : 0;
}

int vgb_batch_merged_info(vgb_batch* b, vgb_merged* out)
{
	NEED_PIPE(b);
	if (!out) return fail(VGB_ERR_INVALID, "null argument");
	uint64_t v[4];
	std::string err;
	const int rc = b->pipe->mergedInfo(v, nullptr, &err);
	if (rc) return fail(rc, err.c_str());
	out->vertex_base = v[0];
	out->index_base = v[1];
	out->total_vertices = v[2];
	out->total_indices = v[3];
	return 0;
}

int vgb_group_gather(vgb_group* g, vgb_batch* b, int owner, void** pos, void** color, void** index, vgb_merged* info, float* ms)
{
	NEED_PIPE(b);
	if (!g) return fail(VGB_ERR_INVALID, "null group");
	uint64_t v[4] = {0, 0, 0, 0};
	std::string err;
	const int rc = groupGather(g, b->pipe, 4, owner, pos, color, index, v, ms, &err);
	if (rc) return fail(rc, err.c_str());
	if (info) {
		info->vertex_base = v[0];
		info->index_base = v[1];
		info->total_vertices = v[2];
		info->total_indices = v[3];
	}
	return 0;
}

int vgb_batch_sync(vgb_batch* b)
{
	NEED_PIPE(b);
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	std::string err;
	const int rc = b->pipe->sync(&err);
	if (rc) return fail(rc, err.c_str());
	return 0;
}

int vgb_batch_get_sizes(vgb_batch* b, vgb_sizes* out)
{
	NEED_PIPE(b);
	if (!b || !out) return fail(VGB_ERR_INVALID, "null argument");
	b->pipe->sizes(out);
	return 0;
}

int vgb_batch_device_ptrs(vgb_batch* b, void** pos, void** color, void** index, void** mvo, void** mio, void** pts, void** subs)
{
	NEED_PIPE(b);
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	b->pipe->devicePtrs(pos, color, index, mvo, mio, pts, subs);
	return 0;
}

int vgb_batch_download(vgb_batch* b, float* pos, uint32_t* color, void* index, uint64_t vcap, uint64_t icap, uint64_t* mvo, uint64_t* mio)
{
	NEED_PIPE(b);
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	std::string err;
	const int rc = b->pipe->download(pos, color, index, vcap, icap, mvo, mio, &err);
	if (rc) return fail(rc, err.c_str());
	return 0;
}

int vgb_batch_download_paths(vgb_batch* b, float* pts, uint64_t ptsCap, vgb_subpath* subs, uint32_t subCap)
{
	NEED_PIPE(b);
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	std::string err;
	const int rc = b->pipe->downloadPaths(pts, ptsCap, subs, subCap, &err);
	if (rc) return fail(rc, err.c_str());
	return 0;
}

int vgb_batch_replay(vgb_batch* b, const float* xforms, uint32_t n, const float* inv)
{
	NEED_PIPE(b);
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	if (n && !xforms) return fail(VGB_ERR_INVALID, "null transforms");
	for (size_t i = 0; i < (size_t)n * 6; ++i) {
		if (!std::isfinite(xforms[i])) return fail(VGB_ERR_NONFINITE, "non-finite instance transform");
	}
	for (int i = 0; inv && i < 6; ++i) {
		if (!std::isfinite(inv[i])) return fail(VGB_ERR_NONFINITE, "non-finite inverse transform");
	}
	std::string err;
	const int rc = b->pipe->replay(xforms, n, inv, &err);
	if (rc) return fail(rc, err.c_str());
	return 0;
}

int vgb_batch_replay_sizes(vgb_batch* b, vgb_sizes* out)
{
	NEED_PIPE(b);
	if (!b || !out) return fail(VGB_ERR_INVALID, "null argument");
	std::string err;
	const int rc = b->pipe->replaySizes(out, &err);
	if (rc) return fail(rc, err.c_str());
	return 0;
}

int vgb_batch_replay_device_ptrs(vgb_batch* b, void** pos, void** color, void** index)
{
	NEED_PIPE(b);
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	b->pipe->replayPtrs(pos, color, index);
	return 0;
}

int vgb_batch_download_replay(vgb_batch* b, float* pos, uint32_t* color, void* index, uint64_t vcap, uint64_t icap)
{
	NEED_PIPE(b);
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	std::string err;
	const int rc = b->pipe->downloadReplay(pos, color, index, vcap, icap, &err);
	if (rc) return fail(rc, err.c_str());
	return 0;
}

int vgb_batch_set_final_layout(vgb_batch* b, const vgb_final_cfg* cfg)
{
	NEED_PIPE(b);
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	std::string err;
	const int rc = b->pipe->setFinal(cfg, &err);
	if (rc) return fail(rc, err.c_str());
	return 0;
}

int vgb_batch_download_final(vgb_batch* b, int16_t* uv, uint64_t vcap, vgb_draw_cmd* cmds, uint32_t cmdCap, uint32_t* numCmds, void** uvDev)
{
	NEED_PIPE(b);
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	std::string err;
	const int rc = b->pipe->downloadFinal(uv, vcap, cmds, cmdCap, numCmds, uvDev, &err);
	if (rc) return fail(rc, err.c_str());
	return 0;
}

int vgb_batch_execute(vgb_batch* b, uint32_t flags, float* pos, uint32_t* color, void* index, uint64_t vcap, uint64_t icap,
                      uint64_t* mvo, uint64_t* mio, vgb_sizes* sizesOut)
{
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	int rc = vgb_batch_upload(b);
	if (rc) return rc;
	rc = vgb_batch_run(b, flags, 0);
	if (rc) return rc;
	rc = vgb_batch_download(b, pos, color, index, vcap, icap, mvo, mio);
	if (rc) return rc;
	if (sizesOut) b->pipe->sizes(sizesOut);
	return 0;
}

int vgb_batch_last_run_ms(vgb_batch* b, float* total, float* flatten, float* strokeCount, float* emit)
{
	NEED_PIPE(b);
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	std::string err;
	b->pipe->sync(&err);
	const RunTimes t = b->pipe->times();
	if (total) *total = t.total;
	if (flatten) *flatten = t.flatten;
	if (strokeCount) *strokeCount = t.strokeCount;
	if (emit) *emit = t.emit;
	return 0;
}

int vgb_batch_mark(vgb_batch* b, uint32_t slot)
{
	NEED_PIPE(b);
	if (!b) return fail(VGB_ERR_INVALID, "null batch");
	std::string err;
	const int rc = b->pipe->mark(slot, &err);
	if (rc) return fail(rc, err.c_str());
	return 0;
}

int vgb_batch_elapsed_ms(vgb_batch* b, uint32_t a, uint32_t c, float* ms)
{
	NEED_PIPE(b);
	if (!b || !ms) return fail(VGB_ERR_INVALID, "null argument");
	std::string err;
	const int rc = b->pipe->elapsed(a, c, ms, &err);
	if (rc) return fail(rc, err.c_str());
	return 0;
}

void* vgb_host_alloc(size_t bytes) { return hostAlloc(bytes); }
void vgb_host_free(void* p) { hostFree(p); }

int vgb_copy_to_host(void* dst, const void* src_device, size_t bytes)
{
	if (!dst || !src_device) return fail(VGB_ERR_INVALID, "null argument");
	if (copyToHost(dst, src_device, bytes)) return fail(VGB_ERR_CUDA, "cudaMemcpy device -> host failed");
	return 0;
}
void* vgb_batch_stream(vgb_batch* b) { return (b && b->pipe) ? b->pipe->stream() : nullptr; }

/* =========================================================================================
 * per-object API (include/vg/path.h:19-38, include/vg/stroker.h:11-70)
 * ======================================================================================= */
struct ObjAlloc
{
	vgb_realloc_fn fn;
	void* user;
	void* get(void* p, size_t bytes) const
	{
		if (fn) return fn(user, p, bytes, 16);
		if (bytes == 0) {
			free(p);
			return nullptr;
		}
		if (p) free(p); /* contents are always rewritten: no need to preserve */
		void* q = nullptr;
		if (posix_memalign(&q, 16, bytes)) return nullptr;
		return q;
	}
};

struct vgb_path
{
	ObjAlloc alloc;
	vgb_batch* batch; /* lazily created: needs a GPU */
	float scale, tol;
	bool flattened;
	float* verts;
	vgb_subpath* subs;
	uint32_t numVerts, numSubs;
	size_t vertCap, subCap;
};

struct vgb_stroker
{
	ObjAlloc alloc;
	vgb_batch* batch;
	float scale, tol, fringe;
	float* pos;
	uint32_t* col;
	uint16_t* idx;
	size_t vcap, icap;
};

static vgb_batch* lazyBatch(vgb_batch** slot)
{
	if (!*slot) {
		*slot = vgb_batch_create(0);
		if (!*slot) {
			/* the reference API has no error channel: fail loudly rather than fall back to a CPU path */
			fprintf(stderr, "libvgb200: %s\n", t_err.c_str());
			abort();
		}
	}
	return *slot;
}

vgb_path* vgb_path_create(vgb_realloc_fn fn, void* user)
{
	vgb_path* p = (vgb_path*)calloc(1, sizeof(vgb_path));
	p->alloc.fn = fn;
	p->alloc.user = user;
	p->scale = 1.0f;
	p->tol = 0.25f;
	return p;
}

void vgb_path_destroy(vgb_path* p)
{
	if (!p) return;
	if (p->verts) p->alloc.get(p->verts, 0);
	if (p->subs) p->alloc.get(p->subs, 0);
	vgb_batch_destroy(p->batch);
	free(p);
}

void vgb_path_reset(vgb_path* p, float scale, float tol)
{
	vgb_batch* b = lazyBatch(&p->batch);
	p->scale = scale;
	p->tol = tol;
	p->flattened = false;
	p->numVerts = p->numSubs = 0;
	vgb_batch_reset(b, 1.0f);
	vgb_batch_begin_path(b, scale, tol, nullptr);
}

#define PATH_CMD(call)          \
	do {                        \
		vgb_batch* b = lazyBatch(&p->batch); \
		if (!b->inPath) vgb_batch_begin_path(b, p->scale, p->tol, nullptr); \
		p->flattened = false;   \
		(void)(call);           \
	} while (0)

void vgb_path_move_to(vgb_path* p, float x, float y) { PATH_CMD(vgb_batch_move_to(b, x, y)); }
void vgb_path_line_to(vgb_path* p, float x, float y) { PATH_CMD(vgb_batch_line_to(b, x, y)); }
void vgb_path_cubic_to(vgb_path* p, float c1x, float c1y, float c2x, float c2y, float x, float y) { PATH_CMD(vgb_batch_cubic_to(b, c1x, c1y, c2x, c2y, x, y)); }
void vgb_path_quadratic_to(vgb_path* p, float cx, float cy, float x, float y) { PATH_CMD(vgb_batch_quadratic_to(b, cx, cy, x, y)); }
void vgb_path_arc_to(vgb_path* p, float x1, float y1, float x2, float y2, float r) { PATH_CMD(vgb_batch_arc_to(b, x1, y1, x2, y2, r)); }
void vgb_path_rect(vgb_path* p, float x, float y, float w, float h) { PATH_CMD(vgb_batch_rect(b, x, y, w, h)); }
void vgb_path_rounded_rect(vgb_path* p, float x, float y, float w, float h, float r) { PATH_CMD(vgb_batch_rounded_rect(b, x, y, w, h, r)); }
void vgb_path_rounded_rect_varying(vgb_path* p, float x, float y, float w, float h, float rtl, float rtr, float rbr, float rbl) { PATH_CMD(vgb_batch_rounded_rect_varying(b, x, y, w, h, rtl, rtr, rbr, rbl)); }
void vgb_path_circle(vgb_path* p, float x, float y, float r) { PATH_CMD(vgb_batch_circle(b, x, y, r)); }
void vgb_path_ellipse(vgb_path* p, float x, float y, float rx, float ry) { PATH_CMD(vgb_batch_ellipse(b, x, y, rx, ry)); }
void vgb_path_arc(vgb_path* p, float x, float y, float r, float a0, float a1, uint32_t dir) { PATH_CMD(vgb_batch_arc(b, x, y, r, a0, a1, dir)); }
void vgb_path_polyline(vgb_path* p, const float* c, uint32_t n) { PATH_CMD(vgb_batch_polyline(b, c, n)); }
void vgb_path_close(vgb_path* p) { PATH_CMD(vgb_batch_close(b)); }

static void pathFlatten(vgb_path* p)
{
	if (p->flattened) return;
	vgb_batch* b = lazyBatch(&p->batch);
	if (!b->inPath) vgb_batch_begin_path(b, p->scale, p->tol, nullptr);
	if (vgb_batch_run(b, 0, 0) != 0) {
		fprintf(stderr, "libvgb200: %s\n", t_err.c_str());
		abort();
	}
	vgb_sizes sz;
	vgb_batch_get_sizes(b, &sz);
	if (sz.num_points + 2 > p->vertCap) {
		p->vertCap = (size_t)sz.num_points * 3 / 2 + 16; /* growth x1.5 like path.cpp:751 */
		p->verts = (float*)p->alloc.get(p->verts, p->vertCap * 2 * sizeof(float));
	}
	if (sz.num_subpaths + 1 > p->subCap) {
		p->subCap = sz.num_subpaths + 16;
		p->subs = (vgb_subpath*)p->alloc.get(p->subs, p->subCap * sizeof(vgb_subpath));
	}
	if (vgb_batch_download_paths(b, p->verts, p->vertCap, p->subs, (uint32_t)p->subCap) != 0) {
		fprintf(stderr, "libvgb200: %s\n", t_err.c_str());
		abort();
	}
	p->numVerts = (uint32_t)sz.num_points;
	p->numSubs = sz.num_subpaths;
	p->flattened = true;
}

const float* vgb_path_get_vertices(vgb_path* p) { pathFlatten(p); return p->verts; }
uint32_t vgb_path_get_num_vertices(vgb_path* p) { pathFlatten(p); return p->numVerts; }
const vgb_subpath* vgb_path_get_subpaths(vgb_path* p) { pathFlatten(p); return p->subs; }
uint32_t vgb_path_get_num_subpaths(vgb_path* p) { pathFlatten(p); return p->numSubs; }

vgb_stroker* vgb_stroker_create(vgb_realloc_fn fn, void* user)
{
	vgb_stroker* s = (vgb_stroker*)calloc(1, sizeof(vgb_stroker));
	s->alloc.fn = fn;
	s->alloc.user = user;
	s->scale = 1.0f;
	s->tol = 0.25f;
	s->fringe = 1.0f;
	return s;
}

void vgb_stroker_destroy(vgb_stroker* s)
{
	if (!s) return;
	if (s->pos) s->alloc.get(s->pos, 0);
	if (s->col) s->alloc.get(s->col, 0);
	if (s->idx) s->alloc.get(s->idx, 0);
	vgb_batch_destroy(s->batch);
	free(s);
}

void vgb_stroker_reset(vgb_stroker* s, float scale, float tol, float fringe)
{
	s->scale = scale;
	s->tol = tol;
	s->fringe = fringe;
}

static void strokerRun(vgb_stroker* s, vgb_mesh* mesh, uint32_t op, const float* vtx, uint32_t n, int closed, uint32_t color, float width, uint32_t cap, uint32_t join, bool hasColor)
{
	vgb_batch* b = lazyBatch(&s->batch);
	memset(mesh, 0, sizeof(*mesh));
	const uint32_t first = 0;
	const uint8_t cl = closed ? 1 : 0;
	vgb_batch_reset(b, s->fringe);
	int rc = vgb_batch_set_polylines(b, vtx, n, &first, &n, &cl, 1, s->scale, s->tol);
	if (!rc) rc = vgb_batch_draw_polyline(b, 0, op, color, width, cap, join);
	if (!rc) rc = vgb_batch_run(b, VGB_RUN_INDEX_U16, 0);
	if (rc == VGB_ERR_INVALID || rc == VGB_ERR_NONFINITE) {
		/* input-dependent (a NaN vertex out of the caller's transform, a non-finite width, scale <= 0): the reference
		 * has no error channel and produces a degenerate mesh; here: an EMPTY mesh, the reason in vgb_last_error() */
		return;
	}
	if (rc) {
		fprintf(stderr, "libvgb200: %s\n", t_err.c_str()); /* device failure: there is no CPU path to fall back to */
		abort();
	}
	vgb_sizes sz;
	vgb_batch_get_sizes(b, &sz);
	if (sz.num_vertices + 2 > s->vcap) {
		s->vcap = (size_t)sz.num_vertices * 3 / 2 + 64;
		s->pos = (float*)s->alloc.get(s->pos, s->vcap * 2 * sizeof(float));
		s->col = (uint32_t*)s->alloc.get(s->col, s->vcap * sizeof(uint32_t));
	}
	if (sz.num_indices + 8 > s->icap) {
		s->icap = (size_t)sz.num_indices * 3 / 2 + 64;
		s->idx = (uint16_t*)s->alloc.get(s->idx, s->icap * sizeof(uint16_t));
	}
	rc = vgb_batch_download(b, s->pos, s->col, s->idx, s->vcap, s->icap, nullptr, nullptr);
	if (rc) {
		fprintf(stderr, "libvgb200: %s\n", t_err.c_str());
		abort();
	}
	mesh->pos = s->pos;
	mesh->color = hasColor ? s->col : nullptr;
	mesh->index = s->idx;
	mesh->num_vertices = (uint32_t)sz.num_vertices;
	mesh->num_indices = (uint32_t)sz.num_indices;
}

void vgb_stroker_polyline_stroke(vgb_stroker* s, vgb_mesh* mesh, const float* vtx, uint32_t n, int closed, float width, uint32_t cap, uint32_t join)
{
	strokerRun(s, mesh, VGB_OP_STROKE, vtx, n, closed, 0xFFFFFFFFu, width, cap, join, false);
}
void vgb_stroker_polyline_stroke_aa(vgb_stroker* s, vgb_mesh* mesh, const float* vtx, uint32_t n, int closed, uint32_t color, float width, uint32_t cap, uint32_t join)
{
	strokerRun(s, mesh, VGB_OP_STROKE_AA, vtx, n, closed, color, width, cap, join, true);
}
void vgb_stroker_polyline_stroke_aa_thin(vgb_stroker* s, vgb_mesh* mesh, const float* vtx, uint32_t n, int closed, uint32_t color, uint32_t cap, uint32_t join)
{
	strokerRun(s, mesh, VGB_OP_STROKE_AA_THIN, vtx, n, closed, color, 0.0f, cap, join, true);
}
void vgb_stroker_convex_fill(vgb_stroker* s, vgb_mesh* mesh, const float* vtx, uint32_t n)
{
	strokerRun(s, mesh, VGB_OP_CONVEX_FILL, vtx, n, 1, 0xFFFFFFFFu, 0.0f, 0, 0, false);
	mesh->pos = vtx; /* the reference returns the caller's own polygon (stroker.cpp:360) */
}
void vgb_stroker_convex_fill_aa(vgb_stroker* s, vgb_mesh* mesh, const float* vtx, uint32_t n, uint32_t color)
{
	strokerRun(s, mesh, VGB_OP_CONVEX_FILL_AA, vtx, n, 1, color, 0.0f, 0, 0, true);
}

} /* extern "C" */
