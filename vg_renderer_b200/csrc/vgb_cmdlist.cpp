/*
 * vgb_cmdlist.cpp -- command-list front end: decodes a vg-renderer CommandList byte
 * stream and records it into a batch, so that a whole recorded frame becomes ONE device run instead of
 * one Path/Stroker round per command.
 *
 * Format (reference src/vg.cpp): records of { CommandHeader{ uint32 type; uint32 size }, padded to 16
 * bytes; payload of `size` bytes, itself padded to 16 } written by clAllocCommand (vg.cpp:5694-5723);
 * command ids CommandType::Enum (vg.cpp:177-247); payload layouts are the CMD_WRITE sequences of the
 * cl* writers (vg.cpp:2402-2960).  Execution follows ctxSubmitCommandList (vg.cpp:4273-4637) and the
 * ctx* functions it calls, restricted to what lands on the hot path:
 *   path commands                       -> vgb_batch_* builders        (vg.cpp:2969-3058)
 *   FillPathColor, convex               -> vgb_batch_convex_fill       (vg.cpp:3061-3132)
 *   StrokePathColor                     -> vgb_batch_stroke            (vg.cpp:3401-3492)
 *   Push/PopState, Transform*, SetViewBox, SetGlobalAlpha, *Scissor (culling only)
 *                                                                      (vg.cpp:3934-4127, 4927-4944)
 *   SubmitCommandList                   -> the named list, in place (vg.cpp:4620-4626, depth limit 4278-4281)
 * Gradient / image-pattern paints, concave fills, text, clip regions and indexed tri-lists are outside
 * SURVEY 8's scope: they are skipped and counted in vgb_cl_stats.unsupported.
 * Host code only; floats follow the reference's expression order (built with -ffp-contract=off).
 */
#include "vgb.h"
#include "vgb_math.h"

#include <string.h>
#include <vector>

extern "C" void vgb_set_error_message(const char* msg); /* vgb_capi.cpp */

namespace
{
enum ClType : uint32_t { /* CommandType::Enum, vg.cpp:179-246 */
	CL_BeginPath = 0, CL_MoveTo, CL_LineTo, CL_CubicTo, CL_QuadraticTo, CL_ArcTo, CL_Arc, CL_Rect, CL_RoundedRect,
	CL_RoundedRectVarying, CL_Circle, CL_Ellipse, CL_Polyline, CL_ClosePath,
	CL_FillPathColor, CL_FillPathGradient, CL_FillPathImagePattern, CL_StrokePathColor, CL_StrokePathGradient, CL_StrokePathImagePattern,
	CL_IndexedTriList,
	CL_BeginClip, CL_EndClip, CL_ResetClip, CL_CreateLinearGradient, CL_CreateBoxGradient, CL_CreateRadialGradient, CL_CreateImagePattern,
	CL_PushState, CL_PopState, CL_ResetScissor, CL_SetScissor, CL_IntersectScissor,
	CL_TransformIdentity, CL_TransformScale, CL_TransformTranslate, CL_TransformRotate, CL_TransformMult, CL_SetViewBox, CL_SetGlobalAlpha,
	CL_Text, CL_TextBox, CL_SubmitCommandList, CL__Count
};
constexpr uint32_t kAlign = 16;      /* VG_CONFIG_COMMAND_LIST_ALIGNMENT, vg.cpp:40 */
constexpr uint32_t kHeaderSize = 16; /* alignSize(sizeof(CommandHeader), 16), vg.cpp:708 */
constexpr uint32_t kMaxStates = 32;  /* ContextConfig::m_MaxStateStackSize default, vg.cpp:723 */

struct State /* the members of vg.cpp's State this path reads */
{
	float mtx[6];
	float scissor[4];
	float globalAlpha;
	float avgScale;
};

void updateState(State* s) /* vg.cpp:4927-4944 (font scale not needed) */
{
	const float* t = s->mtx;
	const float sx = VGB_SQRT(t[0] * t[0] + t[2] * t[2]);
	const float sy = VGB_SQRT(t[1] * t[1] + t[3] * t[3]);
	s->avgScale = (sx + sy) * 0.5f;
}

inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); } /* bx::clamp = min(max(v, lo), hi) */
inline float maxf(float a, float b) { return a > b ? a : b; }
inline float minf(float a, float b) { return a < b ? a : b; }

template<class T> inline T rd(const uint8_t*& p)
{
	T v;
	memcpy(&v, p, sizeof(T));
	p += sizeof(T);
	return v;
}

int fail(int code, const char* msg)
{
	vgb_set_error_message(msg);
	return code;
}
} // namespace

namespace
{
/* (uint16_t)scissor[k] as vg.cpp:5392-5395 stores it.  A mirrored transform gives ctxSetScissor a negative size
 * (vg.cpp:3991-4005), and the reference's x86 build turns that into the value modulo 2^16 (cvttss2si, then the low half):
 * spelled out here so that it is defined behaviour and the same everywhere (the batch does the same) */
static inline uint16_t u16trunc(float v)
{
	return (v > -2147483520.0f && v < 2147483520.0f) ? (uint16_t)(int32_t)v : (uint16_t)0;
}

struct Run /* what one top-level submit shares with the lists it nests */
{
	vgb_batch* b;
	vgb_cl_context* ctx;
	vgb_cl_stats st;
	std::vector<State> stack;
	float fringe;
	bool inClip;
	bool deadPath = false;   /* the current path was begun under a collapsed transform (avgScale <= 0) */
	uint32_t depth, maxDepth;
	/* scissor of the last recorded draw (uint16-truncated like DrawCommand::m_ScissorRect, vg.cpp:5392-5395) */
	uint16_t lastScissor[4] = {0, 0, 0, 0};
	bool anyDraw = false;
	bool force = false; /* m_ForceNewDrawCommand: a scissor command ran since the last recorded draw */
	/* before a draw is recorded: hand the state's scissor to the batch (vg.cpp stamps it on the DrawCommand) */
	void noteDraw(const float* sc)
	{
		const uint16_t r[4] = {u16trunc(sc[0]), u16trunc(sc[1]), u16trunc(sc[2]), u16trunc(sc[3])};
		if (anyDraw && force) ++st.scissor_changes;
		/* a non-finite rect (NaN operands in the list) is refused by the batch: the draw then runs under the previous
		 * rect and the mark stays up for the next one */
		const bool ok = vgb_batch_set_scissor(b, sc, force ? 1 : 0) == 0;
		memcpy(lastScissor, r, sizeof(r));
		anyDraw = true;
		if (ok) force = false;
	}
	/* PopState: a new command only if the restored rect differs from the last draw command's (vg.cpp:3950-3966) */
	void notePop(const float* sc)
	{
		const uint16_t r[4] = {u16trunc(sc[0]), u16trunc(sc[1]), u16trunc(sc[2]), u16trunc(sc[3])};
		if (anyDraw && memcmp(r, lastScissor, sizeof(r)) != 0) force = true;
	}
};

int runList(Run& R, const void* buffer, uint32_t size, uint32_t listFlags)
{
	if (size % kAlign) return fail(VGB_ERR_INVALID, "command buffer size is not a multiple of 16");
	if (R.depth >= R.maxDepth) return 0; /* vg.cpp:4278-4281: recursion limit reached, the list is not executed */
	++R.depth;
	vgb_batch* b = R.b;
	vgb_cl_context* ctx = R.ctx;
	vgb_cl_stats& st = R.st;
	std::vector<State>& stack = R.stack;
	bool& inClip = R.inClip;
	const float fringe = R.fringe;
	const bool cullCmds = (listFlags & 2u) != 0; /* CommandListFlags::AllowCommandCulling, vg.cpp:4307 */
	bool skipCmds = false;                        /* local to one execution of one list, vg.cpp:4324 */
	int rc = 0;

	const uint8_t* cmd = (const uint8_t*)buffer;
	const uint8_t* end = cmd + size;
	while (cmd < end && rc == 0) {
		if ((size_t)(end - cmd) < kHeaderSize) return fail(VGB_ERR_INVALID, "truncated command header");
		uint32_t type, dataSize;
		memcpy(&type, cmd, 4);
		memcpy(&dataSize, cmd + 4, 4);
		cmd += kHeaderSize;
		if (dataSize % kAlign || dataSize > (size_t)(end - cmd)) return fail(VGB_ERR_INVALID, "bad command size");
		if (type >= CL__Count) return fail(VGB_ERR_INVALID, "unknown command");
		const uint8_t* next = cmd + dataSize;
		const uint8_t* p = cmd;
		float f[8] = {0, 0, 0, 0, 0, 0, 0, 0}; /* leading floats of the payload (the caller's buffer need not be aligned) */
		memcpy(f, cmd, dataSize < sizeof(f) ? dataSize : sizeof(f));
		++st.commands;
		State* s = &stack.back();
		const bool isStroker = type >= CL_FillPathColor && type <= CL_StrokePathImagePattern;
		if (skipCmds && isStroker) { /* vg.cpp:4338-4341 */
			++st.culled;
			cmd = next;
			continue;
		}
		if (R.deadPath && (isStroker || (type >= CL_MoveTo && type <= CL_ClosePath))) {
			if (isStroker) ++st.culled;
			cmd = next;
			continue;
		}
		auto need = [&](uint32_t bytes) { return dataSize >= bytes; };
#define BAD fail(VGB_ERR_INVALID, "malformed command payload")
		switch (type) {
		case CL_BeginPath: /* ctxBeginPath, vg.cpp:2969-2981: flatten scale = the state's average scale NOW */
			/* A collapsed transform (TransformScale(0, 0) ...: avgScale 0) is legal in the reference: tessTol becomes
			 * inf, every stroke gets zero alpha, every fill a zero-area polygon -- nothing visible, and the list keeps
			 * executing.  Here such a path is DEAD: its commands and draws are skipped (counted as culled). */
			R.deadPath = !(s->avgScale > 0.0f) || !(s->avgScale < 3.0e38f);
			if (!R.deadPath) rc = vgb_batch_begin_path(b, s->avgScale, ctx->tess_tolerance, nullptr);
			++st.paths;
			break;
		case CL_ClosePath: rc = vgb_batch_close(b); break;
		case CL_MoveTo: rc = need(8) ? vgb_batch_move_to(b, f[0], f[1]) : BAD; break;
		case CL_LineTo: rc = need(8) ? vgb_batch_line_to(b, f[0], f[1]) : BAD; break;
		case CL_CubicTo: rc = need(24) ? vgb_batch_cubic_to(b, f[0], f[1], f[2], f[3], f[4], f[5]) : BAD; break;
		case CL_QuadraticTo: rc = need(16) ? vgb_batch_quadratic_to(b, f[0], f[1], f[2], f[3]) : BAD; break;
		case CL_ArcTo: rc = need(20) ? vgb_batch_arc_to(b, f[0], f[1], f[2], f[3], f[4]) : BAD; break;
		case CL_Arc: {
			if (!need(24)) { rc = BAD; break; }
			uint32_t dir;
			memcpy(&dir, cmd + 20, 4);
			rc = vgb_batch_arc(b, f[0], f[1], f[2], f[3], f[4], dir);
		} break;
		case CL_Rect: rc = need(16) ? vgb_batch_rect(b, f[0], f[1], f[2], f[3]) : BAD; break;
		case CL_RoundedRect: rc = need(20) ? vgb_batch_rounded_rect(b, f[0], f[1], f[2], f[3], f[4]) : BAD; break;
		case CL_RoundedRectVarying: rc = need(32) ? vgb_batch_rounded_rect_varying(b, f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7]) : BAD; break;
		case CL_Circle: rc = need(12) ? vgb_batch_circle(b, f[0], f[1], f[2]) : BAD; break;
		case CL_Ellipse: rc = need(16) ? vgb_batch_ellipse(b, f[0], f[1], f[2], f[3]) : BAD; break;
		case CL_Polyline: {
			if (!need(4)) { rc = BAD; break; }
			const uint32_t n = rd<uint32_t>(p);
			if ((uint64_t)n * 8 + 4 > dataSize) { rc = BAD; break; }
			std::vector<float> pts(2 * (size_t)n);
			if (n) memcpy(pts.data(), p, sizeof(float) * 2 * n);
			rc = vgb_batch_polyline(b, pts.data(), n);
		} break;
		case CL_FillPathColor: {
			if (!need(8)) { rc = BAD; break; }
			const uint32_t flags = rd<uint32_t>(p);
			const uint32_t color = rd<uint32_t>(p);
			if (inClip || (flags & 0x01u) != 0) { /* clip geometry / PathType::Concave (libtess2): not on this path */
				++st.unsupported;
				R.force = true; /* an ordering marker: the caller's own geometry goes between two draw commands */
				break;
			}
			/* ctxFillPathColor, vg.cpp:3071-3076 */
			const uint32_t a = (uint8_t)(s->globalAlpha * (float)(uint8_t)(color >> 24));
			const uint32_t col = (color & 0x00FFFFFFu) | (a << 24);
			if (a == 0) {
				++st.zero_alpha;
				break;
			}
			vgb_batch_set_path_xform(b, s->mtx); /* transformPath: the state transform at the path's FIRST draw, vg.cpp:4957-4975 */
			R.noteDraw(s->scissor);
			rc = vgb_batch_convex_fill(b, col, (flags & 0x04u) != 0);
			++st.draws;
		} break;
		case CL_StrokePathColor: {
			if (!need(12)) { rc = BAD; break; }
			const float width = rd<float>(p);
			const uint32_t flags = rd<uint32_t>(p);
			const uint32_t color = rd<uint32_t>(p);
			if (inClip) {
				++st.unsupported;
				R.force = true;
				break;
			}
			/* ctxStrokePathColor, vg.cpp:3411-3437 */
			const float scaled = (flags & (1u << 5)) != 0 ? width : clampf(width * s->avgScale, 0.0f, 200.0f);
			const bool thin = scaled <= fringe;
			float alphaScale = s->globalAlpha;
			if (thin) {
				const float c = clampf(scaled, 0.0f, fringe);
				alphaScale = s->globalAlpha * (c * c);
			}
			const uint32_t a = (uint8_t)(alphaScale * (float)(uint8_t)(color >> 24));
			const uint32_t col = (color & 0x00FFFFFFu) | (a << 24);
			if (a == 0) {
				++st.zero_alpha;
				break;
			}
			const bool aa = (flags & 0x10u) != 0;
			const uint32_t op = !aa ? VGB_OP_STROKE : (thin ? VGB_OP_STROKE_AA_THIN : VGB_OP_STROKE_AA);
			vgb_batch_set_path_xform(b, s->mtx);
			R.noteDraw(s->scissor);
			rc = vgb_batch_stroke(b, op, col, thin ? fringe : scaled, (flags >> 2) & 3u, flags & 3u);
			++st.draws;
		} break;
		case CL_FillPathGradient: case CL_FillPathImagePattern: case CL_StrokePathGradient: case CL_StrokePathImagePattern:
		case CL_IndexedTriList: case CL_Text: case CL_TextBox: case CL_ResetClip:
			/* draws of another DrawCommand type (or clip state) in the reference: what the batch records around them
			 * must not share a draw command, so the caller can put its own geometry in between */
			++st.unsupported;
			R.force = true;
			break;
		case CL_CreateLinearGradient: case CL_CreateBoxGradient: case CL_CreateRadialGradient: case CL_CreateImagePattern:
			++st.unsupported;
			break;
		case CL_SubmitCommandList: { /* vg.cpp:4620-4626 */
			if (!need(2)) { rc = BAD; break; }
			const uint16_t id = rd<uint16_t>(p);
			if (!ctx->lists) { /* no table given: nesting not available to this call */
				++st.unsupported;
				break;
			}
			if (id >= ctx->num_lists || !ctx->lists[id].commands) {
				break; /* isCommandListHandleValid() false: skipped */
			}
			++st.nested;
			rc = runList(R, ctx->lists[id].commands, ctx->lists[id].size, ctx->lists[id].flags);
			s = &stack.back(); /* the nested list may have pushed / popped states */
		} break;
		case CL_BeginClip: inClip = true; ++st.unsupported; R.force = true; break;
		case CL_EndClip: inClip = false; R.force = true; break;
		case CL_PushState: /* vg.cpp:3934-3944 */
			if (stack.size() >= kMaxStates) { rc = fail(VGB_ERR_INVALID, "state stack overflow"); break; }
			stack.push_back(stack.back());
			break;
		case CL_PopState: /* vg.cpp:3946-3969, 4560-4567 */
			if (stack.size() <= 1) { rc = fail(VGB_ERR_INVALID, "state stack underflow"); break; }
			stack.pop_back();
			R.notePop(stack.back().scissor);
			if (cullCmds) skipCmds = stack.back().scissor[2] < 1.0f || stack.back().scissor[3] < 1.0f;
			break;
		case CL_ResetScissor: /* vg.cpp:3971-3979 */
			s->scissor[0] = s->scissor[1] = 0.0f;
			s->scissor[2] = ctx->canvas_width;
			s->scissor[3] = ctx->canvas_height;
			R.force = true;
			skipCmds = false;
			break;
		case CL_SetScissor: { /* vg.cpp:3981-4003 */
			if (!need(16)) { rc = BAD; break; }
			const float* t = s->mtx;
			const float px = t[0] * f[0] + t[2] * f[1] + t[4], py = t[1] * f[0] + t[3] * f[1] + t[5];
			const float sx = t[0] * f[2] + t[2] * f[3], sy = t[1] * f[2] + t[3] * f[3];
			const float minx = clampf(px, 0.0f, ctx->canvas_width), miny = clampf(py, 0.0f, ctx->canvas_height);
			const float maxx = clampf(px + sx, 0.0f, ctx->canvas_width), maxy = clampf(py + sy, 0.0f, ctx->canvas_height);
			s->scissor[0] = minx; s->scissor[1] = miny; s->scissor[2] = maxx - minx; s->scissor[3] = maxy - miny;
			R.force = true;
			if (cullCmds) skipCmds = s->scissor[2] < 1.0f || s->scissor[3] < 1.0f;
		} break;
		case CL_IntersectScissor: { /* vg.cpp:4005-4033 */
			if (!need(16)) { rc = BAD; break; }
			const float* t = s->mtx;
			const float px = t[0] * f[0] + t[2] * f[1] + t[4], py = t[1] * f[0] + t[3] * f[1] + t[5];
			const float sx = t[0] * f[2] + t[2] * f[3], sy = t[1] * f[2] + t[3] * f[3];
			const float minx = maxf(px, s->scissor[0]), miny = maxf(py, s->scissor[1]);
			const float maxx = minf(px + sx, s->scissor[0] + s->scissor[2]), maxy = minf(py + sy, s->scissor[1] + s->scissor[3]);
			const float w = maxf(0.0f, maxx - minx), h = maxf(0.0f, maxy - miny);
			s->scissor[0] = minx; s->scissor[1] = miny; s->scissor[2] = w; s->scissor[3] = h;
			R.force = true;
			if (cullCmds) skipCmds = !(w >= 1.0f && h >= 1.0f);
		} break;
		case CL_TransformIdentity: /* vg.cpp:4035-4045 */
			s->mtx[0] = 1.0f; s->mtx[1] = 0.0f; s->mtx[2] = 0.0f; s->mtx[3] = 1.0f; s->mtx[4] = 0.0f; s->mtx[5] = 0.0f;
			updateState(s);
			break;
		case CL_TransformScale: /* vg.cpp:4047-4055 */
			if (!need(8)) { rc = BAD; break; }
			s->mtx[0] = f[0] * s->mtx[0]; s->mtx[1] = f[0] * s->mtx[1]; s->mtx[2] = f[1] * s->mtx[2]; s->mtx[3] = f[1] * s->mtx[3];
			updateState(s);
			break;
		case CL_TransformTranslate: /* vg.cpp:4057-4063 */
			if (!need(8)) { rc = BAD; break; }
			s->mtx[4] += s->mtx[0] * f[0] + s->mtx[2] * f[1];
			s->mtx[5] += s->mtx[1] * f[0] + s->mtx[3] * f[1];
			updateState(s);
			break;
		case CL_TransformRotate: { /* vg.cpp:4065-4083; bx::cos / bx::sin = the shared math of this build */
			if (!need(4)) { rc = BAD; break; }
			const float c = vgb_cos(f[0]), sn = vgb_sin(f[0]);
			const float* t = s->mtx;
			const float m[6] = {c * t[0] + sn * t[2], c * t[1] + sn * t[3], -sn * t[0] + c * t[2], -sn * t[1] + c * t[3], t[4], t[5]};
			memcpy(s->mtx, m, sizeof(m));
			updateState(s);
		} break;
		case CL_TransformMult: { /* vg.cpp:4085-4101, multiplyMatrix3 vg_util.h:36-44 */
			if (!need(28)) { rc = BAD; break; }
			uint32_t order;
			memcpy(&order, cmd + 24, 4);
			const float* a = order == 1 /* Post */ ? s->mtx : f;
			const float* bb = order == 1 ? f : s->mtx;
			const float r[6] = {a[0] * bb[0] + a[2] * bb[1], a[1] * bb[0] + a[3] * bb[1], a[0] * bb[2] + a[2] * bb[3],
			                    a[1] * bb[2] + a[3] * bb[3], a[0] * bb[4] + a[2] * bb[5] + a[4], a[1] * bb[4] + a[3] * bb[5] + a[5]};
			memcpy(s->mtx, r, sizeof(r));
			updateState(s);
		} break;
		case CL_SetViewBox: { /* vg.cpp:4103-4121 */
			if (!need(16)) { rc = BAD; break; }
			const float scaleX = ctx->canvas_width / f[2], scaleY = ctx->canvas_height / f[3];
			float* t = s->mtx;
			t[0] = scaleX * t[0]; t[1] = scaleX * t[1]; t[2] = scaleY * t[2]; t[3] = scaleY * t[3];
			t[4] -= t[0] * f[0] + t[2] * f[1];
			t[5] -= t[1] * f[0] + t[3] * f[1];
			updateState(s);
		} break;
		case CL_SetGlobalAlpha: /* vg.cpp:4123-4127 */
			if (!need(4)) { rc = BAD; break; }
			s->globalAlpha = f[0];
			break;
		default: break;
		}
		cmd = next;
	}
	--R.depth;
#undef BAD
	return rc;
}
} // namespace

extern "C" int vgb_batch_submit_command_list(vgb_batch* b, const void* buffer, uint32_t size, vgb_cl_context* ctx, vgb_cl_stats* stats)
{
	if (!b || !ctx || (!buffer && size)) return fail(VGB_ERR_INVALID, "null argument");
	Run R;
	R.b = b;
	R.ctx = ctx;
	memset(&R.st, 0, sizeof(R.st));
	R.stack.resize(1);
	memcpy(R.stack[0].mtx, ctx->xform, sizeof(float) * 6);
	memcpy(R.stack[0].scissor, ctx->scissor, sizeof(float) * 4);
	R.stack[0].globalAlpha = ctx->global_alpha;
	updateState(&R.stack[0]);
	R.fringe = vgb_batch_fringe(b);
	R.inClip = false;
	R.depth = 0;
	R.maxDepth = ctx->max_depth ? ctx->max_depth : 16u; /* ContextConfig::m_MaxCommandListDepth default, vg.cpp:728 */
	const int rc = runList(R, buffer, size, ctx->flags);
	/* VG_CONFIG_COMMAND_LIST_PRESERVE_STATE is 0 by default (vg.h:34-35): state changes outlive the list */
	const State& top = R.stack.back();
	memcpy(ctx->xform, top.mtx, sizeof(float) * 6);
	memcpy(ctx->scissor, top.scissor, sizeof(float) * 4);
	ctx->global_alpha = top.globalAlpha;
	R.st.state_depth = (uint32_t)R.stack.size() - 1;
	if (stats) *stats = R.st;
	return rc;
}
