/*
 * vgctx_capi.cpp -- the reference's WHOLE context (src/vg.cpp: command-list recording and execution, state stack,
 * the per-path fill / stroke executors, the merge step createDrawCommand_VertexColor, vg::end's upload and draw
 * submission) compiled where it lies and driven through its own public API, against a record-only bgfx stand-in
 * (oracle/bgfx_stub) and the bx shim.  TEST INFRASTRUCTURE ONLY: this is the oracle for the two rows that sit on
 * either side of the hot path --
 *   (f2) the command-list front end: the bytes vg::clXXX really writes (CommandList::m_CommandBuffer, vg.cpp:5694-5723)
 *        and what ctxSubmitCommandList (vg.cpp:4273-4637) really does with them;
 *   (f1) the final layout: the vertex streams, uint16 index buffer and draw-command table vg::end (vg.cpp:1076-1288)
 *        hands to bgfx after createDrawCommand_VertexColor / allocVertices / allocDrawCommand (vg.cpp:5207-5405).
 * No reference source is copied: vg.cpp is #included from /root/reference (a unity build, so that the command buffer of
 * a list can be read back); path.cpp / stroker.cpp / vg_util.cpp / fontstash are linked from the reference as well.
 * Built by oracle/Makefile into oracle/_ref/libvgctx.so.
 */
#include "vg.cpp" /* /root/reference/src, via -I */

#include <vector>

namespace bgfx
{
Recorder& recorder()
{
	static Recorder r;
	return r;
}
}

namespace
{
bx::DefaultAllocator s_ctxAlloc;

struct Rd
{
	const uint8_t* p;
	template<class T> T get()
	{
		T v;
		memcpy(&v, p, sizeof(T));
		p += sizeof(T);
		return v;
	}
};
}

extern "C" {

struct vgctx_draw
{
	uint32_t vb_id, first_vertex, num_vertices, first_index, num_indices, program;
	uint16_t scissor[4];
};

/*
 * script: a command list in vg-renderer's own byte format (16-byte aligned records {type, size} + payload), restricted to
 * path commands, colour fills / strokes, state push / pop, scissor, transforms, global alpha.  Every record is replayed
 * through the PUBLIC recording API vg::clXXX into a real command list; that list's own bytes come back in clBytes (they
 * must equal the script: the format check), then the list is submitted inside begin()/end() and everything bgfx was
 * given comes back: per vertex buffer its pos / uv / colour streams (concatenated, vbCounts[i] vertices each), the
 * uint16 index buffer, and one record per bgfx::submit.
 */
/* replay the records of a script through the public recording API into list cl */
static int record(vg::Context* ctx, vg::CommandListHandle cl, const uint8_t* script, uint32_t size)
{
	int rc = 0;
	const uint8_t* cmd = script;
	const uint8_t* end = script + size;
	while (cmd < end && rc == 0) {
		Rd r = {cmd};
		const uint32_t type = r.get<uint32_t>();
		const uint32_t dataSize = r.get<uint32_t>();
		r.p = cmd + 16; /* the header {type, size} occupies one 16-byte slot (alignSize(sizeof(CommandHeader), 16), vg.cpp:708) */
		float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
		memcpy(f, r.p, dataSize < sizeof(f) ? dataSize : sizeof(f));
		switch (type) {
		case vg::CommandType::BeginPath: vg::clBeginPath(ctx, cl); break;
		case vg::CommandType::MoveTo: vg::clMoveTo(ctx, cl, f[0], f[1]); break;
		case vg::CommandType::LineTo: vg::clLineTo(ctx, cl, f[0], f[1]); break;
		case vg::CommandType::CubicTo: vg::clCubicTo(ctx, cl, f[0], f[1], f[2], f[3], f[4], f[5]); break;
		case vg::CommandType::QuadraticTo: vg::clQuadraticTo(ctx, cl, f[0], f[1], f[2], f[3]); break;
		case vg::CommandType::ArcTo: vg::clArcTo(ctx, cl, f[0], f[1], f[2], f[3], f[4]); break;
		case vg::CommandType::Arc: {
			uint32_t dir;
			memcpy(&dir, cmd + 16 + 20, 4);
			vg::clArc(ctx, cl, f[0], f[1], f[2], f[3], f[4], (vg::Winding::Enum)dir);
		} break;
		case vg::CommandType::Rect: vg::clRect(ctx, cl, f[0], f[1], f[2], f[3]); break;
		case vg::CommandType::RoundedRect: vg::clRoundedRect(ctx, cl, f[0], f[1], f[2], f[3], f[4]); break;
		case vg::CommandType::RoundedRectVarying: vg::clRoundedRectVarying(ctx, cl, f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7]); break;
		case vg::CommandType::Circle: vg::clCircle(ctx, cl, f[0], f[1], f[2]); break;
		case vg::CommandType::Ellipse: vg::clEllipse(ctx, cl, f[0], f[1], f[2], f[3]); break;
		case vg::CommandType::Polyline: {
			const uint32_t n = r.get<uint32_t>();
			std::vector<float> pts(2 * (size_t)n + 2);
			memcpy(pts.data(), r.p, sizeof(float) * 2 * n);
			vg::clPolyline(ctx, cl, pts.data(), n);
		} break;
		case vg::CommandType::ClosePath: vg::clClosePath(ctx, cl); break;
		case vg::CommandType::FillPathColor: {
			const uint32_t flags = r.get<uint32_t>();
			const uint32_t color = r.get<uint32_t>();
			vg::clFillPath(ctx, cl, (vg::Color)color, flags);
		} break;
		case vg::CommandType::StrokePathColor: {
			const float width = r.get<float>();
			const uint32_t flags = r.get<uint32_t>();
			const uint32_t color = r.get<uint32_t>();
			vg::clStrokePath(ctx, cl, (vg::Color)color, width, flags);
		} break;
		case vg::CommandType::PushState: vg::clPushState(ctx, cl); break;
		case vg::CommandType::PopState: vg::clPopState(ctx, cl); break;
		case vg::CommandType::ResetScissor: vg::clResetScissor(ctx, cl); break;
		case vg::CommandType::SetScissor: vg::clSetScissor(ctx, cl, f[0], f[1], f[2], f[3]); break;
		case vg::CommandType::IntersectScissor: vg::clIntersectScissor(ctx, cl, f[0], f[1], f[2], f[3]); break;
		case vg::CommandType::TransformIdentity: vg::clTransformIdentity(ctx, cl); break;
		case vg::CommandType::TransformScale: vg::clTransformScale(ctx, cl, f[0], f[1]); break;
		case vg::CommandType::TransformTranslate: vg::clTransformTranslate(ctx, cl, f[0], f[1]); break;
		case vg::CommandType::TransformRotate: vg::clTransformRotate(ctx, cl, f[0]); break;
		case vg::CommandType::TransformMult: {
			uint32_t order;
			memcpy(&order, cmd + 16 + 24, 4);
			vg::clTransformMult(ctx, cl, f, (vg::TransformOrder::Enum)order);
		} break;
		case vg::CommandType::SetViewBox: vg::clSetViewBox(ctx, cl, f[0], f[1], f[2], f[3]); break;
		case vg::CommandType::SetGlobalAlpha: vg::clSetGlobalAlpha(ctx, cl, f[0]); break;
		case vg::CommandType::SubmitCommandList: {
			const vg::CommandListHandle child = {r.get<uint16_t>()};
			vg::clSubmitCommandList(ctx, cl, child);
		} break;
		default: rc = -2; break; /* outside the oracle's script vocabulary */
		}
		cmd = cmd + 16 + dataSize;
	}
	return rc;
}

/* xforms (numXforms x 6, may be null): instead of one submission under the identity, the list is submitted once per row
 * inside  vg::pushState / vg::transformMult(row, Pre) / ... / vg::popState  -- with CommandListFlags::Cacheable in listFlags
 * this is the shape cache's life cycle: the first submission tessellates and caches (addCachedCommand, vg.cpp:5803-5841),
 * every later one whose average scale equals the cached one replays (clCacheRender -> submitCachedMesh, vg.cpp:6137-6166).
 * inv0 (6 floats, may be null) receives vgutil::invertMatrix3 of the first row: the matrix beginCachedCommand stores
 * (vg.cpp:5788) for a list without transforms of its own. */
int vgctx_run_instances(const uint8_t* script, uint32_t size, uint32_t listFlags, uint32_t canvasW, uint32_t canvasH, uint32_t maxVBVertices,
                        const float* xforms, uint32_t numXforms, float* inv0,
                        const uint8_t* childBytes, const uint32_t* childSizes, const uint32_t* childFlags, uint32_t numChildren,
                        uint8_t* clBytes, uint32_t clCap, uint32_t* clSize,
                        vgctx_draw* draws, uint32_t drawCap, uint32_t* numDraws,
                        float* pos, int16_t* uv, uint32_t* col, uint64_t vcap, uint32_t* vbCounts, uint32_t vbCap, uint32_t* numVBs,
                        uint16_t* idx, uint64_t icap, uint64_t* numIdx)
{
	vg::ContextConfig cfg = {64, 64, 8, 32, 16, 256, maxVBVertices ? maxVBVertices : 65536u, vg::ImageFlags::Filter_Bilinear, 16, true};
	bgfx::recorder() = bgfx::Recorder();
	bgfx::recorder().resetDraw();
	vg::Context* ctx = vg::createContext(&s_ctxAlloc, &cfg);
	if (!ctx) return -1;
	/* the lists a SubmitCommandList record may name: handles 0 .. numChildren-1 (the allocator hands them out in order),
	 * all created before any is recorded so that they can name each other -- or themselves */
	int rc = 0;
	std::vector<vg::CommandListHandle> kids;
	for (uint32_t i = 0; i < numChildren; ++i) {
		kids.push_back(vg::createCommandList(ctx, childFlags[i]));
		if (kids.back().idx != i) rc = -3;
	}
	for (uint32_t i = 0, off = 0; i < numChildren && rc == 0; off += childSizes[i], ++i) rc = record(ctx, kids[i], childBytes + off, childSizes[i]);
	vg::CommandListHandle cl = vg::createCommandList(ctx, listFlags);
	if (rc == 0) rc = record(ctx, cl, script, size);
	if (rc == 0) {
		/* the list as the reference recorded it */
		const vg::CommandList* l = &ctx->m_CmdLists[cl.idx];
		*clSize = l->m_CommandBufferPos;
		if (l->m_CommandBufferPos <= clCap) memcpy(clBytes, l->m_CommandBuffer, l->m_CommandBufferPos);
		else rc = -5;
	}
	if (rc == 0) {
		vg::begin(ctx, 0, (uint16_t)canvasW, (uint16_t)canvasH, 1.0f);
		if (xforms == nullptr) {
			vg::submitCommandList(ctx, cl);
		} else {
			if (inv0 && numXforms) vgutil::invertMatrix3(xforms, inv0);
			for (uint32_t k = 0; k < numXforms; ++k) {
				vg::pushState(ctx);
				vg::transformMult(ctx, xforms + 6 * (size_t)k, vg::TransformOrder::Pre);
				vg::submitCommandList(ctx, cl);
				vg::popState(ctx);
			}
		}
		vg::end(ctx);
		const bgfx::Recorder& R = bgfx::recorder();
		/* vertex buffers: vg.cpp creates (pos, uv, colour) handles per VertexBuffer, in that order (vg.cpp:1106-1114) */
		const uint32_t nvb = (uint32_t)(R.vbs.size() / 3);
		*numVBs = nvb;
		uint64_t v = 0;
		for (uint32_t i = 0; i < nvb && rc == 0; ++i) {
			const uint32_t count = (uint32_t)(R.vbs[3 * i].size() / 8);
			if (i >= vbCap || v + count > vcap) { rc = -5; break; }
			vbCounts[i] = count;
			memcpy(pos + 2 * v, R.vbs[3 * i].data(), (size_t)count * 8);
			memcpy(uv + 2 * v, R.vbs[3 * i + 1].data(), (size_t)count * 4);
			memcpy(col + v, R.vbs[3 * i + 2].data(), (size_t)count * 4);
			v += count;
		}
		const uint64_t ni = R.ibs.empty() ? 0 : R.ibs[0].size() / 2;
		*numIdx = ni;
		if (ni > icap) rc = -5;
		else if (ni) memcpy(idx, R.ibs[0].data(), ni * 2);
		*numDraws = (uint32_t)R.draws.size();
		if (R.draws.size() > drawCap) rc = -5;
		for (size_t i = 0; i < R.draws.size() && rc == 0; ++i) {
			const bgfx::Draw& d = R.draws[i];
			draws[i].vb_id = d.vb[0] / 3;
			draws[i].first_vertex = d.firstVertex;
			draws[i].num_vertices = d.numVertices;
			draws[i].first_index = d.firstIndex;
			draws[i].num_indices = d.numIndices;
			draws[i].program = d.program;
			memcpy(draws[i].scissor, d.scissor, 8);
		}
		vg::frame(ctx);
	}
	vg::destroyCommandList(ctx, cl);
	for (size_t i = 0; i < kids.size(); ++i) vg::destroyCommandList(ctx, kids[i]);
	vg::destroyContext(ctx);
	return rc;
}

int vgctx_run(const uint8_t* script, uint32_t size, uint32_t listFlags, uint32_t canvasW, uint32_t canvasH, uint32_t maxVBVertices,
              uint8_t* clBytes, uint32_t clCap, uint32_t* clSize,
              vgctx_draw* draws, uint32_t drawCap, uint32_t* numDraws,
              float* pos, int16_t* uv, uint32_t* col, uint64_t vcap, uint32_t* vbCounts, uint32_t vbCap, uint32_t* numVBs,
              uint16_t* idx, uint64_t icap, uint64_t* numIdx)
{
	return vgctx_run_instances(script, size, listFlags, canvasW, canvasH, maxVBVertices, nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0, clBytes, clCap, clSize, draws, drawCap,
	                           numDraws, pos, uv, col, vcap, vbCounts, vbCap, numVBs, idx, icap, numIdx);
}

} /* extern "C" */
