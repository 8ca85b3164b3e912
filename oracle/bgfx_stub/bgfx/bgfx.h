/*
 * RECORD-ONLY stand-in for <bgfx/bgfx.h>: just the names the reference's src/vg.cpp uses (vg.cpp:16-21, 785-1288),
 * so that vg.cpp itself -- the real caller of vg::pathXXX / vg::strokerXXX, with its command-list executor
 * (ctxSubmitCommandList, vg.cpp:4273-4637) and its merge step (createDrawCommand_VertexColor, vg.cpp:5207-5405) --
 * compiles and RUNS here without a graphics device.  Nothing is rendered: buffer updates are copied into a
 * process-wide Recorder and every bgfx::submit appends the draw call's state (program, vertex/index ranges, scissor),
 * which is exactly the final layout the GPU path must hand to bgfx.  TEST INFRASTRUCTURE ONLY (oracle/Makefile:
 * libvgctx.so).  bgfx is not vendored in the reference and not pinned; nothing here is copied from bgfx.
 */
#ifndef BGFX_STUB_H
#define BGFX_STUB_H

#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <vector>

#define BGFX_INVALID_HANDLE { UINT16_MAX }
#define BGFX_BUFFER_ALLOW_RESIZE UINT16_C(0x0800)
#define BGFX_TEXTURE_NONE UINT64_C(0)
#define BGFX_SAMPLER_NONE UINT32_C(0)
#define BGFX_SAMPLER_U_CLAMP UINT32_C(0x2)
#define BGFX_SAMPLER_V_CLAMP UINT32_C(0x8)
#define BGFX_SAMPLER_W_CLAMP UINT32_C(0x20)
#define BGFX_SAMPLER_MIN_POINT UINT32_C(0x40)
#define BGFX_SAMPLER_MAG_POINT UINT32_C(0x100)
#define BGFX_SAMPLER_MIP_POINT UINT32_C(0x400)
#define BGFX_STATE_WRITE_RGB UINT64_C(0x7)
#define BGFX_STATE_WRITE_A UINT64_C(0x8)
#define BGFX_STATE_BLEND_ONE UINT64_C(0x2000)
#define BGFX_STATE_BLEND_SRC_ALPHA UINT64_C(0x5000)
#define BGFX_STATE_BLEND_INV_SRC_ALPHA UINT64_C(0x6000)
#define BGFX_STATE_BLEND_FUNC_SEPARATE(_srcRGB, _dstRGB, _srcA, _dstA) \
	(UINT64_C(0) | (((uint64_t)(_srcRGB) | ((uint64_t)(_dstRGB) << 4))) | (((uint64_t)(_srcA) | ((uint64_t)(_dstA) << 4)) << 8))
#define BGFX_STENCIL_NONE UINT32_C(0)
#define BGFX_STENCIL_FUNC_REF(v) (((uint32_t)(v)) & UINT32_C(0xff))
#define BGFX_STENCIL_FUNC_RMASK(v) ((((uint32_t)(v)) << 8) & UINT32_C(0xff00))
#define BGFX_STENCIL_TEST_EQUAL UINT32_C(0x00030000)
#define BGFX_STENCIL_TEST_NOTEQUAL UINT32_C(0x00060000)
#define BGFX_STENCIL_TEST_ALWAYS UINT32_C(0x00080000)
#define BGFX_STENCIL_OP_FAIL_S_KEEP UINT32_C(0x00100000)
#define BGFX_STENCIL_OP_FAIL_S_REPLACE UINT32_C(0x00200000)
#define BGFX_STENCIL_OP_FAIL_Z_KEEP UINT32_C(0x01000000)
#define BGFX_STENCIL_OP_FAIL_Z_REPLACE UINT32_C(0x02000000)
#define BGFX_STENCIL_OP_PASS_Z_KEEP UINT32_C(0x10000000)
#define BGFX_STENCIL_OP_PASS_Z_REPLACE UINT32_C(0x20000000)

namespace bgfx
{
#define BGFX_STUB_HANDLE(_name) \
	struct _name { uint16_t idx; }; \
	inline bool isValid(_name h) { return h.idx != UINT16_MAX; }
BGFX_STUB_HANDLE(DynamicVertexBufferHandle)
BGFX_STUB_HANDLE(DynamicIndexBufferHandle)
BGFX_STUB_HANDLE(ProgramHandle)
BGFX_STUB_HANDLE(ShaderHandle)
BGFX_STUB_HANDLE(TextureHandle)
BGFX_STUB_HANDLE(UniformHandle)
#undef BGFX_STUB_HANDLE

typedef uint16_t ViewId;
typedef void (*ReleaseFn)(void* ptr, void* userData);

struct Memory
{
	uint8_t* data;
	uint32_t size;
	ReleaseFn releaseFn; /* stub extension: makeRef's callback, run once the data has been consumed */
	void* userData;
	bool owned;
};

struct RendererType { enum Enum { Noop, Count }; };
struct Attrib { enum Enum { Position, Normal, Tangent, Bitangent, Color0, Color1, Color2, Color3, Indices, Weight, TexCoord0, TexCoord1, Count }; };
struct AttribType { enum Enum { Uint8, Uint10, Int16, Half, Float, Count }; };
struct UniformType { enum Enum { Sampler, End, Vec4, Mat3, Mat4, Count }; };
struct TextureFormat { enum Enum { RGBA8, Count }; };

struct Caps
{
	bool homogeneousDepth;
	struct Limits
	{
		uint32_t maxTextureSize; /* sets the white-pixel rectangle of the font atlas, hence the uv of every path vertex (vg.cpp:836-837) */
	} limits;
};

struct VertexLayout
{
	uint16_t stride;
	VertexLayout& begin(RendererType::Enum = RendererType::Noop) { stride = 0; return *this; }
	VertexLayout& add(Attrib::Enum, uint8_t num, AttribType::Enum type, bool = false, bool = false)
	{
		static const uint16_t sz[] = {1, 4, 2, 2, 4};
		stride = (uint16_t)(stride + (type == AttribType::Uint10 ? 4 : num * sz[type]));
		return *this;
	}
	void end() {}
};

struct EmbeddedShader { const char* name; };

/* ---- what the stub records ---------------------------------------------------------------------------- */
struct Draw
{
	uint16_t program;
	uint16_t vb[3];           /* streams 0 (pos), 1 (colour), 2 (uv): dynamic vertex buffer handles (UINT16_MAX = unset) */
	uint32_t firstVertex, numVertices;
	uint16_t ib;
	uint32_t firstIndex, numIndices;
	uint16_t scissor[4];
	uint64_t state;
	uint32_t stencil;
};

struct Recorder
{
	std::vector<std::vector<uint8_t>> vbs; /* by DynamicVertexBufferHandle */
	std::vector<uint16_t> vbStride;
	std::vector<std::vector<uint8_t>> ibs; /* by DynamicIndexBufferHandle */
	std::vector<Draw> draws;
	Draw cur;
	std::vector<uint16_t> scissors; /* 4 per cached scissor id */
	uint16_t numPrograms = 0, numUniforms = 0, numTextures = 0, numShaders = 0;
	void resetDraw()
	{
		memset(&cur, 0, sizeof(cur));
		cur.vb[0] = cur.vb[1] = cur.vb[2] = UINT16_MAX;
		cur.ib = UINT16_MAX;
	}
	void frame()
	{
		draws.clear();
		scissors.clear();
		resetDraw();
	}
};
Recorder& recorder();

inline void consume(const Memory* m)
{
	if (!m) return;
	if (m->releaseFn) m->releaseFn(m->data, m->userData);
	if (m->owned) ::free(m->data);
	delete m;
}

inline const Memory* makeRef(const void* data, uint32_t size, ReleaseFn fn = nullptr, void* ud = nullptr)
{
	return new Memory{(uint8_t*)data, size, fn, ud, false};
}
inline const Memory* alloc(uint32_t size) { return new Memory{(uint8_t*)::malloc(size), size, nullptr, nullptr, true}; }
inline const Memory* copy(const void* data, uint32_t size)
{
	const Memory* m = alloc(size);
	memcpy(m->data, data, size);
	return m;
}

inline const Caps* getCaps() { static Caps c = {false, {16384}}; return &c; }
inline RendererType::Enum getRendererType() { return RendererType::Noop; }

inline DynamicVertexBufferHandle createDynamicVertexBuffer(uint32_t, const VertexLayout& layout, uint16_t = 0)
{
	Recorder& r = recorder();
	r.vbs.emplace_back();
	r.vbStride.push_back(layout.stride);
	return DynamicVertexBufferHandle{(uint16_t)(r.vbs.size() - 1)};
}
inline void update(DynamicVertexBufferHandle h, uint32_t startVertex, const Memory* m)
{
	Recorder& r = recorder();
	std::vector<uint8_t>& b = r.vbs[h.idx];
	const size_t off = (size_t)startVertex * r.vbStride[h.idx];
	if (b.size() < off + m->size) b.resize(off + m->size);
	memcpy(b.data() + off, m->data, m->size);
	b.resize(off + m->size); /* the frame's content is exactly what was uploaded */
	consume(m);
}
inline DynamicIndexBufferHandle createDynamicIndexBuffer(const Memory* m, uint16_t = 0)
{
	Recorder& r = recorder();
	r.ibs.emplace_back(m->data, m->data + m->size);
	consume(m);
	return DynamicIndexBufferHandle{(uint16_t)(r.ibs.size() - 1)};
}
inline void update(DynamicIndexBufferHandle h, uint32_t startIndex, const Memory* m)
{
	std::vector<uint8_t>& b = recorder().ibs[h.idx];
	const size_t off = (size_t)startIndex * 2;
	b.resize(off + m->size);
	memcpy(b.data() + off, m->data, m->size);
	consume(m);
}
inline void destroy(DynamicVertexBufferHandle) {}
inline void destroy(DynamicIndexBufferHandle) {}
inline void destroy(ProgramHandle) {}
inline void destroy(ShaderHandle) {}
inline void destroy(TextureHandle) {}
inline void destroy(UniformHandle) {}

inline ShaderHandle createEmbeddedShader(const EmbeddedShader*, RendererType::Enum, const char*) { return ShaderHandle{recorder().numShaders++}; }
inline ProgramHandle createProgram(ShaderHandle, ShaderHandle, bool = false) { return ProgramHandle{recorder().numPrograms++}; }
inline UniformHandle createUniform(const char*, UniformType::Enum, uint16_t = 1) { return UniformHandle{recorder().numUniforms++}; }
inline TextureHandle createTexture2D(uint16_t, uint16_t, bool, uint16_t, TextureFormat::Enum, uint64_t = 0, const Memory* m = nullptr)
{
	consume(m);
	return TextureHandle{recorder().numTextures++};
}
inline void updateTexture2D(TextureHandle, uint16_t, uint8_t, uint16_t, uint16_t, uint16_t, uint16_t, const Memory* m, uint16_t = UINT16_MAX) { consume(m); }

inline void setViewTransform(ViewId, const void*, const void*) {}
inline void setUniform(UniformHandle, const void*, uint16_t = 1) {}
inline void setTexture(uint8_t, UniformHandle, TextureHandle, uint32_t = UINT32_MAX) {}
inline void setState(uint64_t s, uint32_t = 0) { recorder().cur.state = s; }
inline void setStencil(uint32_t f, uint32_t = BGFX_STENCIL_NONE) { recorder().cur.stencil = f; }
inline void setVertexBuffer(uint8_t stream, DynamicVertexBufferHandle h, uint32_t first, uint32_t num)
{
	Draw& d = recorder().cur;
	if (stream < 3) d.vb[stream] = h.idx;
	d.firstVertex = first;
	d.numVertices = num;
}
inline void setIndexBuffer(DynamicIndexBufferHandle h, uint32_t first, uint32_t num)
{
	Draw& d = recorder().cur;
	d.ib = h.idx;
	d.firstIndex = first;
	d.numIndices = num;
}
inline uint16_t setScissor(uint16_t x, uint16_t y, uint16_t w, uint16_t h)
{
	Recorder& r = recorder();
	const uint16_t v[4] = {x, y, w, h};
	r.scissors.insert(r.scissors.end(), v, v + 4);
	memcpy(r.cur.scissor, v, sizeof(v));
	return (uint16_t)(r.scissors.size() / 4 - 1);
}
inline void setScissor(uint16_t cache = UINT16_MAX)
{
	Recorder& r = recorder();
	if (cache != UINT16_MAX && (size_t)cache * 4 + 3 < r.scissors.size()) memcpy(r.cur.scissor, &r.scissors[(size_t)cache * 4], 8);
	else memset(r.cur.scissor, 0, 8); /* no scissor */
}
inline void submit(ViewId, ProgramHandle p, uint32_t = 0, uint8_t = 0xff)
{
	Recorder& r = recorder();
	r.cur.program = p.idx;
	r.draws.push_back(r.cur);
	r.resetDraw();
}
}

#endif
