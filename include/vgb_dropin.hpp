/*
 * vgb_dropin.hpp -- C++ declarations of the namespace-vg entry points libvgb200.so exports so
 * that the reference's src/vg.cpp links against it unchanged.  Signatures,
 * enum values and POD layouts are those of include/vg/path.h:11-38, include/vg/stroker.h:11-70
 * and include/vg/vg.h:156-175,252-259,353-360 of the reference; the declarations are written
 * from that interface description, not copied.  When building against the real reference tree,
 * include the reference's own <vg/path.h> / <vg/stroker.h> instead -- the symbols are the same
 * (tests/test_abi.py compares the mangled names with the compiled reference's).
 *
 * Concave fills (include/vg/stroker.h:73-85; libtess2 sweep line, inherently serial, out of scope --
 * SURVEY.md 2 row 5) are exported as well so that src/vg.cpp LINKS unchanged (its call sites:
 * vg.cpp:3133-3155, 3246-3264, 3357-3379): Begin/AddContour record nothing and End/EndAA return false
 * with an empty mesh, which is the reference's own "decomposition failed" result -- vg.cpp:3158-3159,
 * 3266, 3381 then skip the draw.  A build that needs concave fills keeps src/stroker.cpp's concave
 * half + libtess2 on the CPU.
 */
#ifndef VGB_DROPIN_HPP
#define VGB_DROPIN_HPP

#include <stddef.h>
#include <stdint.h>

#ifndef BX_ALLOCATOR_H_HEADER_GUARD
namespace bx
{
/* the one-virtual allocator interface of bkaradzic/bx (what createPath/createStroker receive) */
struct AllocatorI
{
	virtual ~AllocatorI() = 0;
	virtual void* realloc(void* ptr, size_t size, size_t align, const char* file, uint32_t line) = 0;
};
}
#endif

namespace vg
{
#ifndef VG_H
typedef uint32_t Color;
struct LineCap { enum Enum : uint32_t { Butt = 0, Round = 1, Square = 2 }; };
struct LineJoin { enum Enum : uint32_t { Miter = 0, Round = 1, Bevel = 2 }; };
struct Winding { enum Enum : uint32_t { CCW = 0, CW = 1 }; };
struct FillRule { enum Enum : uint32_t { NonZero = 0, EvenOdd = 1 }; };
struct Mesh
{
	const float* m_PosBuffer;
	const uint32_t* m_ColorBuffer;
	const uint16_t* m_IndexBuffer;
	uint32_t m_NumVertices;
	uint32_t m_NumIndices;
};
#endif

#ifndef VG_PATH_H
struct Path;
struct SubPath
{
	uint32_t m_FirstVertexID;
	uint32_t m_NumVertices;
	bool m_IsClosed;
};
Path* createPath(bx::AllocatorI* allocator);
void destroyPath(Path* path);
void pathReset(Path* path, float scale, float tesselationTolerance);
void pathMoveTo(Path* path, float x, float y);
void pathLineTo(Path* path, float x, float y);
void pathCubicTo(Path* path, float c1x, float c1y, float c2x, float c2y, float x, float y);
void pathQuadraticTo(Path* path, float cx, float cy, float x, float y);
void pathArcTo(Path* path, float x1, float y1, float x2, float y2, float r);
void pathRect(Path* path, float x, float y, float w, float h);
void pathRoundedRect(Path* path, float x, float y, float w, float h, float r);
void pathRoundedRectVarying(Path* path, float x, float y, float w, float h, float rtl, float rtr, float rbr, float rbl);
void pathCircle(Path* path, float x, float y, float r);
void pathEllipse(Path* path, float x, float y, float rx, float ry);
void pathArc(Path* path, float x, float y, float r, float a0, float a1, Winding::Enum dir);
void pathPolyline(Path* path, const float* coords, uint32_t numPoints);
void pathClose(Path* path);
const float* pathGetVertices(const Path* path);
uint32_t pathGetNumVertices(const Path* path);
const SubPath* pathGetSubPaths(const Path* path);
uint32_t pathGetNumSubPaths(const Path* path);
#endif

#ifndef VG_STROKER_H
struct Stroker;
Stroker* createStroker(bx::AllocatorI* allocator);
void destroyStroker(Stroker* stroker);
void strokerReset(Stroker* stroker, float scale, float tesselationTolerance, float fringeWidth);
void strokerPolylineStroke(Stroker* stroker, Mesh* mesh, const float* vertexList, uint32_t numVertices, bool isClosed, float strokeWidth, LineCap::Enum lineCap, LineJoin::Enum lineJoin);
void strokerPolylineStrokeAA(Stroker* stroker, Mesh* mesh, const float* vertexList, uint32_t numVertices, bool isClosed, Color color, float strokeWidth, LineCap::Enum lineCap, LineJoin::Enum lineJoin);
void strokerPolylineStrokeAAThin(Stroker* stroker, Mesh* mesh, const float* vertexList, uint32_t numVertices, bool isClosed, Color color, LineCap::Enum lineCap, LineJoin::Enum lineJoin);
void strokerConvexFill(Stroker* stroker, Mesh* mesh, const float* vertexList, uint32_t numVertices);
void strokerConvexFillAA(Stroker* stroker, Mesh* mesh, const float* vertexList, uint32_t numVertices, uint32_t color);
bool strokerConcaveFillBegin(Stroker* stroker);
void strokerConcaveFillAddContour(Stroker* stroker, const float* vertexList, uint32_t numVertices);
bool strokerConcaveFillEnd(Stroker* stroker, Mesh* mesh, FillRule::Enum fillRule);
bool strokerConcaveFillEndAA(Stroker* stroker, Mesh* mesh, uint32_t color, FillRule::Enum fillRule);
#endif
}

#endif
