/*
 * vgb_dropin.cpp -- the namespace-vg symbols of include/vg/path.h:19-38 and
 * include/vg/stroker.h:11-70, forwarding to the C ABI (vgb_path_* / vgb_stroker_*), so that the
 * reference's src/vg.cpp (call sites vg.cpp:778-779, 1006-1009, 2978-3058, 3088-3651, 4965-4970)
 * links against libvgb200.so instead of src/path.cpp + src/stroker.cpp.
 */
#include "vgb_dropin.hpp"
#include "vgb.h"

#define VG_EXPORT __attribute__((visibility("default")))

namespace
{
void* bxRealloc(void* user, void* ptr, size_t size, size_t align)
{
	bx::AllocatorI* a = (bx::AllocatorI*)user;
	/* contents are rewritten by the caller after every resize: free + alloc keeps alignment simple */
	if (ptr) a->realloc(ptr, 0, align, nullptr, 0);
	if (size == 0) return nullptr;
	return a->realloc(nullptr, size, align, nullptr, 0);
}
inline vgb_path* P(vg::Path* p) { return (vgb_path*)p; }
inline vgb_path* P(const vg::Path* p) { return (vgb_path*)p; }
inline vgb_stroker* S(vg::Stroker* s) { return (vgb_stroker*)s; }
static_assert(sizeof(vg::SubPath) == sizeof(vgb_subpath), "SubPath layout");
static_assert(sizeof(vg::Mesh) == sizeof(vgb_mesh), "Mesh layout");
}

namespace vg
{
VG_EXPORT Path* createPath(bx::AllocatorI* allocator) { return (Path*)vgb_path_create(allocator ? bxRealloc : nullptr, allocator); }
VG_EXPORT void destroyPath(Path* path) { vgb_path_destroy(P(path)); }
VG_EXPORT void pathReset(Path* path, float scale, float tol) { vgb_path_reset(P(path), scale, tol); }
VG_EXPORT void pathMoveTo(Path* path, float x, float y) { vgb_path_move_to(P(path), x, y); }
VG_EXPORT void pathLineTo(Path* path, float x, float y) { vgb_path_line_to(P(path), x, y); }
VG_EXPORT void pathCubicTo(Path* path, float c1x, float c1y, float c2x, float c2y, float x, float y) { vgb_path_cubic_to(P(path), c1x, c1y, c2x, c2y, x, y); }
VG_EXPORT void pathQuadraticTo(Path* path, float cx, float cy, float x, float y) { vgb_path_quadratic_to(P(path), cx, cy, x, y); }
VG_EXPORT void pathArcTo(Path* path, float x1, float y1, float x2, float y2, float r) { vgb_path_arc_to(P(path), x1, y1, x2, y2, r); }
VG_EXPORT void pathRect(Path* path, float x, float y, float w, float h) { vgb_path_rect(P(path), x, y, w, h); }
VG_EXPORT void pathRoundedRect(Path* path, float x, float y, float w, float h, float r) { vgb_path_rounded_rect(P(path), x, y, w, h, r); }
VG_EXPORT void pathRoundedRectVarying(Path* path, float x, float y, float w, float h, float rtl, float rtr, float rbr, float rbl) { vgb_path_rounded_rect_varying(P(path), x, y, w, h, rtl, rtr, rbr, rbl); }
VG_EXPORT void pathCircle(Path* path, float x, float y, float r) { vgb_path_circle(P(path), x, y, r); }
VG_EXPORT void pathEllipse(Path* path, float x, float y, float rx, float ry) { vgb_path_ellipse(P(path), x, y, rx, ry); }
VG_EXPORT void pathArc(Path* path, float x, float y, float r, float a0, float a1, Winding::Enum dir) { vgb_path_arc(P(path), x, y, r, a0, a1, (uint32_t)dir); }
VG_EXPORT void pathPolyline(Path* path, const float* coords, uint32_t numPoints) { vgb_path_polyline(P(path), coords, numPoints); }
VG_EXPORT void pathClose(Path* path) { vgb_path_close(P(path)); }
VG_EXPORT const float* pathGetVertices(const Path* path) { return vgb_path_get_vertices(P(path)); }
VG_EXPORT uint32_t pathGetNumVertices(const Path* path) { return vgb_path_get_num_vertices(P(path)); }
VG_EXPORT const SubPath* pathGetSubPaths(const Path* path) { return (const SubPath*)vgb_path_get_subpaths(P(path)); }
VG_EXPORT uint32_t pathGetNumSubPaths(const Path* path) { return vgb_path_get_num_subpaths(P(path)); }

VG_EXPORT Stroker* createStroker(bx::AllocatorI* allocator) { return (Stroker*)vgb_stroker_create(allocator ? bxRealloc : nullptr, allocator); }
VG_EXPORT void destroyStroker(Stroker* stroker) { vgb_stroker_destroy(S(stroker)); }
VG_EXPORT void strokerReset(Stroker* stroker, float scale, float tol, float fringe) { vgb_stroker_reset(S(stroker), scale, tol, fringe); }
VG_EXPORT void strokerPolylineStroke(Stroker* stroker, Mesh* mesh, const float* v, uint32_t n, bool closed, float width, LineCap::Enum cap, LineJoin::Enum join)
{
	vgb_stroker_polyline_stroke(S(stroker), (vgb_mesh*)mesh, v, n, closed, width, (uint32_t)cap, (uint32_t)join);
}
VG_EXPORT void strokerPolylineStrokeAA(Stroker* stroker, Mesh* mesh, const float* v, uint32_t n, bool closed, Color color, float width, LineCap::Enum cap, LineJoin::Enum join)
{
	vgb_stroker_polyline_stroke_aa(S(stroker), (vgb_mesh*)mesh, v, n, closed, color, width, (uint32_t)cap, (uint32_t)join);
}
VG_EXPORT void strokerPolylineStrokeAAThin(Stroker* stroker, Mesh* mesh, const float* v, uint32_t n, bool closed, Color color, LineCap::Enum cap, LineJoin::Enum join)
{
	vgb_stroker_polyline_stroke_aa_thin(S(stroker), (vgb_mesh*)mesh, v, n, closed, color, (uint32_t)cap, (uint32_t)join);
}
VG_EXPORT void strokerConvexFill(Stroker* stroker, Mesh* mesh, const float* v, uint32_t n) { vgb_stroker_convex_fill(S(stroker), (vgb_mesh*)mesh, v, n); }
VG_EXPORT void strokerConvexFillAA(Stroker* stroker, Mesh* mesh, const float* v, uint32_t n, uint32_t color) { vgb_stroker_convex_fill_aa(S(stroker), (vgb_mesh*)mesh, v, n, color); }

/* Concave fills (include/vg/stroker.h:73-85): not tessellated here (libtess2 sweep line: serial, out of
 * scope).  The symbols exist so that src/vg.cpp links; End/EndAA report the reference's own failure
 * result (false, empty mesh), on which vg.cpp:3158-3159 (and its siblings near 3266 and 3381) skip the draw. */
namespace
{
inline bool emptyMesh(Mesh* mesh)
{
	if (mesh) {
		mesh->m_PosBuffer = nullptr;
		mesh->m_ColorBuffer = nullptr;
		mesh->m_IndexBuffer = nullptr;
		mesh->m_NumVertices = 0;
		mesh->m_NumIndices = 0;
	}
	return false;
}
}
VG_EXPORT bool strokerConcaveFillBegin(Stroker*) { return false; }
VG_EXPORT void strokerConcaveFillAddContour(Stroker*, const float*, uint32_t) {}
VG_EXPORT bool strokerConcaveFillEnd(Stroker*, Mesh* mesh, FillRule::Enum) { return emptyMesh(mesh); }
VG_EXPORT bool strokerConcaveFillEndAA(Stroker*, Mesh* mesh, uint32_t, FillRule::Enum) { return emptyMesh(mesh); }
}
