/*
 * vgb.h -- C ABI of libvgb200: B200-native batched path flattening + stroke / convex-fill
 * mesh generation (hand-written sm_100a CUDA underneath).
 *
 * This is the drop-in boundary for the ONE hot path of jdryg/vg-renderer:
 *
 *   reference interface                                   replaced by
 *   ---------------------------------------------------   -----------------------------------
 *   include/vg/path.h:19-38    vg::createPath ...          vgb_path_*      (per-object, 1:1)
 *   include/vg/stroker.h:11-70 vg::createStroker ...       vgb_stroker_*   (per-object, 1:1)
 *   src/vg.cpp:2969-3059 ctxBeginPath/ctxMoveTo/...        vgb_batch_begin_path / vgb_batch_*_to
 *   src/vg.cpp:3401-3492 ctxStrokePathColor                vgb_batch_stroke
 *   src/vg.cpp:3061-3179 ctxFillPathColor (convex)         vgb_batch_convex_fill
 *   src/vg.cpp:4957-4975 transformPath                     xform6 argument of begin_path
 *   src/vg.cpp:5207-5244 createDrawCommand_VertexColor     VGB_RUN_INDEX_U32 merged buffers /
 *                                                          vgb_batch_download_final
 *
 * Plain C types only: pointers, sizes, PODs.  No torch / CUDA types in any signature
 * (device pointers are returned as void*).  All functions return 0 on success, a negative
 * VGB_ERR_* otherwise; vgb_last_error() gives the message for the calling thread.
 * A vgb_batch is not re-entrant; distinct batches are independent (one CUDA stream each).
 *
 * The C++ namespace-vg wrappers that let src/vg.cpp link unchanged are in
 * include/vgb_dropin.hpp; the reference-side binding is shown in INTEGRATION.md.
 */
#ifndef VGB_H
#define VGB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VGB_API __attribute__((visibility("default")))

/* ---- error codes -------------------------------------------------------------------- */
enum {
	VGB_OK = 0,
	VGB_ERR_INVALID = -1,   /* bad argument / command sequence (e.g. lineTo before moveTo) */
	VGB_ERR_NONFINITE = -2, /* NaN/Inf coordinate: the reference never terminates on these */
	VGB_ERR_CUDA = -3,      /* CUDA runtime failure (message has the cudaError string)      */
	VGB_ERR_NOGPU = -4,     /* no usable CUDA device: there is NO CPU fallback              */
	VGB_ERR_CAPACITY = -5,  /* caller buffer too small                                      */
	VGB_ERR_STATE = -6      /* call out of order (e.g. download before run)                 */
};

/* ---- enums: numeric values identical to the reference (include/vg/vg.h:156-259) ------ */
enum { VGB_CAP_BUTT = 0, VGB_CAP_ROUND = 1, VGB_CAP_SQUARE = 2 };  /* vg::LineCap  */
enum { VGB_JOIN_MITER = 0, VGB_JOIN_ROUND = 1, VGB_JOIN_BEVEL = 2 }; /* vg::LineJoin */
enum { VGB_WINDING_CCW = 0, VGB_WINDING_CW = 1 };                  /* vg::Winding  */

/* path commands (one per vg::pathXXX builder, include/vg/path.h:23-36); floats consumed */
enum {
	VGB_CMD_MOVE_TO = 0,              /* x y                         */
	VGB_CMD_LINE_TO = 1,              /* x y                         */
	VGB_CMD_CUBIC_TO = 2,             /* c1x c1y c2x c2y x y         */
	VGB_CMD_QUAD_TO = 3,              /* cx cy x y                   */
	VGB_CMD_ARC_TO = 4,               /* x1 y1 x2 y2 r               */
	VGB_CMD_RECT = 5,                 /* x y w h                     */
	VGB_CMD_ROUNDED_RECT = 6,         /* x y w h r                   */
	VGB_CMD_ROUNDED_RECT_VARYING = 7, /* x y w h rtl rtr rbr rbl     */
	VGB_CMD_CIRCLE = 8,               /* cx cy r                     */
	VGB_CMD_ELLIPSE = 9,              /* cx cy rx ry                 */
	VGB_CMD_ARC = 10,                 /* cx cy r a0 a1 dir(0|1)      */
	VGB_CMD_POLYLINE = 11,            /* n(as uint32 bits) then 2n   */
	VGB_CMD_CLOSE = 12,               /* -                           */
	VGB_CMD__COUNT = 13
};

/* mesh operations (one per vg::strokerXXX entry point, include/vg/stroker.h:29-70) */
enum {
	VGB_OP_STROKE_AA = 0,      /* strokerPolylineStrokeAA     pos + color + index */
	VGB_OP_CONVEX_FILL_AA = 1, /* strokerConvexFillAA         pos + color + index */
	VGB_OP_STROKE = 2,         /* strokerPolylineStroke       pos + index (color = draw color) */
	VGB_OP_STROKE_AA_THIN = 3, /* strokerPolylineStrokeAAThin pos + color + index */
	VGB_OP_CONVEX_FILL = 4     /* strokerConvexFill           pos(copy) + index   */
};

/* vgb_batch_run flags */
enum {
	VGB_RUN_INDEX_U16 = 0, /* per-mesh local uint16 indices: exactly vg::Mesh::m_IndexBuffer */
	VGB_RUN_INDEX_U32 = 1, /* merged uint32 indices: local + mesh first vertex + vertex_base  */
	/* strokerConvexFillAA index ORDER of the reference's stock x86 build (VG_CONFIG_ENABLE_SIMD=1, stroker.cpp:616-701):
	 * Q0, F0,Q1, F1,Q2, ... F(N-3),Q(N-2), Q(N-1) instead of the scalar build's [all fan triangles][all fringe quads]
	 * (stroker.cpp:768-796).  Same triangles, same vertices (values stay those of the scalar build: the SSE
	 * build's approximate rsqrt/rcp are hardware dependent and not a parity target). */
	VGB_RUN_FILL_ORDER_SSE = 2
};

/* Final layout: what vg.cpp's merge step createDrawCommand_VertexColor
 * (vg.cpp:5207-5244) builds out of the meshes one by one, produced by the batch in one go:
 *   - pos / color streams as they are (a vertex buffer of the context is a contiguous slice),
 *   - a uv stream holding the white-pixel uv for every vertex (vg.cpp:5220-5227, 4907-4919),
 *   - uint16 indices rebased by the number of vertices already in the mesh's draw command
 *     (vgutil::batchTransformDrawIndices, vg_util.cpp:447-520, called at vg.cpp:5240),
 *   - the draw-command table: a new vertex buffer and a new draw command start whenever the
 *     next mesh would take the current buffer past max_vb_vertices (allocVertices,
 *     vg.cpp:5321-5342; allocDrawCommand, vg.cpp:5359-5405).
 * Set with vgb_batch_set_final_layout before vgb_batch_run(VGB_RUN_INDEX_U16). */
typedef struct vgb_final_cfg {
	uint32_t max_vb_vertices; /* ContextConfig::m_MaxVBVertices, 2..65536                               */
	uint32_t vb_id;           /* id of the context's current vertex buffer (m_NumVertexBuffers - 1)       */
	uint32_t vb_count;        /* vertices already in it (VertexBuffer::m_Count)                           */
	uint32_t cmd_vertices;    /* vertices already in the open draw command this batch continues
	                             (DrawCommand::m_NumVertices; 0 = the batch opens a new command)          */
	uint32_t ib_count;        /* indices already in the active index buffer (IndexBuffer::m_Count)        */
	int16_t uv[2];            /* getWhitePixelUV(), VG_CONFIG_UV_INT16 layout                             */
} vgb_final_cfg;

/* the DrawCommand fields this path decides (vg.cpp:125-129) + where the command's data sits
 * in the batch's streams */
typedef struct vgb_draw_cmd {
	uint32_t vb_id;        /* m_VertexBufferID                                              */
	uint32_t first_vertex; /* m_FirstVertexID: where the batch's vertices start in that buffer */
	uint32_t first_index;  /* m_FirstIndexID  (ib_count + indices of the earlier commands)   */
	uint32_t num_vertices; /* vertices the batch adds to the command                        */
	uint32_t num_indices;
	uint32_t first_mesh;   /* meshes [first_mesh, first_mesh + num_meshes) of the batch       */
	uint32_t num_meshes;
	uint32_t scissor_id;   /* m_ScissorRect: index into the batch's scissor table (vgb_batch_scissor_rects); 0 when no
	                        * scissor was ever set */
	uint64_t stream_vertex; /* offset of the command's first vertex in the pos/uv/color streams */
	uint64_t stream_index;  /* offset of its first index in the index stream                  */
} vgb_draw_cmd;

typedef struct vgb_batch vgb_batch;

/* vg::Mesh (include/vg/vg.h:353-360), C layout */
typedef struct vgb_mesh {
	const float* pos;      /* x,y interleaved                        */
	const uint32_t* color; /* NULL when the op has no per-vertex colour */
	const uint16_t* index;
	uint32_t num_vertices;
	uint32_t num_indices;
} vgb_mesh;

/* vg::SubPath (include/vg/path.h:11-16) */
typedef struct vgb_subpath {
	uint32_t first_vertex;
	uint32_t num_vertices;
	uint8_t is_closed;
} vgb_subpath;

/* result sizes of one run */
typedef struct vgb_sizes {
	uint64_t num_vertices;   /* emitted mesh vertices (the metric's unit)  */
	uint64_t num_indices;
	uint64_t num_points;     /* flattened polyline points                   */
	uint32_t num_subpaths;
	uint32_t num_meshes;     /* = jobs: (draw, sub-path) pairs, some empty  */
	uint32_t kernel_launches;/* kernels launched by the last run            */
	uint32_t reserved;
} vgb_sizes;

/* ---- library ------------------------------------------------------------------------- */
VGB_API const char* vgb_last_error(void);
VGB_API int vgb_device_count(void);
VGB_API const char* vgb_version(void);

/* ---- batched path: record -> run -> read ------------------------------------------------ */
/* device_ordinal >= 0: a CUDA device (fails with VGB_ERR_NOGPU if none is usable).
 * device_ordinal == -1: record-only batch (host tables can be built and inspected; upload/run
 * fail with VGB_ERR_NOGPU) -- used by the CPU-only host-logic tests. */
VGB_API vgb_batch* vgb_batch_create(int device_ordinal);
VGB_API void vgb_batch_destroy(vgb_batch* b);

/* forget everything recorded; fringe_width = Context::m_FringeWidth (src/vg.cpp:1041) */
VGB_API int vgb_batch_reset(vgb_batch* b, float fringe_width);

/* ctxBeginPath (src/vg.cpp:2969-2981): pathReset(scale, tol) + strokerReset(scale, tol, fringe).
 * xform6 = state transform applied between flatten and stroke (transformPath), NULL = identity */
VGB_API int vgb_batch_begin_path(vgb_batch* b, float scale, float tess_tol, const float* xform6);
VGB_API int vgb_batch_move_to(vgb_batch* b, float x, float y);
VGB_API int vgb_batch_line_to(vgb_batch* b, float x, float y);
VGB_API int vgb_batch_cubic_to(vgb_batch* b, float c1x, float c1y, float c2x, float c2y, float x, float y);
VGB_API int vgb_batch_quadratic_to(vgb_batch* b, float cx, float cy, float x, float y);
VGB_API int vgb_batch_arc_to(vgb_batch* b, float x1, float y1, float x2, float y2, float r);
VGB_API int vgb_batch_rect(vgb_batch* b, float x, float y, float w, float h);
VGB_API int vgb_batch_rounded_rect(vgb_batch* b, float x, float y, float w, float h, float r);
VGB_API int vgb_batch_rounded_rect_varying(vgb_batch* b, float x, float y, float w, float h, float rtl, float rtr, float rbr, float rbl);
VGB_API int vgb_batch_circle(vgb_batch* b, float x, float y, float r);
VGB_API int vgb_batch_ellipse(vgb_batch* b, float x, float y, float rx, float ry);
/* |a0|, |a1| <= 1e6 rad (VGB_ERR_INVALID beyond: pathArc's angle normalisation, path.cpp:635-645, stops
 * making progress once 2*pi falls below the float spacing, i.e. the reference never returns) */
VGB_API int vgb_batch_arc(vgb_batch* b, float x, float y, float r, float a0, float a1, uint32_t dir);
VGB_API int vgb_batch_polyline(vgb_batch* b, const float* coords, uint32_t num_points);
VGB_API int vgb_batch_close(vgb_batch* b);

/* draw the current path: one mesh per sub-path, in sub-path order (vg.cpp:3448-3485, 3099-3132).
 * stroke: op in {VGB_OP_STROKE_AA, VGB_OP_STROKE, VGB_OP_STROKE_AA_THIN}; width is the final
 * (already scaled/clamped) strokerXXX width.  fill: aa != 0 -> strokerConvexFillAA. */
/* Scissor state for the draws recorded from now on (vg.cpp keeps it per State and stamps it on every DrawCommand,
 * uint16-truncated: vg.cpp:5392-5395).  rect = {x, y, w, h} as State::m_ScissorRect holds it; force_new_command != 0
 * reproduces m_ForceNewDrawCommand (vg.cpp sets it on every Set/Intersect/ResetScissor, and on PopState when the
 * restored rect differs from the last draw command's): the next non-empty mesh opens a new draw command in the final
 * layout (vgb_batch_set_final_layout) and vgb_draw_cmd::scissor_id names its rect.  Batches that never call this
 * have one scissor (id 0 = {0,0,0,0}, "whatever the context has") and behave as before. */
VGB_API int vgb_batch_set_scissor(vgb_batch* b, const float rect[4], int force_new_command);
/* the batch's scissor table: n rects of 4 uint16 {x, y, w, h}; id 0 is the default entry */
VGB_API int vgb_batch_scissor_rects(vgb_batch* b, uint16_t* rects4, uint32_t cap, uint32_t* n);
/* per recorded draw: scissor id | force-new-command flag << 31 (sizes from vgb_batch_recorded_counts) */
VGB_API int vgb_batch_recorded_draw_scissor(vgb_batch* b, uint32_t* out);

VGB_API int vgb_batch_stroke(vgb_batch* b, uint32_t op, uint32_t color, float width, uint32_t cap, uint32_t join);
VGB_API int vgb_batch_convex_fill(vgb_batch* b, uint32_t color, int aa);

/* bulk recorders for regular workloads (same result as the per-call sequence, one memcpy):
 * n paths of  moveTo(p[0..1]) cubicTo(p[2..7])  + stroke(op, ...)                              */
VGB_API int vgb_batch_add_cubic_strokes(vgb_batch* b, const float* p8, uint32_t n, float scale, float tess_tol,
                                        uint32_t op, uint32_t color, float width, uint32_t cap, uint32_t join);
/* n paths of  roundedRect(x,y,w,h,r) + convex fill                                             */
VGB_API int vgb_batch_add_rounded_rect_fills(vgb_batch* b, const float* xywhr, uint32_t n, float scale, float tess_tol,
                                             uint32_t color, int aa);

/* ingest whole tables (what a command-list front-end would hand over): path p =
 * commands cmd[path_cmd_off[p] .. path_cmd_off[p+1]) with floats from data[path_data_off[p] ..],
 * flatten parameters path_scale/tol[p], optional transform path_xform[6p..] (NULL = identity), and the
 * draws draw_*[draw_begin[p] .. draw_begin[p+1]) issued on it in order.  Same result and validation as
 * the per-call recording (vgb_batch_begin_path / *_to / stroke / convex_fill). */
VGB_API int vgb_batch_add_paths(vgb_batch* b, uint32_t num_paths, const uint32_t* path_cmd_off, const uint32_t* path_data_off,
                                const uint8_t* cmd, const float* data, const float* path_scale, const float* path_tol,
                                const float* path_xform, const uint32_t* draw_begin, const uint32_t* draw_op,
                                const uint32_t* draw_color, const float* draw_width, const uint8_t* draw_cap, const uint8_t* draw_join);

/* zero-copy variant for the host-buffer path: REPLACES everything recorded by the n paths of
 * vgb_batch_add_cubic_strokes, but p8 is only borrowed -- it is read (DMA, straight from the
 * caller's memory: make it pinned) by the next vgb_batch_upload / _run / _execute, where the command,
 * sub-path and job tables are generated on the device; keep it valid and unchanged until the next vgb_batch_sync (or any
 * download) returns: vgb_batch_upload only ENQUEUES the copy on the batch's stream.
 * Non-finite coordinates are detected on the device: the run fails with VGB_ERR_NONFINITE. */
VGB_API int vgb_batch_set_cubic_strokes(vgb_batch* b, const float* p8, uint32_t n, float scale, float tess_tol,
                                        uint32_t op, uint32_t color, float width, uint32_t cap, uint32_t join);

/* stroker-only entry (strokerXXX called on caller polylines, no Path object involved):
 * replaces the flatten stage's output by these polylines; draws then refer to "paths" that are
 * single sub-paths.  pts = x,y interleaved; sub-path s = pts[first[s] .. first[s]+num[s]).      */
VGB_API int vgb_batch_set_polylines(vgb_batch* b, const float* pts, uint32_t num_points,
                                    const uint32_t* first, const uint32_t* num, const uint8_t* closed,
                                    uint32_t num_subpaths, float scale, float tess_tol);
/* one draw per polyline sub-path s (after vgb_batch_set_polylines) */
VGB_API int vgb_batch_draw_polyline(vgb_batch* b, uint32_t subpath, uint32_t op, uint32_t color, float width, uint32_t cap, uint32_t join);
/* the same draw for every polyline sub-path / every recorded path */
VGB_API int vgb_batch_draw_all(vgb_batch* b, uint32_t op, uint32_t color, float width, uint32_t cap, uint32_t join);

/* transform of the current path, settable until its first draw is recorded (later calls are
 * no-ops): vg.cpp applies the state transform lazily at a path's first fill/stroke and reuses the
 * result for its other draws (transformPath, vg.cpp:4957-4975) */
VGB_API int vgb_batch_set_path_xform(vgb_batch* b, const float* xform6);
VGB_API float vgb_batch_fringe(vgb_batch* b);

/* ---- command-list front end ------------------------------------------------
 * Decodes a vg-renderer CommandList byte stream -- CommandHeader{type,size} records padded to 16
 * bytes as clAllocCommand writes them (vg.cpp:5694-5723), command ids CommandType::Enum
 * (vg.cpp:177-247) -- and records it into the batch the way ctxSubmitCommandList executes it
 * (vg.cpp:4273-4637): path commands, FillPathColor (convex), StrokePathColor, the state stack,
 * transforms, global alpha and scissor culling.  Paints, concave fills, text, clip regions,
 * and tri-lists are skipped and counted in stats->unsupported; nested lists are executed when the caller
 * provides the context's list table (vgb_cl_context::lists).  The batch's fringe
 * (vgb_batch_reset) is Context::m_FringeWidth. */
typedef struct vgb_cl_context {
	float xform[6];       /* in: State::m_TransformMtx on entry; out: on exit (the list's state changes persist, vg.h:34-35) */
	float scissor[4];     /* in/out: State::m_ScissorRect (x, y, w, h)                                */
	float global_alpha;   /* in/out: State::m_GlobalAlpha                                             */
	float tess_tolerance; /* Context::m_TesselationTolerance                                          */
	float canvas_width, canvas_height;
	uint32_t flags;       /* CommandListFlags of THIS list: 2 = AllowCommandCulling (vg.h:380-388)   */
	/* nested lists (SubmitCommandList, vg.cpp:4620-4626): the context's command lists by handle index;
	 * a handle outside the table (or an entry with commands == NULL) is an invalid handle and is skipped
	 * as the reference skips it; max_depth = ContextConfig::m_MaxCommandListDepth (0 = the default 16) */
	uint32_t num_lists;
	const struct vgb_cl_list* lists;
	uint32_t max_depth;
	uint32_t reserved;
} vgb_cl_context;

typedef struct vgb_cl_list {
	const void* commands; /* CommandList::m_CommandBuffer    */
	uint32_t size;        /* CommandList::m_CommandBufferPos */
	uint32_t flags;       /* CommandList::m_Flags            */
} vgb_cl_list;

typedef struct vgb_cl_stats {
	uint32_t commands;    /* records decoded                                         */
	uint32_t paths;       /* BeginPath                                               */
	uint32_t draws;       /* fills / strokes recorded into the batch                 */
	uint32_t culled;      /* stroker commands skipped under a zero-sized scissor, or belonging to a
	                       * path begun under a collapsed transform (average scale 0: invisible)   */
	uint32_t zero_alpha;  /* draws dropped because their colour ends up transparent  */
	uint32_t unsupported; /* commands outside this path (see above)                  */
	uint32_t state_depth; /* states left pushed when the list ended                  */
	uint32_t nested;      /* SubmitCommandList commands executed                     */
	uint32_t scissor_changes; /* recorded draws before which vg.cpp forces a new DrawCommand because of the scissor
	                       * (m_ForceNewDrawCommand: every Set/Intersect/ResetScissor, and PopState when the restored rect
	                       * differs from the last draw's; vg.cpp:3946-4033, 5359-5405) -- or before which this decoder
	                       * skipped a draw it does not handle (paints, text, clips, concave fills: the caller's own
	                       * geometry belongs between the two draw commands).  The batch carries the marks
	                       * (vgb_batch_set_scissor): the final layout starts a new draw command there */
} vgb_cl_stats;

VGB_API int vgb_batch_submit_command_list(vgb_batch* b, const void* commands, uint32_t size, vgb_cl_context* ctx, vgb_cl_stats* stats);

/* introspection of what was recorded: out6 = {paths, sub-paths, commands, data floats, draws, jobs};
 * the arrays (each nullable) receive one entry per sub-path / per job */
VGB_API int vgb_batch_recorded_counts(vgb_batch* b, uint32_t* out6);
VGB_API int vgb_batch_recorded_subpaths(vgb_batch* b, uint32_t* cmd_begin, uint32_t* cmd_end, uint32_t* data_begin, uint32_t* path_id);
VGB_API int vgb_batch_recorded_jobs(vgb_batch* b, uint32_t* job_subpath, uint32_t* job_draw);

/* the recorded streams themselves (each nullable; sizes from vgb_batch_recorded_counts): command
 * bytes, data floats, per path {scale, tol, 2x3 transform}, per draw {op, colour, width, cap, join, path} */
VGB_API int vgb_batch_recorded_streams(vgb_batch* b, uint8_t* cmd, float* data, float* path_scale, float* path_tol, float* path_xform6,
                                       uint32_t* draw_op, uint32_t* draw_color, float* draw_width, uint8_t* draw_cap, uint8_t* draw_join,
                                       uint32_t* draw_path);

/* ---- execution ----------------------------------------------------------------------------
 * upload: host tables -> device (async on the batch stream).  run: all kernels; results stay in
 * HBM.  vertex_base / index_base are added to merged uint32 indices / reported offsets (multi-GPU
 * merged buffer: the exclusive scan of per-rank totals).  */
VGB_API int vgb_batch_upload(vgb_batch* b);
VGB_API int vgb_batch_run(vgb_batch* b, uint32_t flags, uint64_t vertex_base);

/* ---- several GPUs, one merged vertex / index space --------------------------------
 * One process per GPU; ranks take disjoint path ranges and run independently.  The reference builds ONE
 * vertex/index space per frame by rebasing each mesh's indices by the vertices that precede it
 * (createDrawCommand_VertexColor, vg.cpp:5207-5244); across GPUs "what precedes" is the exclusive scan of
 * the ranks' vertex totals.  A vgb_group is that communicator (NCCL over NVLink/NVSwitch, resolved at run
 * time: libnccl.so.2).  vgb_group_unique_id is called by ONE rank; the 128 bytes reach the others by
 * whatever launcher the host uses (torchrun store, MPI, a file) and every rank calls vgb_group_create.
 *
 * vgb_batch_run_merged: flatten + count, then an all-gather of the ranks' (vertex count, index count)
 * -- 16 bytes per rank, enqueued on the batch's stream -- whose exclusive prefix, computed on the device,
 * is added to this rank's merged uint32 indices (flags must carry VGB_RUN_INDEX_U32).  Asynchronous like
 * vgb_batch_run: nothing returns to the host between count and emit.  vgb_batch_merged_info reads the
 * result back: {vertex_base, index_base, total_vertices, total_indices} (synchronises).
 *
 * vgb_group_gather: payload placement -- every rank's pos / colour / index arrays are written into ONE
 * buffer on `owner` at the rank's base (ncclSend/ncclRecv: peer copies over NVLink).  On the owner
 * *pos/*color/*index are device pointers to the merged arrays (owned by the group, valid until the next
 * gather or vgb_group_destroy), on the others they are set to NULL; ms = device time of the transfer.
 * NVLink-bound (the owner's ingress), reported separately from the HBM roofline. */
typedef struct vgb_group vgb_group;
#define VGB_GROUP_ID_BYTES 128
typedef struct vgb_merged
{
	uint64_t vertex_base, index_base, total_vertices, total_indices;
} vgb_merged;
VGB_API int vgb_group_unique_id(void* id128);
VGB_API int vgb_group_create(vgb_group** out, int device, int rank, int nranks, const void* id128);
VGB_API void vgb_group_destroy(vgb_group* g);
VGB_API int vgb_batch_run_merged(vgb_batch* b, vgb_group* g, uint32_t flags);
VGB_API int vgb_batch_merged_info(vgb_batch* b, vgb_merged* out);
VGB_API int vgb_group_gather(vgb_group* g, vgb_batch* b, int owner, void** pos, void** color, void** index, vgb_merged* info, float* ms);
VGB_API int vgb_batch_sync(vgb_batch* b);
VGB_API int vgb_batch_get_sizes(vgb_batch* b, vgb_sizes* out);

/* final layout: cfg is copied; NULL switches back to per-mesh local indices.  Applies to the next
 * runs with VGB_RUN_INDEX_U16 (VGB_ERR_INVALID with uint32 indices or a bad max_vb_vertices). */
VGB_API int vgb_batch_set_final_layout(vgb_batch* b, const vgb_final_cfg* cfg);
/* after a final-layout run: the uv stream (2 int16 per vertex; nullable), the draw commands
 * (nullable) and their number.  VGB_ERR_CAPACITY if a buffer is too small or if one mesh alone has
 * max_vb_vertices vertices or more (the reference asserts on that, vg.cpp:5323).  uv_device
 * (nullable) receives the device pointer of the uv stream. */
VGB_API int vgb_batch_download_final(vgb_batch* b, int16_t* uv, uint64_t vtx_capacity, vgb_draw_cmd* cmds, uint32_t cmd_capacity,
                                     uint32_t* num_cmds, void** uv_device);

/* Instance replay -- the device-side analogue of vg.cpp's shape cache
 * (addCachedCommand vg.cpp:5802-5841, submitCachedMesh vg.cpp:6137-6160): the meshes of the last run
 * stay resident and are written out n more times, copy k with
 *     pos = T_k( Tinv( pos_cached ) )      both by vgutil::batchTransformPositions (scalar build,
 *                                          vg_util.h:24-28), the intermediate rounded to float as the
 *                                          cache stores it; inv_xform = NULL skips Tinv
 *     colour copied, uint16 indices copied (per-mesh local) / uint32 merged indices + k * V,
 * V = vertices of the cached run.  The index type is the last run's; not available after a
 * final-layout run.  Mesh (k, m) starts at k*V + mesh_vtx_off[m] / k*I + mesh_idx_off[m].
 * Scope: ONE matrix per instance for all meshes.  That is bit-for-bit the reference's cache replay for lists that run
 * under their entry state throughout (checked against vg.cpp's real shape cache, tests/test_vgctx.py).  The reference
 * keeps an inverse matrix per cached COMMAND (vg.cpp:5788) and re-executes the list's state commands on replay
 * (clCacheRender): a list with relative transforms inside (translate / scale / rotate / mult) then agrees with the
 * one-matrix form to rounding only (<= 2.5e-4 at coordinates ~1000, measured), one that resets the transform
 * (TransformIdentity, SetViewBox) or changes global alpha inside does not -- replay such lists from one batch per group
 * of commands that share a state.  While RECORDING a cache vg.cpp also tessellates draws whose alpha is zero; a plain run
 * (this library's decoder included) skips them. */
VGB_API int vgb_batch_replay(vgb_batch* b, const float* xforms /* n x 6 */, uint32_t n, const float* inv_xform /* 6 or NULL */);
/* sizes of the replay (num_vertices = n*V, num_indices = n*I, num_meshes = n*M), device pointers, host copy */
VGB_API int vgb_batch_replay_sizes(vgb_batch* b, vgb_sizes* out);
VGB_API int vgb_batch_replay_device_ptrs(vgb_batch* b, void** pos, void** color, void** index);
VGB_API int vgb_batch_download_replay(vgb_batch* b, float* pos, uint32_t* color, void* index, uint64_t vtx_capacity, uint64_t idx_capacity);

/* device-resident results (valid until the next run/reset/destroy) */
VGB_API int vgb_batch_device_ptrs(vgb_batch* b, void** pos, void** color, void** index,
                                  void** mesh_vtx_off_u64, void** mesh_idx_off_u64,
                                  void** points, void** subpaths);

/* copy results to host buffers (any may be NULL to skip).  Capacities in elements.
 * mesh_vtx_off / mesh_idx_off get num_meshes+1 exclusive offsets. */
VGB_API int vgb_batch_download(vgb_batch* b, float* pos, uint32_t* color, void* index,
                               uint64_t vtx_capacity, uint64_t idx_capacity,
                               uint64_t* mesh_vtx_off, uint64_t* mesh_idx_off);
/* flattened polylines: points (x,y), sub-path table (vg::SubPath layout) */
VGB_API int vgb_batch_download_paths(vgb_batch* b, float* pts, uint64_t pts_capacity,
                                     vgb_subpath* subpaths, uint32_t subpath_capacity);

/* one call: upload + run + download into caller (ideally pinned) host buffers, D2H overlapped
 * with nothing else; the end-to-end path bench.py times.  sizes_out may be NULL. */
VGB_API int vgb_batch_execute(vgb_batch* b, uint32_t flags, float* pos, uint32_t* color, void* index,
                              uint64_t vtx_capacity, uint64_t idx_capacity,
                              uint64_t* mesh_vtx_off, uint64_t* mesh_idx_off, vgb_sizes* sizes_out);

/* elapsed device time (ms, CUDA events on the batch stream) of the kernels of the last run */
VGB_API int vgb_batch_last_run_ms(vgb_batch* b, float* total_ms, float* flatten_ms, float* stroke_count_ms, float* emit_ms);

/* user timing marks: cudaEventRecord on the batch stream into slot 0..7, and the elapsed device
 * time between two recorded slots (synchronises on the later one) */
VGB_API int vgb_batch_mark(vgb_batch* b, uint32_t slot);
VGB_API int vgb_batch_elapsed_ms(vgb_batch* b, uint32_t slot_a, uint32_t slot_b, float* ms);

/* pinned host memory helpers (cudaHostAlloc / cudaFreeHost) for the host-buffer path */
VGB_API void* vgb_host_alloc(size_t bytes);
VGB_API void vgb_host_free(void* p);
/* plain device -> host copy of a pointer handed out by vgb_batch_device_ptrs / vgb_group_gather (blocking) */
VGB_API int vgb_copy_to_host(void* dst, const void* src_device, size_t bytes);
/* raw stream handle (cudaStream_t) of the batch, for callers that time with their own events */
VGB_API void* vgb_batch_stream(vgb_batch* b);

/* ---- per-object API, 1:1 with include/vg/path.h and include/vg/stroker.h ------------------
 * Same call pattern, ownership and (absence of) error reporting as the reference: getters and
 * vgb_mesh pointers are borrowed views valid until the next mutating call on the object. */
typedef struct vgb_path vgb_path;
typedef struct vgb_stroker vgb_stroker;
/* allocator callback = the single virtual of bx::AllocatorI; NULL -> malloc/free */
typedef void* (*vgb_realloc_fn)(void* user, void* ptr, size_t size, size_t align);

VGB_API vgb_path* vgb_path_create(vgb_realloc_fn alloc, void* user);
VGB_API void vgb_path_destroy(vgb_path* p);
VGB_API void vgb_path_reset(vgb_path* p, float scale, float tess_tol);
VGB_API void vgb_path_move_to(vgb_path* p, float x, float y);
VGB_API void vgb_path_line_to(vgb_path* p, float x, float y);
VGB_API void vgb_path_cubic_to(vgb_path* p, float c1x, float c1y, float c2x, float c2y, float x, float y);
VGB_API void vgb_path_quadratic_to(vgb_path* p, float cx, float cy, float x, float y);
VGB_API void vgb_path_arc_to(vgb_path* p, float x1, float y1, float x2, float y2, float r);
VGB_API void vgb_path_rect(vgb_path* p, float x, float y, float w, float h);
VGB_API void vgb_path_rounded_rect(vgb_path* p, float x, float y, float w, float h, float r);
VGB_API void vgb_path_rounded_rect_varying(vgb_path* p, float x, float y, float w, float h, float rtl, float rtr, float rbr, float rbl);
VGB_API void vgb_path_circle(vgb_path* p, float x, float y, float r);
VGB_API void vgb_path_ellipse(vgb_path* p, float x, float y, float rx, float ry);
VGB_API void vgb_path_arc(vgb_path* p, float x, float y, float r, float a0, float a1, uint32_t dir);
VGB_API void vgb_path_polyline(vgb_path* p, const float* coords, uint32_t num_points);
VGB_API void vgb_path_close(vgb_path* p);
VGB_API const float* vgb_path_get_vertices(vgb_path* p);
VGB_API uint32_t vgb_path_get_num_vertices(vgb_path* p);
VGB_API const vgb_subpath* vgb_path_get_subpaths(vgb_path* p);
VGB_API uint32_t vgb_path_get_num_subpaths(vgb_path* p);

VGB_API vgb_stroker* vgb_stroker_create(vgb_realloc_fn alloc, void* user);
VGB_API void vgb_stroker_destroy(vgb_stroker* s);
VGB_API void vgb_stroker_reset(vgb_stroker* s, float scale, float tess_tol, float fringe_width);
VGB_API void vgb_stroker_polyline_stroke(vgb_stroker* s, vgb_mesh* mesh, const float* vtx, uint32_t n, int closed, float width, uint32_t cap, uint32_t join);
VGB_API void vgb_stroker_polyline_stroke_aa(vgb_stroker* s, vgb_mesh* mesh, const float* vtx, uint32_t n, int closed, uint32_t color, float width, uint32_t cap, uint32_t join);
VGB_API void vgb_stroker_polyline_stroke_aa_thin(vgb_stroker* s, vgb_mesh* mesh, const float* vtx, uint32_t n, int closed, uint32_t color, uint32_t cap, uint32_t join);
VGB_API void vgb_stroker_convex_fill(vgb_stroker* s, vgb_mesh* mesh, const float* vtx, uint32_t n);
VGB_API void vgb_stroker_convex_fill_aa(vgb_stroker* s, vgb_mesh* mesh, const float* vtx, uint32_t n, uint32_t color);

#ifdef __cplusplus
}
#endif

#endif /* VGB_H */
