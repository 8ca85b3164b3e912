/*
 * vgb_device.cu -- the sm_100a device pipeline of libvgb200.
 *
 * Kernels (all HBM-streaming fp32/integer work: no tensor cores -- there is no dense
 * contraction anywhere on this path):
 *
 *   k_flatten_slots    thread per sub-path, ONE pass: walk commands, store (transformed) points into
 *                      the sub-path's recorder-sized slot (overflow arena + two-pass fallback
 *                      k_flatten<false/true> + scan)
 *   k_cmd_index, k_cmd_flatten<count/emit>, k_cmd_sub_sizes   command-parallel flatten for scenes of long
 *                      sub-paths (thread per command from the previous command's nominal end point; sub-paths
 *                      whose end points survive the dedupe are exactly the sequential walk, the rest fall back)
 *   k_bulk_cubic_tables  bulk mode: all tables generated on the device from the caller's coordinates
 *   k_draw_prep        per-draw constants (hsw, da, nHC, nfix ...) with the shared fp32 math
 *   k_job_pieces       pieces per job (mesh) -> scan -> pieceStart;  k_job_prep: 48-byte job records
 *                      + per-piece job ids
 *   k_group_count      lane per piece, 2 groups per warp: sizes -> per-group (V,I) totals, hand-over words,
 *                      within-group mesh starts (geometry-free shortcut when sizes are provable)
 *   scan (u32x2->u64x2), k_mesh_off   group bases, mesh table
 *   k_group_emit       the hot kernel: a warp owns 32 pieces = one contiguous output range; lanes emit
 *                      vertices / colours / indices into a warp-private shared-memory image that
 *                      leaves through the TMA engine (cp.async.bulk), no block barrier anywhere
 *   k_fill_job_sizes, k_fill_emit   pure convex-AA-fill batches: closed-form sizes and addresses
 *   k_seg_next/_double/_cmds/_mesh_base, k_fill_u32   final layout: greedy vertex-buffer
 *                      segmentation by pointer doubling, draw-command table, per-mesh index bases, uv stream
 *   k_split_marks/_pos/_flags/_cmds/_fin   the same table when scissor commands force draw commands inside a
 *                      vertex buffer (m_ForceNewDrawCommand): two scans over per-mesh flags
 *   k_replay_vertices, k_replay_indices   instance replay: cached meshes x n transforms
 *
 * Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (see Makefile).  -fmad=false
 * plus vgb_math.h's __f*_rn intrinsics make every float result identical to the non-FMA CPU
 * oracle, which is what keeps vertex COUNTS (hence index buffers) bit-exact.
 */
#include "vgb_flatten.cuh"
#include "vgb_stroke.cuh"
#include "vgb_internal.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace vgb
{
#define VGB_CUDA(call)                                                                        \
	do {                                                                                      \
		cudaError_t e_ = (call);                                                              \
		if (e_ != cudaSuccess) {                                                              \
			if (err) *err = std::string(#call) + ": " + cudaGetErrorString(e_);               \
			return VGB_ERR_CUDA;                                                              \
		}                                                                                     \
	} while (0)

#ifndef VGB_CMD_FLATTEN_DEFAULT
#define VGB_CMD_FLATTEN_DEFAULT 1
#endif
#ifndef VGB_EMIT_PREFETCH
#define VGB_EMIT_PREFETCH 1
#endif
#ifndef VGB_SINGLE_PASS_DEFAULT
#define VGB_SINGLE_PASS_DEFAULT 0 /* k_group_emit_sp instead of count + scans + k_group_emit when the output buffers exist */
#endif
#ifndef VGB_EMIT_MIN_BLOCKS_SPEC
#define VGB_EMIT_MIN_BLOCKS_SPEC 8
#endif
#ifndef VGB_EMIT_MIN_BLOCKS
#define VGB_EMIT_MIN_BLOCKS 6
#endif
static constexpr uint32_t PPB = 128; /* threads per block of the piece kernels (4 warps = 4 groups) */

/* =========================================================================================
 * scans
 * ======================================================================================= */
template<class T> struct ScanOps;
template<> struct ScanOps<uint32_t>
{
	typedef uint32_t In;
	typedef uint32_t Acc;
	__device__ static Acc zero() { return 0u; }
	__device__ static Acc cvt(In a) { return a; }
	__device__ static Acc add(Acc a, Acc b) { return a + b; }
	__device__ static Acc shflUp(Acc a, int d) { return __shfl_up_sync(0xffffffffu, a, d); }
};
template<> struct ScanOps<uint2>
{
	typedef uint2 In;
	typedef ulonglong2 Acc;
	__device__ static Acc zero() { return make_ulonglong2(0ull, 0ull); }
	__device__ static Acc cvt(In a) { return make_ulonglong2(a.x, a.y); }
	__device__ static Acc add(Acc a, Acc b) { return make_ulonglong2(a.x + b.x, a.y + b.y); }
	__device__ static Acc shflUp(Acc a, int d)
	{
		return make_ulonglong2(__shfl_up_sync(0xffffffffu, a.x, d), __shfl_up_sync(0xffffffffu, a.y, d));
	}
};

static constexpr uint32_t SCAN_THREADS = 256;
static constexpr uint32_t SCAN_ITEMS = 4;
static constexpr uint32_t SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

/* inclusive block scan of one value per thread; returns inclusive value, *total = block sum */
template<class O, bool TRAILING_SYNC = true>
__device__ __forceinline__ typename O::Acc blockInclusive(typename O::Acc v, typename O::Acc* total)
{
	typedef typename O::Acc Acc;
	__shared__ Acc warpSum[32];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const Acc u = O::shflUp(v, d);
		if (lane >= (uint32_t)d) v = O::add(v, u);
	}
	if (lane == 31) warpSum[warp] = v;
	__syncthreads();
	if (warp == 0) {
		Acc w = lane < nw ? warpSum[lane] : O::zero();
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const Acc u = O::shflUp(w, d);
			if (lane >= (uint32_t)d) w = O::add(w, u);
		}
		warpSum[lane] = w;
	}
	__syncthreads();
	if (warp > 0) v = O::add(v, warpSum[warp - 1]);
	*total = warpSum[nw - 1];
	if (TRAILING_SYNC) {
		__syncthreads(); /* warpSum may be reused by a following call */
	}
	return v;
}

template<class O>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const typename O::In* __restrict__ in, uint32_t n, typename O::Acc* __restrict__ partial)
{
	typedef typename O::Acc Acc;
	const uint32_t base = blockIdx.x * SCAN_TILE;
	Acc s = O::zero();
#pragma unroll
	for (uint32_t k = 0; k < SCAN_ITEMS; ++k) {
		const uint32_t i = base + k * SCAN_THREADS + threadIdx.x;
		if (i < n) s = O::add(s, O::cvt(in[i]));
	}
	Acc total;
	blockInclusive<O>(s, &total);
	if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

__device__ __forceinline__ uint32_t accSub(uint32_t a, uint32_t b) { return a - b; }
__device__ __forceinline__ ulonglong2 accSub(ulonglong2 a, ulonglong2 b) { return make_ulonglong2(a.x - b.x, a.y - b.y); }

/* one block: exclusive scan of the partials in place */
template<class O>
__global__ void __launch_bounds__(1024) k_scan_partials(typename O::Acc* __restrict__ partial, uint32_t m)
{
	typedef typename O::Acc Acc;
	__shared__ Acc carry;
	if (threadIdx.x == 0) carry = O::zero();
	__syncthreads();
	for (uint32_t base = 0; base < m; base += blockDim.x) {
		const uint32_t i = base + threadIdx.x;
		const Acc v = i < m ? partial[i] : O::zero();
		Acc total;
		const Acc inc = blockInclusive<O>(v, &total);
		const Acc c = carry;
		if (i < m) partial[i] = accSub(O::add(c, inc), v);
		__syncthreads();
		if (threadIdx.x == 0) carry = O::add(c, total);
		__syncthreads();
	}
}

/* out[i] = exclusive prefix, out[n] = total */
template<class O>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_final(const typename O::In* __restrict__ in, uint32_t n,
                                                            const typename O::Acc* __restrict__ partial, typename O::Acc* __restrict__ out)
{
	typedef typename O::Acc Acc;
	const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
	Acc v[SCAN_ITEMS];
	Acc s = O::zero();
#pragma unroll
	for (uint32_t k = 0; k < SCAN_ITEMS; ++k) {
		v[k] = (base + k < n) ? O::cvt(in[base + k]) : O::zero();
		s = O::add(s, v[k]);
	}
	Acc total;
	const Acc inc = blockInclusive<O>(s, &total);
	Acc run = O::add(partial[blockIdx.x], accSub(inc, s));
#pragma unroll
	for (uint32_t k = 0; k < SCAN_ITEMS; ++k) {
		if (base + k < n) out[base + k] = run;
		run = O::add(run, v[k]);
		if (base + k + 1 == n) out[n] = run;
	}
	if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) out[0] = O::zero();
}

/* =========================================================================================
 * flatten
 * ======================================================================================= */
struct FlattenArgs
{
	const SubRec* subRec;
	uint32_t numSub;
	const uint8_t* cmd;
	const float* data;
	const float* pathScale;
	const float* pathTol;
	const float* pathXform;
	uint32_t* subNum;         /* count pass output                  */
	const uint32_t* subFirst; /* emit pass input (scanned)          */
	uint8_t* subClosed;       /* emit pass output                   */
	float2* pts;              /* emit pass output                   */
	const uint8_t* subMode;   /* nullable; 1 = this sub-path is flattened per COMMAND (k_cmd_flatten): skip it here */
};

template<bool EMIT>
__global__ void __launch_bounds__(128) k_flatten(FlattenArgs a)
{
	__shared__ float sStack[10 * 6 * 128];
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= a.numSub) return;
	if (a.subMode != nullptr && a.subMode[s] != 0) return;
	const SubRec r = a.subRec[s];
	SubPathWalker<EMIT> w;
	w.lx = w.ly = w.px = w.py = w.fx = w.fy = 0.0f;
	w.n = 0;
	w.closed = false;
	w.scale = a.pathScale[r.pathId];
	w.tol = a.pathTol[r.pathId];
	w.m = (EMIT && a.pathXform) ? a.pathXform + 6 * (size_t)r.pathId : nullptr;
	w.out = EMIT ? a.pts + a.subFirst[s] : nullptr;
	w.limit = EMIT ? a.subNum[s] : 0u;
	w.stk = sStack + threadIdx.x;
	w.stride = 128;
	w.run(a.cmd, r.cmdBegin, r.cmdEnd, a.data + r.dataBegin);
	if (EMIT) {
		a.subClosed[s] = w.closed ? 1 : 0;
	} else {
		a.subNum[s] = w.n;
	}
}

/* ---- command-parallel flatten (two passes, dense output) -----------------------------------------
 * A thread per COMMAND instead of per sub-path, for scenes of few long sub-paths (a thread per sub-path
 * then is one wave of long serial chains).  What command c needs from its predecessors is only the path's
 * last kept vertex E_c (P1 of a curve, the dedupe partner of its first vertex; path.cpp:93-97, 761-784).
 * For moveTo / lineTo / cubicTo / quadTo the previous command's nominal end point is in the command data,
 * and E_c equals it unless the epsilon-dedupe dropped that end point.  So every command is flattened on
 * its own from the nominal E_c and reports whether its own end point was kept; a sub-path all of whose
 * commands report so is exactly the sequential walk (tools/proto_command_parallel_flatten.py checks this
 * against the oracle), any other sub-path -- and any with a command whose end point is computed (arcs,
 * shapes, polylines) -- is left to k_flatten.  pathClose (path.cpp:707-726) is applied per sub-path from
 * the first and the last nominal point. */
struct CmdArgs
{
	FlattenArgs f;
	uint32_t numCmd;
	uint32_t* cmdSub;   /* per command: its sub-path                              */
	uint32_t* cmdData;  /* per command: offset of its floats in data[]            */
	uint32_t* cmdCount; /* per command: vertices it contributes                   */
	uint32_t* cmdOff;   /* per command: first vertex inside its sub-path          */
	uint8_t* cmdOk;     /* per command: its nominal end point was kept            */
	uint8_t* subModeW;  /* per sub-path: 1 = handled here, 0 = left to k_flatten   */
};

__device__ __forceinline__ bool cmdIsSimple(uint32_t t)
{
	return t == VGB_CMD_MOVE_TO || t == VGB_CMD_LINE_TO || t == VGB_CMD_CUBIC_TO || t == VGB_CMD_QUAD_TO;
}

/* nominal end point of a simple command: the last two floats of its data */
__device__ __forceinline__ float2 cmdEnd(uint32_t t, const float* __restrict__ d)
{
	const uint32_t n = cmdArity(t);
	return make_float2(d[n - 2], d[n - 1]);
}

/* per sub-path: command -> (sub-path, data offset) and the static part of the eligibility */
__global__ void k_cmd_index(CmdArgs a)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= a.f.numSub) return;
	const SubRec r = a.f.subRec[s];
	uint32_t off = r.dataBegin;
	bool ok = r.cmdEnd > r.cmdBegin && a.f.cmd[r.cmdBegin] == VGB_CMD_MOVE_TO;
	for (uint32_t c = r.cmdBegin; c < r.cmdEnd; ++c) {
		const uint32_t t = a.f.cmd[c];
		a.cmdSub[c] = s;
		a.cmdData[c] = off;
		if (c > r.cmdBegin) {
			const bool lastClose = t == VGB_CMD_CLOSE && c + 1 == r.cmdEnd;
			ok = ok && (lastClose || (cmdIsSimple(t) && t != VGB_CMD_MOVE_TO));
		}
		if (t == VGB_CMD_POLYLINE) {
			off += 1u + 2u * __float_as_uint(a.f.data[off]);
		} else {
			off += cmdArity(t);
		}
	}
	a.subModeW[s] = ok ? 1 : 0;
}

template<bool EMIT>
__global__ void __launch_bounds__(128) k_cmd_flatten(CmdArgs a)
{
	__shared__ float sStack[10 * 6 * 128];
	const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= a.numCmd) return;
	const uint32_t s = a.cmdSub[c];
	if (a.subModeW[s] == 0) return;
	const SubRec r = a.f.subRec[s];
	const uint32_t t = a.f.cmd[c];
	if (t == VGB_CMD_CLOSE) {
		if (!EMIT) {
			a.cmdCount[c] = 0;
			a.cmdOk[c] = 1;
		}
		return;
	}
	const float* d = a.f.data + a.cmdData[c];
	const bool isMove = c == r.cmdBegin;
	SubPathWalker<EMIT> w;
	w.px = w.py = w.fx = w.fy = 0.0f;
	w.lx = w.ly = 0.0f;
	w.n = 0;
	if (!isMove) {
		const float2 e = cmdEnd(a.f.cmd[c - 1], a.f.data + a.cmdData[c - 1]);
		w.lx = e.x;
		w.ly = e.y;
		w.n = 1; /* the sub-path is not empty: the dedupe against (lx, ly) applies */
	}
	w.closed = false;
	w.scale = a.f.pathScale[r.pathId];
	w.tol = a.f.pathTol[r.pathId];
	w.m = (EMIT && a.f.pathXform) ? a.f.pathXform + 6 * (size_t)r.pathId : nullptr;
	uint32_t off = 0;
	if (EMIT) {
		off = a.cmdOff[c];
		w.out = a.f.pts + a.f.subFirst[s] + off - (isMove ? 0u : 1u);
		w.limit = (a.f.subNum[s] - off) + (isMove ? 0u : 1u); /* a vertex dropped by pathClose is never stored */
	} else {
		w.out = nullptr;
		w.limit = 0;
	}
	w.stk = sStack + threadIdx.x;
	w.stride = 128;
	w.run(a.f.cmd, c, c + 1, d);
	if (!EMIT) {
		const float2 e = cmdEnd(t, d);
		a.cmdCount[c] = w.n - (isMove ? 0u : 1u);
		a.cmdOk[c] = (w.lx == e.x && w.ly == e.y) ? 1 : 0;
	}
}

/* per sub-path: all end points kept? -> vertex offsets of its commands, its size, pathClose */
__global__ void k_cmd_sub_sizes(CmdArgs a)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= a.f.numSub) return;
	if (a.subModeW[s] == 0) return;
	const SubRec r = a.f.subRec[s];
	uint32_t total = 0;
	bool ok = true;
	for (uint32_t c = r.cmdBegin; c < r.cmdEnd; ++c) {
		ok = ok && a.cmdOk[c] != 0;
		a.cmdOff[c] = total;
		total += a.cmdCount[c];
	}
	if (!ok) {
		a.subModeW[s] = 0; /* an end point was deduped away: the sequential walk decides */
		return;
	}
	bool closed = false;
	const uint32_t lastCmd = r.cmdEnd - 1;
	if (a.f.cmd[lastCmd] == VGB_CMD_CLOSE && total > 2) { /* pathClose, path.cpp:707-726 */
		closed = true;
		const float2 first = cmdEnd(VGB_CMD_MOVE_TO, a.f.data + r.dataBegin);
		const float2 last = cmdEnd(a.f.cmd[lastCmd - 1], a.f.data + a.cmdData[lastCmd - 1]);
		const float dx = last.x - first.x;
		const float dy = last.y - first.y;
		const float distSqr = dx * dx + dy * dy;
		if (distSqr < VGB_EPSILON) {
			--total;
		}
	}
	a.f.subNum[s] = total;
	a.f.subClosed[s] = closed ? 1 : 0;
}

/* Single-pass flatten: every sub-path owns a slot of the point buffer sized by the recorder, so
 * one walk both counts and stores -- no count pass, no scan, no size read-back.  A sub-path that
 * outgrows its slot (the walk keeps counting past the slot's end) claims space in an overflow
 * arena with one atomicAdd and walks again; the point buffer is therefore not dense, which is
 * fine because everything downstream addresses points through (subFirst, subNum). */
struct SlotArgs
{
	FlattenArgs f;
	const uint32_t* slotFirst; /* numSub + 1 */
	uint32_t* subFirstOut;
	uint32_t arenaBase, arenaCap;
	uint32_t* counters; /* [0] arena cursor, [1] overflow-error flag, [2] total points */
};

__global__ void __launch_bounds__(128) k_flatten_slots(SlotArgs a)
{
	__shared__ float sStack[10 * 6 * 128];
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t n = 0;
	if (s < a.f.numSub) {
		const SubRec r = a.f.subRec[s];
		uint32_t first = a.slotFirst[s];
		uint32_t cap = a.slotFirst[s + 1] - first;
		bool closed = false;
		for (int pass = 0; pass < 2; ++pass) {
			SubPathWalker<true> w;
			w.lx = w.ly = w.px = w.py = w.fx = w.fy = 0.0f;
			w.n = 0;
			w.closed = false;
			w.scale = a.f.pathScale[r.pathId];
			w.tol = a.f.pathTol[r.pathId];
			w.m = a.f.pathXform ? a.f.pathXform + 6 * (size_t)r.pathId : nullptr;
			w.out = a.f.pts + first;
			w.limit = cap;
			w.stk = sStack + threadIdx.x;
			w.stride = 128;
			w.run(a.f.cmd, r.cmdBegin, r.cmdEnd, a.f.data + r.dataBegin);
			n = w.n;
			closed = w.closed;
			if (n <= cap) {
				break;
			}
			/* slot too small: claim n points of the arena and walk once more */
			const uint32_t base = atomicAdd(&a.counters[0], n);
			if (pass == 1 || base + n > a.arenaCap) {
				atomicOr(&a.counters[1], 1u);
				n = 0;
				break;
			}
			first = a.arenaBase + base;
			cap = n;
		}
		a.subFirstOut[s] = first;
		a.f.subNum[s] = n;
		a.f.subClosed[s] = closed ? 1 : 0;
	}
	const uint32_t tot = __reduce_add_sync(0xffffffffu, n);
	if ((threadIdx.x & 31) == 0 && tot) {
		atomicAdd(&a.counters[2], tot);
	}
}

/* bulk mode (vgb_batch_set_cubic_strokes): the coordinates arrive straight from the caller's
 * buffer; every table the pipeline needs for "path i = moveTo + cubicTo, one stroke" is generated
 * here instead of being built and shipped by the host.  A sub-path with a non-finite coordinate
 * gets an empty command range (the reference would never return from pathCubicTo) and raises the
 * error flag. */
static constexpr uint32_t BULK_SLOT = 52; /* point slots per bulk cubic (moveTo 1 + cubic 48 + spare) */

__global__ void k_bulk_cubic_tables(const float* __restrict__ p8, uint32_t n, float scale, float tol, uint8_t* __restrict__ cmd,
                                    SubRec* __restrict__ subRec, float* __restrict__ pathScale, float* __restrict__ pathTol,
                                    uint32_t* __restrict__ jobSub, uint32_t* __restrict__ jobDraw, uint32_t* __restrict__ slotFirst,
                                    uint32_t* __restrict__ errFlag)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > n) return;
	slotFirst[i] = i * BULK_SLOT;
	if (i == n) return;
	const float4 a = reinterpret_cast<const float4*>(p8)[2 * i], b = reinterpret_cast<const float4*>(p8)[2 * i + 1];
	const uint32_t bits = (__float_as_uint(a.x) & 0x7F800000u) == 0x7F800000u || (__float_as_uint(a.y) & 0x7F800000u) == 0x7F800000u ||
	                      (__float_as_uint(a.z) & 0x7F800000u) == 0x7F800000u || (__float_as_uint(a.w) & 0x7F800000u) == 0x7F800000u ||
	                      (__float_as_uint(b.x) & 0x7F800000u) == 0x7F800000u || (__float_as_uint(b.y) & 0x7F800000u) == 0x7F800000u ||
	                      (__float_as_uint(b.z) & 0x7F800000u) == 0x7F800000u || (__float_as_uint(b.w) & 0x7F800000u) == 0x7F800000u;
	if (bits) {
		atomicOr(errFlag, 1u);
	}
	reinterpret_cast<uchar2*>(cmd)[i] = make_uchar2(VGB_CMD_MOVE_TO, VGB_CMD_CUBIC_TO);
	SubRec r;
	r.cmdBegin = 2 * i;
	r.cmdEnd = bits ? 2 * i : 2 * i + 2;
	r.dataBegin = 8 * i;
	r.pathId = i;
	subRec[i] = r;
	pathScale[i] = scale;
	pathTol[i] = tol;
	jobSub[i] = i;
	jobDraw[i] = 0;
}

/* merged run: every rank's (V, I) -> this rank's exclusive base and the group totals.  One warp. */
__global__ void k_rank_base(const uint64_t* __restrict__ rankTotals, uint32_t rank, uint32_t nranks, uint64_t* __restrict__ out)
{
	uint64_t bv = 0, bi = 0, tv = 0, ti = 0;
	for (uint32_t r = threadIdx.x; r < nranks; r += 32) {
		const uint64_t v = rankTotals[2 * r], i = rankTotals[2 * r + 1];
		tv += v;
		ti += i;
		if (r < rank) {
			bv += v;
			bi += i;
		}
	}
	for (int d = 16; d > 0; d >>= 1) {
		bv += __shfl_xor_sync(0xffffffffu, bv, d);
		bi += __shfl_xor_sync(0xffffffffu, bi, d);
		tv += __shfl_xor_sync(0xffffffffu, tv, d);
		ti += __shfl_xor_sync(0xffffffffu, ti, d);
	}
	if (threadIdx.x == 0) {
		out[0] = bv;
		out[1] = bi;
		out[2] = tv;
		out[3] = ti;
	}
}

/* mesh table as two plain u64 arrays (what vgb_batch_download hands out) */
__global__ void k_split_mesh_off(const ulonglong2* __restrict__ meshOff, uint32_t n, uint64_t* __restrict__ v, uint64_t* __restrict__ i)
{
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	const ulonglong2 o = meshOff[j];
	v[j] = o.x;
	i[j] = o.y;
}

/* dense copy of a slotted point buffer (only for reading paths back: download_paths) */
__global__ void k_gather_points(const float2* __restrict__ pts, const uint32_t* __restrict__ subFirst, const uint32_t* __restrict__ subNum,
                                const uint32_t* __restrict__ denseFirst, uint32_t numSub, float2* __restrict__ out)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= numSub) return;
	const uint32_t f = subFirst[s], n = subNum[s], d = denseFirst[s];
	for (uint32_t i = 0; i < n; ++i) {
		out[d + i] = pts[f + i];
	}
}

/* =========================================================================================
 * draws / jobs / pieces
 * ======================================================================================= */
__global__ void k_draw_prep(const DrawRec* __restrict__ draws, uint32_t n, const float* __restrict__ pathScale,
                            const float* __restrict__ pathTol, float fringe, DrawDerived* __restrict__ out)
{
	const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
	if (d >= n) return;
	const DrawRec r = draws[d];
	const float scale = pathScale[r.pathId];
	const float tol = pathTol[r.pathId];
	const uint32_t op = r.ocj & 0xFF;
	DrawDerived o;
	o.fringe = fringe;
	o.color = r.color;
	o.c0 = r.color & 0x00FFFFFFu; /* colorSetAlpha(color, 0), vg.inl:92-95 */
	o.ocj = r.ocj;
	o.hsw = o.hsw_aa = o.da = 0.0f;
	o.nHC = 2;
	if (op == VGB_OP_STROKE_AA) {
		/* stroker.cpp:1396-1399 */
		o.hsw = (r.width - fringe) * 0.5f;
		o.hsw_aa = o.hsw + fringe;
		o.da = vgb_acos((scale * o.hsw) / ((scale * o.hsw) + tol)) * 2.0f;
		o.nHC = vgb_umax(2u, vgb_f2u(vgb_ceil(VGB_PI / o.da)));
		if (3.0f * o.da > VGB_PI + 1e-3f) {
			o.ocj |= 1u << 24; /* nfix: see quickCountAA */
		}
	} else if (op == VGB_OP_STROKE) {
		/* stroker.cpp:1012-1014; hsw_aa carries the side-test scale (v * hsw, 1096-1099) */
		o.hsw = r.width * 0.5f;
		o.hsw_aa = o.hsw;
		o.da = vgb_acos((scale * o.hsw) / ((scale * o.hsw) + tol)) * 2.0f;
		o.nHC = vgb_umax(2u, vgb_f2u(vgb_ceil(VGB_PI / o.da)));
	} else if (op == VGB_OP_STROKE_AA_THIN) {
		/* stroker.cpp:1999 (half width = fringe) and the dispatch of 311-332: round cap -> square,
		 * round join -> bevel */
		o.hsw_aa = fringe;
		uint32_t cap = (r.ocj >> 8) & 0xFF, join = (r.ocj >> 16) & 0xFF;
		if (cap == VGB_CAP_ROUND) cap = VGB_CAP_SQUARE;
		if (join == VGB_JOIN_ROUND) join = VGB_JOIN_BEVEL;
		o.ocj = op | (cap << 8) | (join << 16);
	}
	out[d] = o;
}

__global__ void k_job_pieces(const uint32_t* __restrict__ jobSub, const uint32_t* __restrict__ jobDraw, uint32_t numJobs,
                             const uint32_t* __restrict__ subNum, const DrawRec* __restrict__ draws, uint32_t* __restrict__ jobPieces, uint32_t sseFill)
{
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= numJobs) return;
	jobPieces[j] = pieceCount(draws[jobDraw[j]].ocj & 0xFF, subNum[jobSub[j]], sseFill != 0);
}

/* per job: the 48-byte record the piece kernels read, and for each of its pieces the record
 * {job id, index of the piece's own polyline vertex in pts[]} -- with the point index in hand a piece
 * kernel starts the job-record load and the point load together (one dependent round trip less than
 * piece -> job record -> points).  One warp handles 32 jobs; for every job of the warp all 32 lanes
 * write its piece range, so the stores are coalesced whatever the job lengths are. */
__global__ void __launch_bounds__(256) k_job_prep(const uint32_t* __restrict__ jobSub, const uint32_t* __restrict__ jobDraw, uint32_t numJobs,
                                                  const uint32_t* __restrict__ subFirst, const uint32_t* __restrict__ subNum,
                                                  const uint8_t* __restrict__ subClosed, const DrawDerived* __restrict__ dd,
                                                  const uint32_t* __restrict__ pieceStart, JobRec* __restrict__ jobRec,
                                                  uint2* __restrict__ pieceRec, uint32_t sseFill)
{
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t lane = threadIdx.x & 31;
	uint32_t ps = 0, pe = 0, first = 0, fan = 0;
	if (j < numJobs) {
		const uint32_t sp = jobSub[j];
		const DrawDerived d = dd[jobDraw[j]];
		JobRec r;
		r.first = subFirst[sp];
		r.N = subNum[sp];
		const uint32_t op = d.ocj & 0xFF, cap = (d.ocj >> 8) & 0xFF, join = (d.ocj >> 16) & 0xFF;
		r.flags = (uint32_t)(subClosed[sp] & 1) | (((d.ocj >> 24) & 1u) << 1) | ((sseFill & 1u) << 2) | (op << 8) | (cap << 16) | (join << 24);
		ps = pieceStart[j];
		pe = pieceStart[j + 1];
		first = r.first;
		/* fills: N-2 fan-triangle pieces (no vertex of their own) precede the N vertex pieces */
		fan = ((op == VGB_OP_CONVEX_FILL_AA && !sseFill) || op == VGB_OP_CONVEX_FILL) && r.N >= 2 ? r.N - 2 : 0u;
		r.pieceStart = ps;
		r.hsw = d.hsw;
		r.hsw_aa = d.hsw_aa;
		r.da = d.da;
		r.fringe = d.fringe;
		r.nHC = d.nHC;
		r.color = d.color;
		r.c0 = d.c0;
		r.pad = 0;
		jobRec[j] = r;
	}
	const uint32_t jw = j - lane; /* first job of this warp */
	for (uint32_t l = 0; l < 32; ++l) {
		const uint32_t s0 = __shfl_sync(0xffffffffu, ps, l);
		const uint32_t e0 = __shfl_sync(0xffffffffu, pe, l);
		if (e0 > s0) {
			const uint32_t f0 = __shfl_sync(0xffffffffu, first, l);
			const uint32_t n0 = __shfl_sync(0xffffffffu, fan, l);
			for (uint32_t k = s0 + lane; k < e0; k += 32) {
				const uint32_t q = k - s0;
				pieceRec[k] = make_uint2(jw + l, f0 + (q >= n0 ? q - n0 : 0u));
			}
		}
	}
}

struct PieceArgs
{
	uint32_t numPieces, numJobs;
	const uint2* pieceRec;    /* per piece: {job, index of its polyline vertex} */
	const JobRec* jobRec;     /* per job   */
	const float2* pts;
	/* count outputs */
	uint2* blockTotals;
	uint2* meshTotals;    /* zeroed before the count kernel */
	uint32_t* blockTail;  /* packTail() of each block's last piece (the next block's hand-over) */
	/* emit inputs */
	const ulonglong2* blockBase; /* numBlocks + 1 */
	const ulonglong2* meshOff;   /* numJobs + 1   */
	uint64_t vertexBase;
	const uint64_t* vertexBasePtr; /* merged run: this rank's base, computed on the device (nullptr otherwise) */
	const uint32_t* meshBase; /* final layout: per-mesh index base (nullptr otherwise) */
	uint32_t* ticket;         /* emit: dynamic group tickets (zeroed by k_mesh_off) */
	uint32_t chunkShift;      /* emit: a ticket is worth 1 << chunkShift consecutive groups */
	/* single-pass emit (k_group_emit_sp): group descriptors of the chained scan, output capacity, mesh table */
	uint4* spA;                /* per group: totals word  (see k_group_emit_sp) */
	uint4* spP;                /* per group: prefix word */
	uint32_t* spFlags;         /* bit 0: output capacity exceeded, bit 1: the look-back gave up (never expected) */
	uint64_t capV, capI;
	const uint32_t* pieceStart; /* numJobs + 1 */
	ulonglong2* meshOffW;       /* numJobs + 1, written by the groups that hold a job's first piece */
	/* emit outputs */
	float2* pos;
	uint32_t* col;
	void* idx;
};

__device__ __forceinline__ JobCtx loadJob(const PieceArgs& a, uint32_t j, uint32_t* pieceStart)
{
	const uint4* r = reinterpret_cast<const uint4*>(a.jobRec + j);
	const uint4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
	JobCtx J;
	J.first = r0.x;
	J.N = r0.y;
	J.closed = r0.z & 1u;
	J.nfix = (r0.z >> 1) & 1u;
	J.sse = (r0.z >> 2) & 1u;
	J.op = (r0.z >> 8) & 0xFF;
	J.cap = (r0.z >> 16) & 0xFF;
	J.join = r0.z >> 24;
	*pieceStart = r0.w;
	J.hsw = __uint_as_float(r1.x);
	J.hsw_aa = __uint_as_float(r1.y);
	J.da = __uint_as_float(r1.z);
	J.fringe = __uint_as_float(r1.w);
	J.nHC = r2.x;
	J.color = r2.y;
	J.c0 = r2.z;
	return J;
}

/* ---- shared -> global copy-out through the TMA engine (1-D bulk copy, no tensor map) ----------
 * The block's output image sits in shared memory laid out so that element e is at
 * sbuf[e - alignedBase]: 16-byte chunks of the image line up with 16-byte chunks of HBM.  The
 * aligned body goes out with ONE cp.async.bulk per stream issued by one thread (the copy runs
 * asynchronously in the TMA unit: no LSU instructions, no registers); only the ragged head/tail
 * (fewer than 16 bytes each) is stored by threads. */
__device__ __forceinline__ void fenceProxyAsync()
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulkStore(void* gdst, const void* ssrc, uint32_t bytes)
{
	const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(ssrc);
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkCommit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulkWaitRead() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

/* One WARP streams its three arrays out of its staging image.  The image holds element e of a stream
 * at sbuf[lo + e], lo = (global element index of e = 0) mod K, K = elements per 16 bytes, so 16-byte
 * chunks of the image line up with 16-byte chunks of HBM.  The aligned body [b0, b1) of each stream
 * leaves with ONE cp.async.bulk; the ragged ends (fewer than K elements each) are stored by lanes.
 * The three plans are evaluated side by side, one stream per group of lanes (32-bit arithmetic on
 * image coordinates; the only 64-bit quantity is the stream's 16-byte aligned global address):
 *   lanes 0-1   positions  (K = 2: one head, one tail element at most)      issuer: lane 0
 *   lanes 2-7   colours    (K = 4: up to three each)                        issuer: lane 2
 *   lanes 8-23  indices    (K = 8 or 4: up to seven each)                   issuer: lane 8 */
template<class IDX>
__device__ __forceinline__ void copyOutWarp3(float2* __restrict__ gPos, uint32_t* __restrict__ gCol, IDX* __restrict__ gIdx, uint64_t vs, uint64_t is,
                                             uint32_t totV, uint32_t totI, const float2* sPos, const uint32_t* sCol, const IDX* sIdx, uint32_t lane)
{
	constexpr uint32_t KI = 16 / sizeof(IDX);
	const uint32_t sid = lane < 2 ? 0u : (lane < 8 ? 1u : 2u);
	const uint32_t r = lane - (sid == 0 ? 0u : (sid == 1 ? 2u : 8u)); /* role inside the stream's lanes */
	const uint32_t K = sid == 0 ? 2u : (sid == 1 ? 4u : KI);
	const uint32_t half = sid == 0 ? 1u : (sid == 1 ? 3u : 8u);
	const uint32_t sh = sid == 0 ? 3u : (sid == 1 ? 2u : (sizeof(IDX) == 2 ? 1u : 2u)); /* log2(element bytes) */
	const uint64_t gs = sid == 2 ? is : vs;
	const uint32_t count = sid == 2 ? totI : totV;
	const uint32_t lo = (uint32_t)gs & (K - 1);
	const uint32_t end = lo + count;
	const uint32_t b0 = lo ? K : 0u;
	const uint32_t b1 = end & ~(K - 1);
	const bool body = b1 > b0;
	/* shared / global BYTE addresses of image coordinate 0 */
	const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sid == 0 ? (const void*)sPos : (sid == 1 ? (const void*)sCol : (const void*)sIdx));
	char* gbase = (sid == 0 ? (char*)gPos : (sid == 1 ? (char*)gCol : (char*)gIdx)) + ((gs - lo) << sh);
	if (r == 0 && lane < 9) {
		if (body) {
			asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gbase + (b0 << sh)), "r"(sbase + (b0 << sh)), "r"((b1 - b0) << sh)
			             : "memory");
		}
		bulkCommit();
	}
	/* ragged ends: head elements [lo, min(b0, end)), tail elements [max(b0, b1), end) */
	const bool isHead = r < half;
	const uint32_t e = isHead ? lo + r : (body ? b1 : b0) + (r - half);
	const bool ok = lane < 24 && (isHead ? e < (b0 < end ? b0 : end) : e < end);
	if (ok) {
		if (sid == 0) {
			reinterpret_cast<float2*>(gbase)[e] = sPos[e];
		} else if (sid == 1) {
			reinterpret_cast<uint32_t*>(gbase)[e] = sCol[e];
		} else {
			reinterpret_cast<IDX*>(gbase)[e] = sIdx[e];
		}
	}
}

template<uint32_t OPS, int JT, int CT, class SINK>
__device__ __forceinline__ void emitPiece(const float2* __restrict__ pts, const JobCtx& J, const Piece& P, uint32_t F, uint4 prev, const SINK& S)
{
	if ((OPS & OPS_OTHER) && (J.op == VGB_OP_STROKE || J.op == VGB_OP_STROKE_AA_THIN)) {
		if (P.kind == PK_JOIN) {
			const bool firstLeft = (P.flags & PF_CLOSING) ? firstJoinLeft(J, pts) : false;
			emitRibbonJoin(J, P, F, prev, firstLeft, S);
		} else {
			emitRibbonCap(J, P, F, prev, S);
		}
		return;
	}
	if ((OPS & OPS_OTHER) && P.kind == PK_FAN_TRI) {
		S.tri(0, 0, P.li + 1, P.li + 2);
		return;
	}
	if ((OPS & OPS_OTHER) && P.kind == PK_FAN_VTX) {
		S.V(0, P.p.x, P.p.y, J.color);
		return;
	}
	if (false) {
	} else if ((OPS & OPS_FILL_AA) && P.kind == PK_FILL_VTX) {
		emitFillVtx(J, P, pts, S);
	} else if ((OPS & OPS_FILL_AA) && P.kind == PK_FILL_FAN) {
		emitFillFan(P, S);
	}
}

template<uint32_t OPS, int JT, int CT, class SINK>
__device__ __forceinline__ void emitAny(const float2* __restrict__ pts, const JobCtx& J, const Piece& P, uint32_t F, uint4 prev, const SINK& S)
{
	if ((OPS & OPS_STROKE_AA) && (!(OPS & ~OPS_STROKE_AA) || J.op == VGB_OP_STROKE_AA)) {
		/* joins and round caps share one code path (see emitStrokeAA) */
		const bool firstLeft = (P.flags & PF_CLOSING) ? firstJoinLeft(J, pts) : false;
		emitStrokeAA<JT, CT>(J, P, F, prev, firstLeft, S);
	} else {
		emitPiece<OPS, JT, CT>(pts, J, P, F, prev, S);
	}
}

/* the same through the plain-pointer sink: it only runs for warps whose pieces do not fit the staging image.
 * Inlined on purpose: as a __noinline__ call its JobCtx / Piece arguments had to live in local memory, which put
 * STL/LDL on the COMMON path (the structs' addresses escape); inlined, the overflow code is a cold block behind a
 * warp-uniform branch and only costs instruction-cache footprint. */
template<uint32_t OPS, int JT, int CT, class SINK>
__device__ __forceinline__ void emitAnyCold(const float2* __restrict__ pts, const JobCtx& J, const Piece& P, uint32_t F, uint4 prev, const SINK& S)
{
	emitAny<OPS, JT, CT>(pts, J, P, F, prev, S);
}

/* =========================================================================================
 * count / emit: the unit that owns a contiguous output range is a WARP (a "group" of 32 pieces).
 * Offsets come from a shuffle scan, the hand-over from the previous lane (or the previous warp's
 * tail word), the staging image is private to the warp and leaves through the warp's own bulk
 * copies -- there is not a single block barrier, so a warp waiting on memory never holds its
 * three siblings.  (Round-1 history: a block-granular version with a block scan and caps detached
 * to warp 0 reached 1.65 ms on the headline workload; this one 1.16 ms -- profiles/README.md.)
 * ======================================================================================= */
#ifndef VGB_CVW
#define VGB_CVW 288
#endif
#ifndef VGB_CIW
#define VGB_CIW 1280
#endif
static constexpr uint32_t GRP = 32;                      /* pieces per offset group (one warp) */
static constexpr uint32_t CVW = VGB_CVW;                 /* staged vertices per warp */
static constexpr uint32_t CIW_COUNT = VGB_CIW;           /* staged indices per warp (uint16 or uint32) */

__device__ __forceinline__ uint2 warpInclusive(uint2 v, uint32_t lane)
{
	/* both components in one word when no carry can cross (32 x 1023 < 2^16, 32 x 2047 < 2^16): one shuffle
	 * and one add per step; fat pieces (wide strokes: long round caps / joins) take the two-component scan */
	if (__all_sync(0xffffffffu, v.x <= 1023u && v.y <= 2047u)) {
		uint32_t x = v.x | (v.y << 16);
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
			if (lane >= (uint32_t)d) {
				x += y;
			}
		}
		return make_uint2(x & 0xFFFFu, x >> 16);
	}
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t x = __shfl_up_sync(0xffffffffu, v.x, d);
		const uint32_t y = __shfl_up_sync(0xffffffffu, v.y, d);
		if (lane >= (uint32_t)d) {
			v.x += x;
			v.y += y;
		}
	}
	return v;
}

/* mesh table: (first vertex, first index) of job j = base of the group holding its first piece +
 * the within-group offset recorded by k_group_count; an empty job starts where the next piece starts */
__global__ void k_mesh_off(const uint32_t* __restrict__ pieceStart, const uint2* __restrict__ pieceRec, uint32_t numJobs, uint32_t numPieces,
                           uint32_t numGroups, const ulonglong2* __restrict__ groupBase, const uint2* __restrict__ meshLocal,
                           ulonglong2* __restrict__ meshOff, uint32_t* __restrict__ ticket)
{
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j == 0) *ticket = 0u; /* the emit kernel's group tickets */
	if (j > numJobs) return;
	const uint32_t g = j < numJobs ? pieceStart[j] : numPieces;
	ulonglong2 o;
	if (g >= numPieces) {
		o = groupBase[numGroups];
	} else {
		const uint2 l = meshLocal[pieceRec[g].x];
		o = groupBase[g >> 5];
		o.x += l.x;
		o.y += l.y;
	}
	meshOff[j] = o;
}

/* ---- final layout: vertex-buffer segmentation + draw-command table ---------------
 * The reference packs meshes into vertex buffers first-fit in submission order (allocVertices,
 * vg.cpp:5321-5342): mesh m joins the current buffer iff count + nv(m) <= max, else a new buffer and
 * a new draw command start with it.  That greedy rule is a chain  s -> next(s)  over mesh ids, with
 *   next(s) = the largest e > s such that  P[e] - P[s] <= max      (P = exclusive vertex offsets),
 * a pure function of s.  So: (1) next(s) for every mesh by binary search in P, in parallel;
 * (2) K rounds of pointer doubling J <- J o J, each of which also doubles the list of known
 * segment starts (starts[j + 2^k] = J_k[starts[j]]);  (3) one record per segment;  (4) per-mesh
 * index base = P[m] - P[start of its segment] (+ the vertices already in the continued command). */
struct FinalHdr
{
	uint32_t vbId0;    /* vertex buffer of segment 0                                   */
	uint32_t first0;   /* first vertex of segment 0 in its buffer                      */
	uint32_t cmdV0;    /* vertices already in the draw command segment 0 continues     */
	uint32_t oversize; /* a mesh with >= max vertices was seen                         */
	uint32_t numCmds;
	uint32_t pad[3];
};

__global__ void k_seg_next(const ulonglong2* __restrict__ meshOff, uint32_t M, uint32_t maxV, uint32_t vbId, uint32_t vbCount, uint32_t cmdV,
                           uint32_t* __restrict__ next, uint32_t* __restrict__ starts, FinalHdr* __restrict__ hdr)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s > M) return;
	if (s == M) {
		next[M] = M;
		return;
	}
	const uint64_t P0 = meshOff[s].x;
	if (meshOff[s + 1].x - P0 >= (uint64_t)maxV) atomicOr(&hdr->oversize, 1u); /* VG_CHECK at vg.cpp:5323 */
	uint64_t T = P0 + maxV;
	bool fresh = true; /* segment starts at vertex 0 of a new buffer */
	if (s == 0 && vbCount > 0) {
		/* the first mesh continues the context's current buffer if it fits */
		if (vbCount + (meshOff[1].x - P0) <= (uint64_t)maxV) {
			T = P0 + maxV - vbCount;
			fresh = false;
		}
	}
	/* largest e in (s, M] with P[e] <= T */
	uint32_t lo = s + 1, hi = M;
	while (lo < hi) {
		const uint32_t mid = lo + (hi - lo + 1) / 2;
		if (meshOff[mid].x <= T) lo = mid;
		else hi = mid - 1;
	}
	next[s] = lo;
	if (s == 0) {
		starts[0] = 0;
		hdr->vbId0 = fresh && vbCount > 0 ? vbId + 1 : vbId;
		hdr->first0 = fresh ? 0u : vbCount;
		hdr->cmdV0 = fresh ? 0u : cmdV;
	}
}

/* one doubling round: Jout = Jin o Jin, and starts[j + half] = Jin[starts[j]] for j < half */
__global__ void k_seg_double(const uint32_t* __restrict__ Jin, uint32_t* __restrict__ Jout, uint32_t M, uint32_t* __restrict__ starts, uint32_t half,
                             uint32_t writeJ)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s < half) starts[s + half] = Jin[starts[s]];
	if (writeJ && s <= M) Jout[s] = Jin[Jin[s]];
}

__global__ void k_seg_cmds(const uint32_t* __restrict__ starts, uint32_t numStarts, const ulonglong2* __restrict__ meshOff, uint32_t M, uint32_t ibCount,
                           const uint32_t* __restrict__ jobDraw, const uint32_t* __restrict__ drawScissor, FinalHdr* __restrict__ hdr,
                           vgb_draw_cmd* __restrict__ cmds)
{
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= numStarts) return;
	const uint32_t s = starts[j];
	if (s >= M) return;
	const uint32_t e = j + 1 < numStarts ? starts[j + 1] : M;
	const ulonglong2 a = meshOff[s], b = meshOff[e], o = meshOff[0];
	vgb_draw_cmd c;
	c.vb_id = hdr->vbId0 + j;
	c.first_vertex = j == 0 ? hdr->first0 : 0u;
	c.first_index = ibCount + (uint32_t)(a.y - o.y);
	c.num_vertices = (uint32_t)(b.x - a.x);
	c.num_indices = (uint32_t)(b.y - a.y);
	c.first_mesh = s;
	c.num_meshes = e - s;
	c.scissor_id = drawScissor != nullptr ? drawScissor[jobDraw[s]] & 0x7fffffffu : 0u;
	c.stream_vertex = a.x - o.x;
	c.stream_index = a.y - o.y;
	cmds[j] = c;
	if (e >= M) hdr->numCmds = j + 1;
}

/* index base of every mesh: vertices that precede it in its draw command (vg.cpp:5240) */
__global__ void k_seg_mesh_base(const uint32_t* __restrict__ starts, uint32_t numStarts, const ulonglong2* __restrict__ meshOff, uint32_t M,
                                const FinalHdr* __restrict__ hdr, uint32_t* __restrict__ meshBase)
{
	const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= M) return;
	uint32_t lo = 0, hi = numStarts - 1; /* last j with starts[j] <= m */
	while (lo < hi) {
		const uint32_t mid = lo + (hi - lo + 1) / 2;
		if (starts[mid] <= m) lo = mid;
		else hi = mid - 1;
	}
	meshBase[m] = (uint32_t)(meshOff[m].x - meshOff[starts[lo]].x) + (lo == 0 ? hdr->cmdV0 : 0u);
}

/* ---- scissor breaks inside a vertex buffer (m_ForceNewDrawCommand, vg.cpp:3946-4033 + 5368-5402) ----
 * A draw recorded after a scissor change carries a mark; the flag stays up until a mesh with vertices consumes it
 * (allocDrawCommand is only reached with vertices), and that mesh opens a new draw command in the SAME vertex
 * buffer.  With  X[m] = (marks before m, non-empty meshes before m)  a non-empty mesh m opens a command iff there
 * is a mark at or before it and no non-empty mesh between the latest such mark and m.  Draw commands are then the
 * union of the vertex-buffer starts (the chain above) and those meshes; two more scans number them. */
__global__ void k_split_marks(const ulonglong2* __restrict__ meshOff, uint32_t M, const uint32_t* __restrict__ jobDraw,
                              const uint32_t* __restrict__ drawScissor, uint2* __restrict__ in)
{
	const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= M) return;
	const uint32_t d = jobDraw[m];
	const uint32_t mark = (drawScissor[d] >> 31) && (m == 0 || jobDraw[m - 1] != d) ? 1u : 0u; /* first mesh of a marked draw */
	in[m] = make_uint2(mark, meshOff[m + 1].x > meshOff[m].x ? 1u : 0u);
}

__global__ void k_split_pos(const uint2* __restrict__ in, const ulonglong2* __restrict__ X, uint32_t M, uint32_t* __restrict__ markPos)
{
	const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m < M && in[m].x) markPos[X[m].x] = m;
}

/* flags[m] = (m starts a vertex buffer, m starts a draw command); overwrites the (mark, non-empty) pairs in place */
__global__ void k_split_flags(uint2* __restrict__ io, const ulonglong2* __restrict__ X, uint32_t M, const uint32_t* __restrict__ markPos,
                              const uint32_t* __restrict__ starts, uint32_t numStarts)
{
	const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= M) return;
	const uint2 f = io[m];
	uint32_t lo = 0, hi = numStarts - 1; /* last j with starts[j] <= m */
	while (lo < hi) {
		const uint32_t mid = lo + (hi - lo + 1) / 2;
		if (starts[mid] <= m) lo = mid;
		else hi = mid - 1;
	}
	const uint32_t vb = starts[lo] == m ? 1u : 0u;
	uint32_t sc = 0;
	const uint64_t k = X[m].x + f.x; /* marks at or before m */
	if (f.y && k > 0) sc = X[m].y == X[markPos[k - 1]].y ? 1u : 0u;
	io[m] = make_uint2(vb, vb | sc);
}

__global__ void k_split_cmds(const uint2* __restrict__ flags, const ulonglong2* __restrict__ Y, uint32_t M, const uint32_t* __restrict__ starts,
                             const ulonglong2* __restrict__ meshOff, uint32_t ibCount, const uint32_t* __restrict__ jobDraw,
                             const uint32_t* __restrict__ drawScissor, FinalHdr* __restrict__ hdr, vgb_draw_cmd* __restrict__ cmds,
                             uint32_t* __restrict__ cmdMesh)
{
	const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m == 0) hdr->numCmds = (uint32_t)Y[M].y;
	if (m >= M || !flags[m].y) return;
	const uint32_t c = (uint32_t)Y[m].y;
	const uint32_t j = (uint32_t)(Y[m].x + flags[m].x) - 1u; /* vertex buffer (= chain segment) of m */
	const ulonglong2 a = meshOff[m], o = meshOff[0];
	vgb_draw_cmd r;
	r.vb_id = hdr->vbId0 + j;
	r.first_vertex = (uint32_t)(a.x - meshOff[starts[j]].x) + (j == 0 ? hdr->first0 : 0u);
	r.first_index = ibCount + (uint32_t)(a.y - o.y);
	r.num_vertices = 0;
	r.num_indices = 0;
	r.first_mesh = m;
	r.num_meshes = 0;
	r.scissor_id = drawScissor[jobDraw[m]] & 0x7fffffffu;
	r.stream_vertex = a.x - o.x;
	r.stream_index = a.y - o.y;
	cmds[c] = r;
	cmdMesh[c] = m;
}

/* sizes of every command (up to the next command's first mesh) and the index base of every mesh */
__global__ void k_split_fin(const uint2* __restrict__ flags, const ulonglong2* __restrict__ Y, uint32_t M, const ulonglong2* __restrict__ meshOff,
                            const FinalHdr* __restrict__ hdr, const uint32_t* __restrict__ cmdMesh, vgb_draw_cmd* __restrict__ cmds,
                            uint32_t* __restrict__ meshBase)
{
	const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t numCmds = (uint32_t)Y[M].y;
	if (m < numCmds) {
		const uint32_t s = cmdMesh[m], e = m + 1 < numCmds ? cmdMesh[m + 1] : M;
		cmds[m].num_vertices = (uint32_t)(meshOff[e].x - meshOff[s].x);
		cmds[m].num_indices = (uint32_t)(meshOff[e].y - meshOff[s].y);
		cmds[m].num_meshes = e - s;
	}
	if (m >= M) return;
	const uint32_t c = (uint32_t)(Y[m].y + flags[m].y) - 1u; /* mesh 0 always starts a command */
	/* command 0 continues the context's open command (cmdV0 vertices in it) unless mesh 0 itself was forced */
	meshBase[m] = (uint32_t)(meshOff[m].x - meshOff[cmdMesh[c]].x) + (c == 0 ? hdr->cmdV0 : 0u);
}

/* the uv stream: every vertex of a colour mesh samples the white pixel (vg.cpp:5220-5227) */
__global__ void k_fill_u32(uint32_t* __restrict__ dst, uint64_t n, uint32_t v)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	uint4* d4 = reinterpret_cast<uint4*>(dst);
	const uint64_t n4 = n / 4;
	const uint4 v4 = make_uint4(v, v, v, v);
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) d4[i] = v4;
	for (uint64_t i = n4 * 4 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = v;
}

/* ---- instance replay (submitCachedMesh, vg.cpp:6137-6160) -------------------------
 * Pure streaming: a thread owns one cached vertex and writes it for a chunk of instances (the source is
 * read once per chunk, the six matrix entries are warp-uniform loads); stores of a warp are consecutive
 * in every instance.  transformPos2D keeps the reference's scalar association (vg_util.h:24-28). */
static constexpr uint32_t REPLAY_CHUNK = 8; /* instances per block row */

__global__ void k_replay_vertices(const float2* __restrict__ pos, const uint32_t* __restrict__ col, uint64_t V, const float* __restrict__ xf, uint32_t n,
                                  const float* __restrict__ inv, float2* __restrict__ outPos, uint32_t* __restrict__ outCol)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= V) return;
	float2 p = pos[i];
	const uint32_t c = col[i];
	if (inv != nullptr) {
		/* what the cache holds: the mesh in the command list's local space, rounded to float */
		const float x = inv[0] * p.x + inv[2] * p.y + inv[4];
		const float y = inv[1] * p.x + inv[3] * p.y + inv[5];
		p = make_float2(x, y);
	}
	for (uint32_t k0 = blockIdx.y * REPLAY_CHUNK; k0 < n; k0 += gridDim.y * REPLAY_CHUNK) {
		const uint32_t k1 = min(n, k0 + REPLAY_CHUNK);
		for (uint32_t k = k0; k < k1; ++k) {
			const float* m = xf + 6 * (size_t)k;
			const float x = __ldg(m + 0) * p.x + __ldg(m + 2) * p.y + __ldg(m + 4);
			const float y = __ldg(m + 1) * p.x + __ldg(m + 3) * p.y + __ldg(m + 5);
			__stcs(outPos + (uint64_t)k * V + i, make_float2(x, y));
			__stcs(outCol + (uint64_t)k * V + i, c);
		}
	}
}

/* indices: a thread owns one cached word and writes it for a chunk of instances -- one coalesced load, then
 * fully coalesced stores in every copy.  W = uint32 merged index (+ k*V per copy), uint16 local index, or a
 * PAIR of uint16 local indices moved as one 32-bit word (when the cached stream has an even length, so that
 * every copy starts 4-byte aligned) */
template<class W, bool ADD>
__global__ void k_replay_indices(const W* __restrict__ src, uint64_t I, uint64_t V, uint32_t n, W* __restrict__ out)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= I) return;
	const W v = src[i];
	for (uint32_t k0 = blockIdx.y * REPLAY_CHUNK; k0 < n; k0 += gridDim.y * REPLAY_CHUNK) {
		const uint32_t k1 = min(n, k0 + REPLAY_CHUNK);
		for (uint32_t k = k0; k < k1; ++k) {
			__stcs(out + (uint64_t)k * I + i, ADD ? (W)(v + (W)(k * V)) : v);
		}
	}
}

/* ---- the polyline window of a warp: 32 consecutive pieces = (mostly) 32 consecutive points ----------
 * Every lane loads ONLY its own point; its neighbours come from the adjacent lanes by shuffle when
 * those hold the neighbouring piece of the same job.  What shuffles cannot give: the point before
 * lane 0's and after lane 31's (loaded speculatively together with the lane's own point: `edge`) and
 * the wrap-around neighbours of a closed polyline's first / last piece (a dependent load, rare). */
struct Window
{
	float2 pm, pn;
	bool linkedPrev; /* lane-1 holds this piece's predecessor */
};

__device__ __forceinline__ float2 shflUp2(float2 v)
{
	return make_float2(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1));
}
__device__ __forceinline__ float2 shflDown2(float2 v)
{
	return make_float2(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1));
}
__device__ __forceinline__ float2 shflIdx2(float2 v, int src)
{
	return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

/* the speculative edge load: lane 0's previous point, lane 31's next point (same address as the own
 * point for every other lane, so it costs one predicated load) */
__device__ __forceinline__ float2 loadEdge(const float2* __restrict__ pts, uint32_t pt, float2 p, uint32_t lane)
{
	float2 e = p;
	if (lane == 0 || lane == 31) {
		e = pts[lane == 0 ? (pt ? pt - 1 : 0u) : pt + 1];
	}
	return e;
}

/* all 32 lanes must call this (shuffles); results are meaningful for valid stroke pieces only */
__device__ __forceinline__ Window windowOf(const float2* __restrict__ pts, const JobCtx& J, uint32_t j, uint32_t q, bool valid, float2 p, float2 edge,
                                           uint32_t lane)
{
	const uint32_t jm = __shfl_up_sync(0xffffffffu, j, 1), jn = __shfl_down_sync(0xffffffffu, j, 1);
	const unsigned vmask = __ballot_sync(0xffffffffu, valid);
	Window w;
	w.linkedPrev = valid && lane > 0 && jm == j;
	const bool linkedNext = valid && lane < 31 && ((vmask >> (lane + 1)) & 1u) && jn == j;
	const float2 pmS = shflUp2(p), pnS = shflDown2(p);
	const uint32_t N = J.N;
	w.pm = w.linkedPrev ? pmS : ((lane == 0 && q > 0) ? edge : p);
	w.pn = linkedNext ? pnS : ((lane == 31 && q + 1 < N) ? edge : p);
	if (valid && J.closed) {
		const float2* v = pts + J.first;
		if (q == 0) w.pm = v[N - 1];
		if (q == N - 1) w.pn = v[0];
	}
	return w;
}

/* AA-stroke pieces of a whole warp: each lane computes ONE segment direction (its outgoing one) and takes
 * the incoming one from the lane before it.  The direction nobody owns -- into lane 0's point, or into a
 * closed polyline's first point -- is computed by a lane that has no outgoing segment of its own (an end
 * cap, or a lane past the end); only when the warp has no such lane, or more than one orphan, do the
 * orphans compute theirs separately.  Bit-identical to two vdir() per lane: same inputs, same function. */
template<int JT, int CT>
__device__ __forceinline__ Piece classifyWarpAA(const float2* __restrict__ pts, const JobCtx& J, uint32_t j, uint32_t q, bool valid, float2 p, float2 edge,
                                                uint32_t lane)
{
	const Window w = windowOf(pts, J, j, q, valid, p, edge, lane);
	const bool open = !J.closed;
	const bool needA = valid && !(open && q == 0);
	const bool needB = valid && !(open && q == J.N - 1);
	const bool orphan = needA && !w.linkedPrev;
	const unsigned om = __ballot_sync(0xffffffffu, orphan), fm = __ballot_sync(0xffffffffu, !needB);
	const bool have = om != 0 && fm != 0; /* warp-uniform */
	const int o = __ffs(om) - 1, f = __ffs(fm) - 1;
	float2 x0 = p, x1 = w.pn;
	if (have) {
		const float2 a = shflIdx2(w.pm, o), b = shflIdx2(p, o);
		if ((int)lane == f) {
			x0 = a;
			x1 = b;
		}
	}
	const float2 d = vdir(x0, x1);
	float2 dA = shflUp2(d);
	if (have) {
		const float2 t = shflIdx2(d, f);
		if ((int)lane == o) dA = t;
	}
	if (orphan && !(have && (int)lane == o)) {
		dA = vdir(w.pm, p);
	}
	return classifyStrokeAA<JT, CT>(J, q, p, dA, d);
}

/* each warp counts GC consecutive groups; the loads of all of them are issued before any is used
 * (the kernel is a chain of dependent loads: piece -> job record -> points) */
#ifndef VGB_GC
#define VGB_GC 2
#endif
#ifndef VGB_COUNT_MIN_BLOCKS
#define VGB_COUNT_MIN_BLOCKS 12
#endif
static constexpr uint32_t GC = VGB_GC;

template<uint32_t OPS, int JT, int CT>
__global__ void __launch_bounds__(128, VGB_COUNT_MIN_BLOCKS) k_group_count(PieceArgs a)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t w0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * GC;
	uint32_t g[GC], ps[GC];
	uint2 rec[GC];
	bool valid[GC];
	JobCtx J[GC];
	float2 pc[GC], pe[GC];
#pragma unroll
	for (uint32_t k = 0; k < GC; ++k) {
		g[k] = (w0 + k) * 32 + lane;
		valid[k] = g[k] < a.numPieces;
		rec[k] = valid[k] ? a.pieceRec[g[k]] : make_uint2(0u, 0u);
	}
#pragma unroll
	for (uint32_t k = 0; k < GC; ++k) {
		ps[k] = 0;
		J[k] = loadJob(a, rec[k].x, &ps[k]);
		pc[k] = a.pts[rec[k].y];
		pe[k] = loadEdge(a.pts, rec[k].y, pc[k], lane);
	}
#pragma unroll
	for (uint32_t k = 0; k < GC; ++k) {
		const uint32_t w = w0 + k;
		const uint32_t q = g[k] - ps[k];
		const bool aa = (OPS & OPS_STROKE_AA) && (!(OPS & ~OPS_STROKE_AA) || J[k].op == VGB_OP_STROKE_AA);
		Window win;
		if (OPS & OPS_STROKE_AA) {
			win = windowOf(a.pts, J[k], rec[k].x, q, valid[k] && aa, pc[k], pe[k], lane);
		}
		uint32_t vc = 0, ic = 0, tail = 0;
		if (valid[k]) {
			Piece P;
			if (!(aa && quickCountAA<JT, CT>(J[k], q, win.pm, pc[k], win.pn, lane == 31, P))) {
				P = classify<OPS, JT, CT>(J[k], q, a.pts);
			}
			vc = P.vcount;
			ic = P.icount;
			tail = packTail(P);
		}
		/* within-group offsets: the lane owning a mesh's first piece records where that mesh starts
		 * inside its group; k_mesh_off adds the group base after the scan (no atomics, no per-mesh scan) */
		const uint2 inc = warpInclusive(make_uint2(vc, ic), lane);
		if (valid[k] && g[k] == ps[k]) {
			a.meshTotals[rec[k].x] = make_uint2(inc.x - vc, inc.y - ic);
		}
		const uint32_t tv = __shfl_sync(0xffffffffu, inc.x, 31);
		const uint32_t ti = __shfl_sync(0xffffffffu, inc.y, 31);
		if (w * 32 < a.numPieces) {
			if (lane == 0) {
				a.blockTotals[w] = make_uint2(tv, ti);
			}
			if (lane == 31) {
				a.blockTail[w] = tail;
			}
		}
	}
}

/* BIGI: index image sized for ~40 indices per piece (uint32 output of fat pieces); otherwise half of
 * that for uint32, so that thin meshes keep 9 blocks per SM.  uint16 always uses the full count. */
/* Resident blocks per SM: the (cap, join)-specialised AA-stroke kernels run best at 8 (64 registers), the general
 * mixed-op kernel at 6 (80 registers, no spills; its uint32 image limits it to 6 blocks of shared memory anyway --
 * tiger-like scene x 600, emit: 2.96 ms at 6, 3.29 at 7, 3.71 at 8, 3.34 at 9 with 72 bytes of spills). */
template<class IDX, uint32_t OPS, int JT, int CT, bool BIGI>
__global__ void __launch_bounds__(128, (OPS == OPS_STROKE_AA && JT < 3) ? VGB_EMIT_MIN_BLOCKS_SPEC : VGB_EMIT_MIN_BLOCKS) k_group_emit(PieceArgs a)
{
	constexpr uint32_t CIW = (sizeof(IDX) == 2 || BIGI) ? CIW_COUNT : CIW_COUNT / 2;
	constexpr uint32_t KI = 16 / sizeof(IDX);
	constexpr uint32_t POS_BYTES = (CVW + 2) * 8, COL_BYTES = (CVW + 4) * 4, IDX_BYTES = (CIW + KI) * sizeof(IDX);
	constexpr uint32_t WBYTES = POS_BYTES + COL_BYTES + IDX_BYTES;
	static_assert(POS_BYTES % 16 == 0 && COL_BYTES % 16 == 0 && IDX_BYTES % 16 == 0, "16-byte aligned staging streams");
	__shared__ __align__(16) uint8_t sStage[4 * WBYTES];

	/* PERSISTENT warps: warp w of the grid takes groups w, w + W, w + 2W ...  What the NEXT group needs is
	 * fetched while the current one is being emitted: its piece records travel in registers two groups
	 * ahead (recB), and with the records of the group one ahead (recA) in hand the lines of its points, job
	 * records, group base, hand-over word and mesh offsets are pulled into L2 -- so the dependent chain
	 * piece record -> {job record, point} that every group used to wait for at its head is an L2 hit.
	 * The wait for the previous group's bulk copy (its reads of the staging image) sits right before
	 * the image is written again, i.e. the copy-out of group t overlaps the classification of t + 1. */
	const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const uint32_t W = gridDim.x * 4; /* warps of the grid = statically assigned chunks */
	const uint32_t numGroups = (a.numPieces + 31) >> 5;
	/* groups are handed out in CHUNKS of TC consecutive ones: the first chunk of every warp is static, the rest
	 * are tickets (one same-address atomic per TC groups: at one per group the ~0.7 M atomics of a launch ran
	 * into the L2's same-address atomic rate on some boards and doubled the kernel time) */
	const uint32_t TC = 1u << a.chunkShift; /* 4 for large batches, 1 for small ones (all warps busy at once) */
	uint32_t w = (blockIdx.x * 4 + wib) * TC;
	if (w >= numGroups) {
		return; /* whole warp out of range */
	}
	uint8_t* stage = sStage + wib * WBYTES;
	float2* sPos = reinterpret_cast<float2*>(stage);
	uint32_t* sCol = reinterpret_cast<uint32_t*>(stage + POS_BYTES);
	IDX* sIdx = reinterpret_cast<IDX*>(stage + POS_BYTES + COL_BYTES);
	bool copyPending = false;
	uint2 recA = make_uint2(0u, 0u), recB = make_uint2(0u, 0u);
	uint32_t w1 = w + 1, w2 = w + 2; /* TC = 1: fixed up below */
	uint32_t nextChunk = 0; /* lane 0: the prefetched ticket (valid in lane 0 only until it is consumed) */
	{
		if (TC == 1) {
			/* every group is its own chunk: the next two groups are tickets already */
			uint32_t t = 0;
			if (lane == 0) t = atomicAdd(a.ticket, 2u);
			t = __shfl_sync(0xffffffffu, t, 0);
			w1 = W + t;
			w2 = W + t + 1;
		}
		if (lane == 0) nextChunk = atomicAdd(a.ticket, 1u);
		const uint32_t g0 = w * 32 + lane;
		if (g0 < a.numPieces) recA = a.pieceRec[g0];
		const uint32_t g1 = w1 * 32 + lane;
		if (w1 < numGroups && g1 < a.numPieces) recB = a.pieceRec[g1];
	}
	for (;;) {
	const uint32_t g = w * 32 + lane;
	const bool valid = g < a.numPieces;
	const uint2 rec = recA;
	const uint32_t j = rec.x;
	recA = recB;
	{
		/* the records of the group two ahead, L2 prefetch for the group one ahead */
		const uint32_t g2 = w2 * 32 + lane;
		recB = make_uint2(0u, 0u);
		if (w2 < numGroups && g2 < a.numPieces) recB = a.pieceRec[g2];
#if VGB_EMIT_PREFETCH
		if (w1 < numGroups) {
			if (w1 * 32 + lane < a.numPieces) {
				asm volatile("prefetch.global.L2 [%0];" ::"l"(a.pts + recA.y));
				asm volatile("prefetch.global.L2 [%0];" ::"l"(a.jobRec + recA.x));
				asm volatile("prefetch.global.L2 [%0];" ::"l"(a.meshOff + recA.x));
			}
			if (lane == 0) {
				asm volatile("prefetch.global.L2 [%0];" ::"l"(a.blockBase + w1));
				asm volatile("prefetch.global.L2 [%0];" ::"l"(a.blockTail + (w1 - 1)));
			}
		}
#endif
	}
#if VGB_EMIT_PREFETCH
	/* the group base, the previous group's hand-over word and the mesh offsets are only needed after the
	 * pieces are classified: start those loads now (into L1, no registers held) */
	if (lane == 0) {
		asm volatile("prefetch.global.L1 [%0];" ::"l"(a.blockBase + w));
		if (w > 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.blockTail + (w - 1)));
	}
	if (valid) {
		asm volatile("prefetch.global.L1 [%0];" ::"l"(a.meshOff + j));
	}
#endif
	JobCtx J;
	Piece P;
	uint32_t q = 0;
	if (OPS == OPS_STROKE_AA) {
		/* pure AA-stroke batch: one point load per lane, neighbours and segment directions by shuffle */
		uint32_t ps;
		const float2 p = a.pts[rec.y];
		const float2 edge = loadEdge(a.pts, rec.y, p, lane);
		J = loadJob(a, j, &ps);
		q = g - ps;
		P = classifyWarpAA<JT, CT>(a.pts, J, j, q, valid, p, edge, lane);
		if (!valid) {
			P.kind = PK_NONE;
			P.vcount = P.icount = 0;
			P.flags = 0;
		}
	} else {
		P.kind = PK_NONE;
		P.vcount = P.icount = 0;
		P.flags = 0;
		if (valid) {
			uint32_t ps;
			J = loadJob(a, j, &ps);
			q = g - ps;
			P = classify<OPS, JT, CT>(J, q, a.pts);
		}
	}
	const uint32_t myTail = valid ? packTail(P) : 0u;
	uint32_t tw = __shfl_up_sync(0xffffffffu, myTail, 1);
	if (lane == 0) {
		tw = w > 0 ? a.blockTail[w - 1] : 0u;
	}
	const uint2 inc = warpInclusive(make_uint2(P.vcount, P.icount), lane);
	const uint32_t vl = inc.x - P.vcount, il = inc.y - P.icount;
	const uint32_t totV = __shfl_sync(0xffffffffu, inc.x, 31), totI = __shfl_sync(0xffffffffu, inc.y, 31);
	const ulonglong2 bb = a.blockBase[w];
	uint32_t F = 0;
	ulonglong2 mo = make_ulonglong2(0, 0);
	uint4 prev = make_uint4(0, 0, 0, 0);
	if (valid) {
		mo = a.meshOff[j];
		F = (uint32_t)(bb.x + vl - mo.x);
		if (P.flags & PF_HAS_PREV) {
			prev = ((OPS & OPS_OTHER) && J.op != VGB_OP_STROKE_AA) ? ribbonOutFromTail(J, tw, F)
			                                                        : out4FromTail(CT < 3 ? (uint32_t)CT : J.cap, JT < 3 ? (uint32_t)JT : J.join, tw, F);
		}
	}
	/* uint16: BIGI doubles as "final layout" (per-mesh base from the draw-command segmentation);
	 * the plain per-mesh-local variant keeps base = 0 at compile time */
	uint32_t idxBase = sizeof(IDX) == 2 ? 0u : (uint32_t)(mo.x + a.vertexBase + (a.vertexBasePtr ? __ldg(a.vertexBasePtr) : 0ull));
	if (sizeof(IDX) == 2 && BIGI && valid) idxBase = a.meshBase[j];
	const bool fits = totV <= CVW && totI <= CIW; /* warp-uniform */
	if (fits) {
		/* the staging image: typed shared-memory stores, packed index words */
		if (copyPending) {
			/* the previous group's bulk copies must have READ the image before it is overwritten */
			if (lane == 0 || lane == 2 || lane == 8) {
				bulkWaitRead();
			}
			__syncwarp();
		}
		SinkS<IDX, !(OPS & OPS_OTHER)> S;
		S.base = idxBase;
		S.pos = sPos + (uint32_t)(bb.x & 1) + vl;
		S.col = sCol + (uint32_t)(bb.x & 3) + vl;
		S.idx = sIdx + (uint32_t)(bb.y & (KI - 1)) + il;
		S.par = ((uint32_t)bb.y + il) & 1u;
		if (valid) {
			emitAny<OPS, JT, CT>(a.pts, J, P, F, prev, S);
		}
		fenceProxyAsync();
		__syncwarp();
		copyOutWarp3<IDX>(a.pos, a.col, reinterpret_cast<IDX*>(a.idx), bb.x, bb.y, totV, totI, sPos, sCol, sIdx, lane);
		copyPending = true;
	} else if (valid) {
		/* a warp whose pieces exceed the staging image writes straight to HBM */
		Sink<IDX, !(OPS & OPS_OTHER)> S;
		S.base = idxBase;
		S.par = 0;
		S.pos = a.pos + bb.x + vl;
		S.col = a.col + bb.x + vl;
		S.idx = reinterpret_cast<IDX*>(a.idx) + bb.y + il;
		emitAnyCold<OPS, JT, CT>(a.pts, J, P, F, prev, S);
	}
	/* next group: tickets keep the groups in flight a contiguous window of the output whatever the warps' pace */
	w = w1;
	w1 = w2;
	if (((w2 + 1) & (TC - 1)) != 0) {
		++w2;
	} else {
		/* chunk boundary: consume the prefetched ticket (it was issued TC groups ago) and ask for the next */
		w2 = (W + __shfl_sync(0xffffffffu, nextChunk, 0)) << a.chunkShift;
		if (lane == 0) nextChunk = atomicAdd(a.ticket, 1u);
	}
	if (w >= numGroups) {
		break;
	}
	} /* groups of this warp */
	if (copyPending && (lane == 0 || lane == 2 || lane == 8)) {
		bulkWaitRead(); /* the image must outlive its bulk copies */
	}
}

/* ---- single-pass emit: classification, offsets and output in ONE kernel ---------------------------
 * (experimental: VGB_SINGLE_PASS=1; results are bit-identical to the two-pass pipeline, see profiles/README.md for
 * where it stands on speed)
 * k_group_count + the two scans + k_mesh_off compute, for every group of 32 pieces, exactly what k_group_emit
 * computes again right before it writes: the pieces' sizes.  This variant classifies once.  A group's output
 * offset is the sum of the sizes of all groups before it, which a chained scan with decoupled look-back
 * (Merrill & Garland) delivers inside the kernel: every group publishes its totals as soon as its pieces are
 * classified (descriptor word A), then sums its predecessors' descriptors backwards until it meets one that
 * already knows its own inclusive prefix (word P), and publishes P itself -- before it emits anything.
 * Two more things used to come from the count pass and travel through the descriptors as well:
 *   - the hand-over word of the previous group's last piece (lane 0's `prev` vertices),
 *   - F, the number of vertices the lane's mesh already holds: for a job that began in an earlier group this is
 *     the job's share of the group it began in plus the whole of every group in between (`spans`), i.e. a second,
 *     segmented sum over the same descriptors.
 * The mesh table is written by the lanes that hold a job's first piece.  The output buffers must exist before the
 * sizes are known: the kernel is given their capacity, a group that would not fit writes nothing and raises a
 * flag (the host grows the buffers to the totals the scan produced anyway, and runs the kernel again).
 * Waiting is bounded: a look-back that polls one descriptor SP_POLL_LIMIT times gives up, raises a flag that stops
 * every other look-back, and the host runs the batch through the two-pass pipeline. */
/* A descriptor is TWO self-validating 16-byte words, each written with one aligned 128-bit store and read with one
 * 128-bit load (what cub::ScanTileState does for its status + value words): no fence on either side, a reader either
 * sees a whole word or the zeros the host left there.
 *   word A (totals known):   x = vertices | 0x5 << 28, y = indices, z = last piece's hand-over word,
 *                            w = the last job's vertices inside the group | spans << 31
 *   word P (prefix known):   x = prefix vertices (low 32), y = high 16 of both prefixes, z = prefix indices (low 32),
 *                            w = the last job's vertices up to the end of the group | 0xA << 28 */
__device__ __forceinline__ void spStore(uint4* p, uint4 v)
{
	asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 spLoad(const uint4* p)
{
	uint4 v;
	asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ uint4 spWordA(uint32_t totV, uint32_t totI, uint32_t tail, uint32_t lastLocal, uint32_t spans)
{
	return make_uint4((totV & 0x0FFFFFFFu) | 0x50000000u, totI, tail, (lastLocal & 0x7FFFFFFFu) | (spans << 31));
}
__device__ __forceinline__ uint4 spWordP(uint64_t pv, uint64_t pi, uint32_t carry)
{
	return make_uint4((uint32_t)pv, (uint32_t)((pv >> 32) & 0xFFFFu) | ((uint32_t)((pi >> 32) & 0xFFFFu) << 16), (uint32_t)pi,
	                  (carry & 0x0FFFFFFFu) | 0xA0000000u);
}
__device__ __forceinline__ uint64_t warpSum64(uint64_t v)
{
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		v += __shfl_xor_sync(0xffffffffu, v, d);
	}
	return v;
}

static constexpr uint32_t SP_POLL_LIMIT = 1u << 22; /* polls of one descriptor before the look-back gives up */

struct SpLook
{
	uint64_t exV, exI; /* exclusive prefix of the group */
	uint32_t carry;    /* vertices lane 0's job holds before this group (if that job began earlier) */
	uint32_t tail;     /* hand-over word of the previous group's last piece */
	bool dead;
};

/* all 32 lanes; w > 0.  Lane l looks at group base - l; the window moves back 32 groups at a time. */
__device__ __forceinline__ SpLook spLookBack(const PieceArgs& a, uint32_t w, uint32_t lane)
{
	SpLook r;
	r.exV = r.exI = 0;
	r.carry = r.tail = 0;
	r.dead = false;
	bool carryDone = false, first = true;
	int64_t base = (int64_t)w - 1;
	for (;;) {
		const int64_t u = base - (int64_t)lane;
		/* before group 0 everything is zero and known */
		uint4 wa = make_uint4(0x50000000u, 0u, 0u, 0u), wp = make_uint4(0u, 0u, 0u, 0xA0000000u);
		bool gaveUp = false;
		if (u >= 0) {
			uint32_t polls = 0;
			for (;;) {
				wa = spLoad(a.spA + u);
				if ((wa.x >> 28) == 0x5u) break;
				__nanosleep(20);
				if (++polls >= SP_POLL_LIMIT || ((polls & 1023u) == 0 && (*(volatile uint32_t*)a.spFlags & 2u))) {
					gaveUp = true;
					break;
				}
			}
			wp = spLoad(a.spP + u);
		}
		if (__any_sync(0xffffffffu, gaveUp)) {
			if (lane == 0) atomicOr(a.spFlags, 2u);
			r.dead = true;
			return r;
		}
		const bool res = (wp.w >> 28) == 0xAu;
		const unsigned pm = __ballot_sync(0xffffffffu, res);
		const uint32_t np = pm ? (uint32_t)__ffs(pm) - 1u : 32u; /* nearest resolved predecessor of this window */
		uint64_t cv = 0, ci = 0;
		if (lane < np) {
			cv = wa.x & 0x0FFFFFFFu;
			ci = wa.y;
		} else if (lane == np) {
			cv = (uint64_t)wp.x | ((uint64_t)(wp.y & 0xFFFFu) << 32);
			ci = (uint64_t)wp.z | ((uint64_t)(wp.y >> 16) << 32);
		}
		r.exV += warpSum64(cv);
		r.exI += warpSum64(ci);
		if (!carryDone) {
			/* the walk over the job that reaches into this group ends at the first group it does not span */
			const bool stop = res || !(wa.w >> 31);
			const unsigned sm = __ballot_sync(0xffffffffu, stop);
			const uint32_t k = sm ? (uint32_t)__ffs(sm) - 1u : 32u;
			uint32_t c = 0;
			if (lane < k) c = wa.w & 0x7FFFFFFFu;
			else if (lane == k) c = res ? (wp.w & 0x0FFFFFFFu) : (wa.w & 0x7FFFFFFFu);
			r.carry += __reduce_add_sync(0xffffffffu, c);
			carryDone = k < 32u;
		}
		if (first) {
			r.tail = __shfl_sync(0xffffffffu, wa.z, 0);
			first = false;
		}
		if (np < 32u) break;
		base -= 32;
	}
	return r;
}

template<class IDX, uint32_t OPS, int JT, int CT, bool BIGI>
__global__ void __launch_bounds__(128, (OPS == OPS_STROKE_AA && JT < 3) ? VGB_EMIT_MIN_BLOCKS_SPEC : VGB_EMIT_MIN_BLOCKS) k_group_emit_sp(PieceArgs a)
{
	constexpr uint32_t CIW = (sizeof(IDX) == 2 || BIGI) ? CIW_COUNT : CIW_COUNT / 2;
	constexpr uint32_t KI = 16 / sizeof(IDX);
	constexpr uint32_t POS_BYTES = (CVW + 2) * 8, COL_BYTES = (CVW + 4) * 4, IDX_BYTES = (CIW + KI) * sizeof(IDX);
	constexpr uint32_t WBYTES = POS_BYTES + COL_BYTES + IDX_BYTES;
	static_assert(POS_BYTES % 16 == 0 && COL_BYTES % 16 == 0 && IDX_BYTES % 16 == 0, "16-byte aligned staging streams");
	__shared__ __align__(16) uint8_t sStage[4 * WBYTES];

	const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const uint32_t numGroups = (a.numPieces + 31) >> 5;
	/* group order: a warp's first group is its index in the grid, every later one a ticket drawn ONE group ahead (at the
	 * top of the group before it).  What keeps the look-back decoupled is that every group publishes its totals a
	 * roughly fixed time after its id became known to be taken: immediately for the first groups, one group later for
	 * ticketed ones.  (Handing a warp several consecutive groups at once breaks exactly that: see launchEmitSp.) */
	const uint32_t W = gridDim.x * 4;
	uint32_t w = blockIdx.x * 4 + wib;
	if (w >= numGroups) {
		return;
	}
	uint8_t* stage = sStage + wib * WBYTES;
	float2* sPos = reinterpret_cast<float2*>(stage);
	uint32_t* sCol = reinterpret_cast<uint32_t*>(stage + POS_BYTES);
	IDX* sIdx = reinterpret_cast<IDX*>(stage + POS_BYTES + COL_BYTES);
	bool copyPending = false;
	uint2 rec = make_uint2(0u, 0u);
	if (w * 32 + lane < a.numPieces) rec = a.pieceRec[w * 32 + lane];
	for (;;) {
	const uint32_t g = w * 32 + lane;
	const bool valid = g < a.numPieces;
	const uint32_t j = rec.x;
	/* chunkShift doubles as the ticket policy here: 0 = drawn one group ahead (its piece records are prefetched, but the
	 * id is taken a whole group time before its totals appear), 1 = drawn when the group starts (nothing prefetched) */
	const bool lateTicket = a.chunkShift != 0;
	uint32_t ticketNext = 0;
	if (!lateTicket && lane == 0) ticketNext = atomicAdd(a.ticket, 1u); /* consumed after the totals are published */
	JobCtx J;
	Piece P;
	uint32_t q = 0, ps = 0;
	if (OPS == OPS_STROKE_AA) {
		const float2 p = a.pts[rec.y];
		const float2 edge = loadEdge(a.pts, rec.y, p, lane);
		J = loadJob(a, j, &ps);
		q = g - ps;
		P = classifyWarpAA<JT, CT>(a.pts, J, j, q, valid, p, edge, lane);
		if (!valid) {
			P.kind = PK_NONE;
			P.vcount = P.icount = 0;
			P.flags = 0;
		}
	} else {
		P.kind = PK_NONE;
		P.vcount = P.icount = 0;
		P.flags = 0;
		if (valid) {
			J = loadJob(a, j, &ps);
			q = g - ps;
			P = classify<OPS, JT, CT>(J, q, a.pts);
		}
	}
	const uint32_t myTail = valid ? packTail(P) : 0u;
	const uint2 inc = warpInclusive(make_uint2(P.vcount, P.icount), lane);
	const uint32_t vl = inc.x - P.vcount, il = inc.y - P.icount;
	const uint32_t totV = __shfl_sync(0xffffffffu, inc.x, 31), totI = __shfl_sync(0xffffffffu, inc.y, 31);
	/* where the lane's job begins: a lane of this group (>= 0) or an earlier group (-1) */
	const uint32_t g0 = w * 32;
	const bool inside = valid && ps >= g0;
	const uint32_t vlAtStart = __shfl_sync(0xffffffffu, vl, inside ? (int)(ps - g0) : 0);
	const uint32_t ownV = inc.x - (inside ? vlAtStart : 0u); /* the job's vertices in this group up to and including the lane */
	/* descriptor fields: the last piece's hand-over word; the last job's share of this group; does it span the group */
	const uint32_t tail31 = __shfl_sync(0xffffffffu, myTail, 31);
	const uint32_t lastLocal = __shfl_sync(0xffffffffu, valid ? ownV : 0u, 31);
	const uint32_t spans = __shfl_sync(0xffffffffu, (valid && !inside) ? 1u : 0u, 31);
	/* totals known: publish them, then fetch the next group's id and piece records while the look-back runs */
	if (lane == 0) {
		spStore(a.spA + w, spWordA(totV, totI, tail31, lastLocal, spans));
	}
	uint32_t wn = W + __shfl_sync(0xffffffffu, ticketNext, 0);
	uint2 recN = make_uint2(0u, 0u);
	if (!lateTicket && wn < numGroups && wn * 32 + lane < a.numPieces) recN = a.pieceRec[wn * 32 + lane];
	uint64_t exV = 0, exI = 0;
	uint32_t carryIn = 0, tw0 = 0;
	if (w > 0) {
		const SpLook lb = spLookBack(a, w, lane);
		if (lb.dead) {
			break;
		}
		exV = lb.exV;
		exI = lb.exI;
		carryIn = lb.carry;
		tw0 = lb.tail;
	}
	const uint64_t runV = exV + totV, runI = exI + totI;
	if (lane == 0) {
		spStore(a.spP + w, spWordP(runV, runI, spans ? carryIn + totV : lastLocal));
	}
#if VGB_EMIT_PREFETCH
	if (!lateTicket && wn < numGroups && wn * 32 + lane < a.numPieces) {
		asm volatile("prefetch.global.L2 [%0];" ::"l"(a.pts + recN.y));
		asm volatile("prefetch.global.L2 [%0];" ::"l"(a.jobRec + recN.x));
	}
#endif
	uint32_t tw = __shfl_up_sync(0xffffffffu, myTail, 1);
	if (lane == 0) {
		tw = tw0;
	}
	const ulonglong2 bb = make_ulonglong2(exV, exI);
	uint32_t F = 0;
	uint64_t mox = 0;
	uint4 prev = make_uint4(0, 0, 0, 0);
	if (valid) {
		F = inside ? vl - vlAtStart : carryIn + vl;
		mox = bb.x + vl - F; /* first vertex of the lane's mesh */
		if (P.flags & PF_HAS_PREV) {
			prev = ((OPS & OPS_OTHER) && J.op != VGB_OP_STROKE_AA) ? ribbonOutFromTail(J, tw, F)
			                                                        : out4FromTail(CT < 3 ? (uint32_t)CT : J.cap, JT < 3 ? (uint32_t)JT : J.join, tw, F);
		}
		/* the mesh table: a job's first piece knows where the mesh starts; empty jobs start where the next piece does */
		if (g == ps) {
			const ulonglong2 o = make_ulonglong2(bb.x + vl, bb.y + il);
			a.meshOffW[j] = o;
			for (uint32_t jj = j; jj-- > 0 && a.pieceStart[jj] == g;) {
				a.meshOffW[jj] = o;
			}
		}
		if (g + 1 == a.numPieces) {
			const ulonglong2 o = make_ulonglong2(runV, runI);
			for (uint32_t jj = j + 1; jj <= a.numJobs; ++jj) {
				a.meshOffW[jj] = o;
			}
		}
	}
	const uint32_t idxBase = sizeof(IDX) == 2 ? 0u : (uint32_t)(mox + a.vertexBase);
	const bool room = runV <= a.capV && runI <= a.capI; /* warp-uniform; prefixes grow, so later groups agree */
	const bool fits = totV <= CVW && totI <= CIW;        /* warp-uniform */
	if (!room) {
		if (lane == 0) atomicOr(a.spFlags, 1u);
	} else if (fits) {
		if (copyPending) {
			if (lane == 0 || lane == 2 || lane == 8) {
				bulkWaitRead();
			}
			__syncwarp();
		}
		SinkS<IDX, !(OPS & OPS_OTHER)> S;
		S.base = idxBase;
		S.pos = sPos + (uint32_t)(bb.x & 1) + vl;
		S.col = sCol + (uint32_t)(bb.x & 3) + vl;
		S.idx = sIdx + (uint32_t)(bb.y & (KI - 1)) + il;
		S.par = ((uint32_t)bb.y + il) & 1u;
		if (valid) {
			emitAny<OPS, JT, CT>(a.pts, J, P, F, prev, S);
		}
		fenceProxyAsync();
		__syncwarp();
		copyOutWarp3<IDX>(a.pos, a.col, reinterpret_cast<IDX*>(a.idx), bb.x, bb.y, totV, totI, sPos, sCol, sIdx, lane);
		copyPending = true;
	} else if (valid) {
		Sink<IDX, !(OPS & OPS_OTHER)> S;
		S.base = idxBase;
		S.par = 0;
		S.pos = a.pos + bb.x + vl;
		S.col = a.col + bb.x + vl;
		S.idx = reinterpret_cast<IDX*>(a.idx) + bb.y + il;
		emitAnyCold<OPS, JT, CT>(a.pts, J, P, F, prev, S);
	}
	/* next group */
	if (lateTicket) {
		uint32_t t = 0;
		if (lane == 0) t = atomicAdd(a.ticket, 1u);
		wn = W + __shfl_sync(0xffffffffu, t, 0);
		recN = make_uint2(0u, 0u);
		if (wn < numGroups && wn * 32 + lane < a.numPieces) recN = a.pieceRec[wn * 32 + lane];
	}
	w = wn;
	rec = recN;
	if (w >= numGroups) {
		break;
	}
	} /* groups of this warp */
	if (copyPending && (lane == 0 || lane == 2 || lane == 8)) {
		bulkWaitRead();
	}
}

/* =========================================================================================
 * pure convex-AA-fill batches (BASELINE configs[3]): sizes are closed-form (V = 2N, I = 9N - 6,
 * stroker.cpp:726-730), so there is no count pass at all, and every output of a polygon vertex has
 * a closed-form address in its mesh: the vertex pair at 2i, its fan triangle at 3i (i < N-2), its
 * fringe quad at 3(N-2) + 6i.  One thread per polygon vertex writes them straight to HBM --
 * consecutive lanes write consecutive addresses in each of the four streams.
 * ======================================================================================= */
__global__ void k_fill_job_sizes(const uint32_t* __restrict__ jobSub, uint32_t numJobs, const uint32_t* __restrict__ subNum,
                                 uint32_t* __restrict__ jobPieces, uint2* __restrict__ meshTotals)
{
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= numJobs) return;
	const uint32_t N = subNum[jobSub[j]];
	const bool ok = N >= 3; /* vg.cpp:3101 */
	jobPieces[j] = ok ? N : 0u;
	meshTotals[j] = ok ? make_uint2(2 * N, 9 * N - 6) : make_uint2(0u, 0u);
}

template<class IDX, bool SSE>
__global__ void __launch_bounds__(256) k_fill_emit(PieceArgs a)
{
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= a.numPieces) return;
	const uint32_t j = a.pieceRec[g].x;
	uint32_t ps;
	const JobCtx J = loadJob(a, j, &ps);
	const uint32_t i = g - ps, N = J.N;
	const ulonglong2 mo = a.meshOff[j];
	const float2* v = a.pts + J.first;
	const float2 p = v[i];
	const float2 d01 = vdir(v[i == 0 ? N - 1 : i - 1], p);
	const float2 d12 = vdir(p, v[i == N - 1 ? 0 : i + 1]);
	/* orientation from the first triangle, stroker.cpp:721-723 */
	const float2 p0 = v[0], p1 = v[1], p2 = v[2];
	const float ax = p1.x - p0.x, ay = p1.y - p0.y, bx = p2.x - p0.x, by = p2.y - p0.y;
	const float aa = J.fringe * 0.5f * vgb_sign(ax * by - bx * ay);
	const float2 e = extrusion(d01, d12);
	const float ex = e.x * aa, ey = e.y * aa;
	reinterpret_cast<float4*>(a.pos + mo.x)[i] = make_float4(p.x + ex, p.y + ey, p.x - ex, p.y - ey);
	reinterpret_cast<uint2*>(a.col + mo.x)[i] = make_uint2(J.color, J.c0);
	const uint32_t base = sizeof(IDX) == 2 ? (a.meshBase != nullptr ? a.meshBase[j] : 0u) : (uint32_t)(mo.x + a.vertexBase + (a.vertexBasePtr ? __ldg(a.vertexBasePtr) : 0ull));
	IDX* idx = reinterpret_cast<IDX*>(a.idx) + mo.y;
	const uint32_t f = 2 * i;
	/* scalar order (stroker.cpp:768-796): fan triangle i at 3i, quad i at 3(N-2) + 6i;
	 * SIMD order (stroker.cpp:616-701): Q0 at 0, F(k) at 6 + 9k, Q(i) at 9i for 1 <= i <= N-2, Q(N-1) at 9N - 12 */
	if (i + 2 < N) {
		IDX* t = idx + (SSE ? 6 + 9 * i : 3 * i);
		t[0] = (IDX)(base);
		t[1] = (IDX)(base + f + 2);
		t[2] = (IDX)(base + f + 4);
	}
	IDX* q = idx + (SSE ? (i == 0 ? 0u : (i + 1 == N ? 9 * N - 12 : 9 * i)) : 3 * (N - 2) + 6 * i);
	const uint32_t nA = (i + 1 < N) ? f + 3 : 1u, nB = (i + 1 < N) ? f + 2 : 0u; /* next vertex's (outer, inner), wrapping to (1, 0) */
	q[0] = (IDX)(base + f);
	q[1] = (IDX)(base + f + 1);
	q[2] = (IDX)(base + nA);
	q[3] = (IDX)(base + f);
	q[4] = (IDX)(base + nA);
	q[5] = (IDX)(base + nB);
}

/* =========================================================================================
 * host side
 * ======================================================================================= */
struct DevBuf
{
	void* p = nullptr;
	size_t cap = 0;
	int ensure(size_t bytes, std::string* err)
	{
		if (bytes <= cap) return 0;
		if (p) cudaFree(p);
		p = nullptr;
		cap = 0;
		const size_t want = bytes + bytes / 8 + 256;
		VGB_CUDA(cudaMalloc(&p, want));
		cap = want;
		return 0;
	}
	void release()
	{
		if (p) cudaFree(p);
		p = nullptr;
		cap = 0;
	}
	template<class T> T* as() const { return reinterpret_cast<T*>(p); }
};

class PipelineImpl : public Pipeline
{
public:
	int dev = 0;
	cudaStream_t st = nullptr;
	cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
	cudaEvent_t userEv[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	/* inputs */
	DevBuf dSubRec, dCmd, dData, dScale, dTol, dXform, dDraws, dJobSub, dJobDraw;
	/* intermediates */
	DevBuf dSubNum, dSubFirst, dSubClosed, dPts, dDD, dJobPieces, dPieceStart, dJobRec, dPieceJob;
	DevBuf dBlockTotals, dBlockBase, dMeshTotals, dMeshOff, dScanTmp, dBlockTail;
	DevBuf dSlotFirst, dCounters, dDenseFirst, dDensePts, dBulkErr, dMeshSplit, dTicket, dRankTotals, dRankBase;
	bool mergedRan = false;
	bool sseFill = false; /* VGB_RUN_FILL_ORDER_SSE of the current run */
	bool bulkMode = false;
	/* sizes of the uploaded batch, learnt by the first run (they are facts about the input, and the
	 * kernels are deterministic): later runs of the SAME upload reuse them instead of draining the
	 * stream three times to read them back, so the host can enqueue a whole step ahead */
	bool sizesKnown = false;
	uint32_t kNumPts = 0, kNumPieces = 0;
	uint64_t kTotals[2] = {0, 0};
	bool slotMode = false;   /* single-pass flatten into recorder-sized slots */
	bool cmdMode = false;    /* command-parallel flatten (scenes of long sub-paths) */
	DevBuf dCmdSub, dCmdData, dCmdCount, dCmdOff, dCmdOk, dSubMode;
	bool slotsUsed = false;  /* the last run's point buffer is slotted (not dense) */
	uint32_t totalSlots = 0;
	/* outputs */
	DevBuf dPos, dCol, dIdx;
	/* final layout (A9) */
	bool finalOn = false, finalRan = false;
	vgb_final_cfg fcfg{};
	uint32_t finalCmdCap = 0;
	DevBuf dSegNext[2], dSegStarts, dSegHdr, dSegCmds, dMeshBase, dUV;
	DevBuf dDrawScissor, dSplitIn, dSplitX, dSplitY, dMarkPos, dCmdMesh; /* scissor breaks in the final layout */
	DevBuf dSpDesc; /* single-pass emit: group descriptors (two 16-byte words per group) */
	int spMode = -1;                /* -1: read VGB_SINGLE_PASS on first use */
	bool spUnchecked = false;       /* a single-pass run of known sizes whose flags have not been looked at yet */
	bool haveScissor = false;
	uint32_t numScissorBreaks = 0;
	/* instance replay */
	DevBuf dRXf, dRPos, dRCol, dRIdx;
	uint32_t replayN = 0;
	bool replayRan = false;
	void* hPinned = nullptr; /* small pinned readback area */
	HostTables ht;
	bool hasXform = false, rawMode = false, uploaded = false, ran = false;
	uint32_t numSub = 0, numJobs = 0, numDraws = 0, numPaths = 0;
	uint32_t idxBytes = 2;
	uint32_t opsMask = OPS_ALL; /* which ops the uploaded draws contain */
	uint32_t uniCap = 3, uniJoin = 3; /* cap / join shared by every draw (3 = mixed): selects a specialised kernel */
	vgb_sizes sz;
	RunTimes rt;
	uint32_t launches = 0;

	~PipelineImpl() override
	{
		cudaSetDevice(dev);
		DevBuf* all[] = {&dSubRec, &dCmd, &dData, &dScale, &dTol, &dXform, &dDraws, &dJobSub, &dJobDraw, &dSubNum, &dSubFirst,
		                 &dSubClosed, &dPts, &dDD, &dJobPieces, &dPieceStart, &dJobRec, &dPieceJob, &dBlockTotals, &dBlockBase,
		                 &dMeshTotals, &dMeshOff, &dScanTmp, &dBlockTail, &dSlotFirst, &dCounters, &dDenseFirst, &dDensePts, &dBulkErr, &dMeshSplit, &dTicket, &dRankTotals, &dRankBase, &dPos, &dCol, &dIdx,
		                 &dCmdSub, &dCmdData, &dCmdCount, &dCmdOff, &dCmdOk, &dSubMode, &dSegNext[0], &dSegNext[1], &dSegStarts, &dSegHdr, &dSegCmds, &dMeshBase, &dUV, &dDrawScissor, &dSplitIn, &dSplitX, &dSplitY, &dMarkPos, &dCmdMesh, &dSpDesc, &dRXf, &dRPos, &dRCol, &dRIdx};
		for (DevBuf* b : all) b->release();
		for (auto& e : ev) if (e) cudaEventDestroy(e);
		for (auto& e : userEv) if (e) cudaEventDestroy(e);
		if (hPinned) cudaFreeHost(hPinned);
		if (st) cudaStreamDestroy(st);
	}

	int init(int device, std::string* err)
	{
		int n = 0;
		if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
			if (err) *err = "no CUDA device visible: libvgb200 has no CPU fallback";
			return VGB_ERR_NOGPU;
		}
		if (device < 0 || device >= n) {
			if (err) *err = "device ordinal out of range";
			return VGB_ERR_INVALID;
		}
		dev = device;
		VGB_CUDA(cudaSetDevice(dev));
		VGB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
		for (auto& e : ev) VGB_CUDA(cudaEventCreate(&e));
		for (auto& e : userEv) VGB_CUDA(cudaEventCreate(&e));
		VGB_CUDA(cudaHostAlloc(&hPinned, 256, cudaHostAllocDefault));
		memset(&sz, 0, sizeof(sz));
		return 0;
	}

	template<class T>
	int up(DevBuf& b, const T* src, size_t n, std::string* err)
	{
		if (b.ensure(n * sizeof(T) + 16, err)) return VGB_ERR_CUDA;
		if (n) VGB_CUDA(cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
		return 0;
	}

	int upload(const HostTables& t, std::string* err) override
	{
		VGB_CUDA(cudaSetDevice(dev));
		ht = t;
		rawMode = t.rawPts != nullptr;
		slotMode = false;
		cmdMode = false;
		numSub = t.numSub;
		numJobs = t.numJobs;
		numDraws = t.numDraws;
		numPaths = t.numPaths;
		hasXform = t.pathXform != nullptr;
		int rc = 0;
		bulkMode = t.bulkP8 != nullptr;
		if (bulkMode) {
			const uint32_t n = t.bulkN;
			numPaths = n;
			if (dScale.ensure((size_t)n * 4 + 16, err) || dTol.ensure((size_t)n * 4 + 16, err) || dCmd.ensure((size_t)n * 2 + 16, err) ||
			    dSubRec.ensure((size_t)n * sizeof(SubRec) + 16, err) || dJobSub.ensure((size_t)n * 4 + 16, err) ||
			    dJobDraw.ensure((size_t)n * 4 + 16, err) || dSlotFirst.ensure((size_t)(n + 1) * 4, err) || dBulkErr.ensure(16, err))
				return VGB_ERR_CUDA;
			rc |= up(dData, t.bulkP8, 8 * (size_t)n, err); /* DMA straight from the caller's buffer */
			VGB_CUDA(cudaMemsetAsync(dBulkErr.p, 0, 4, st));
			k_bulk_cubic_tables<<<(n + 1 + 255) / 256, 256, 0, st>>>(dData.as<float>(), n, t.bulkScale, t.bulkTol, dCmd.as<uint8_t>(), dSubRec.as<SubRec>(),
			                                                         dScale.as<float>(), dTol.as<float>(), dJobSub.as<uint32_t>(), dJobDraw.as<uint32_t>(),
			                                                         dSlotFirst.as<uint32_t>(), dBulkErr.as<uint32_t>());
			VGB_CUDA(cudaGetLastError());
			hasXform = false;
			slotMode = n > 0 && (uint64_t)n * BULK_SLOT < 0xE0000000ull && getenv("VGB_TWO_PASS_FLATTEN") == nullptr;
			totalSlots = n * BULK_SLOT;
		} else {
			rc |= up(dScale, t.pathScale, t.numPaths, err);
			rc |= up(dTol, t.pathTol, t.numPaths, err);
		}
		if (bulkMode) {
			/* nothing else to ship */
		} else if (rawMode) {
			rc |= up(dPts, t.rawPts, 2 * (size_t)t.numRawPts, err);
			rc |= up(dSubNum, t.rawNum, t.numSub, err);
			rc |= up(dSubFirst, t.rawFirst, t.numSub, err);
			rc |= up(dSubClosed, t.rawClosed, t.numSub, err);
		} else {
			rc |= up(dSubRec, t.subRec, t.numSub, err);
			rc |= up(dCmd, t.cmd, t.numCmd, err);
			rc |= up(dData, t.data, t.numData, err);
			if (hasXform) rc |= up(dXform, t.pathXform, 6 * (size_t)t.numPaths, err);
			slotMode = t.subSlotFirst != nullptr && t.numSub > 0 && getenv("VGB_TWO_PASS_FLATTEN") == nullptr;
			/* long sub-paths (8+ commands each on average; with VGB_CMD_FLATTEN=1: 4+): flatten per command.
			 * Measured on the tiger-like scene (13 commands per sub-path, x300 instances): flatten 1.13 -> 0.97 ms;
			 * short sub-paths keep the single-pass slot flatten.  VGB_CMD_FLATTEN=0 switches it off. */
			{
				const char* e = getenv("VGB_CMD_FLATTEN");
				const bool dflt = VGB_CMD_FLATTEN_DEFAULT != 0;
				const uint64_t minCmds = (e && e[0] == '1') ? 4ull : 8ull;
				cmdMode = t.numSub > 0 && (uint64_t)t.numCmd >= minCmds * t.numSub && (e ? e[0] == '1' : dflt);
				if (cmdMode) slotMode = false;
			}
			if (slotMode) {
				totalSlots = t.subSlotFirst[t.numSub];
				rc |= up(dSlotFirst, t.subSlotFirst, (size_t)t.numSub + 1, err);
			}
		}
		opsMask = 0;
		uniCap = uniJoin = 3;
		for (uint32_t d = 0; d < t.numDraws; ++d) {
			const uint32_t op = t.draws[d].ocj & 0xFF;
			const uint32_t cap = (t.draws[d].ocj >> 8) & 0xFF, join = (t.draws[d].ocj >> 16) & 0xFF;
			uniCap = d == 0 ? cap : (uniCap == cap ? cap : 3u);
			uniJoin = d == 0 ? join : (uniJoin == join ? join : 3u);
			opsMask |= (op == VGB_OP_STROKE_AA) ? OPS_STROKE_AA : (op == VGB_OP_CONVEX_FILL_AA ? OPS_FILL_AA : OPS_OTHER);
		}
		if (opsMask == 0 || (opsMask & (opsMask - 1))) opsMask = OPS_ALL; /* mixed batches use the general kernel */
		rc |= up(dDraws, t.draws, t.numDraws, err);
		haveScissor = t.drawScissor != nullptr;
		numScissorBreaks = haveScissor ? t.numScissorBreaks : 0u;
		if (haveScissor) rc |= up(dDrawScissor, t.drawScissor, t.numDraws, err);
		if (!bulkMode) {
			rc |= up(dJobSub, t.jobSub, t.numJobs, err);
			rc |= up(dJobDraw, t.jobDraw, t.numJobs, err);
		}
		if (rc) return VGB_ERR_CUDA;
		uploaded = true;
		ran = false;
		sizesKnown = false;
		return 0;
	}

	template<class O>
	int scan(const typename O::In* in, uint32_t n, typename O::Acc* out, std::string* err)
	{
		typedef typename O::Acc Acc;
		const uint32_t blocks = (n + SCAN_TILE - 1) / SCAN_TILE;
		if (dScanTmp.ensure((size_t)(blocks + 1) * sizeof(Acc), err)) return VGB_ERR_CUDA;
		Acc* partial = dScanTmp.as<Acc>();
		if (blocks == 0) {
			VGB_CUDA(cudaMemsetAsync(out, 0, sizeof(Acc), st));
			return 0;
		}
		k_scan_reduce<O><<<blocks, SCAN_THREADS, 0, st>>>(in, n, partial);
		k_scan_partials<O><<<1, 1024, 0, st>>>(partial, blocks);
		k_scan_final<O><<<blocks, SCAN_THREADS, 0, st>>>(in, n, partial, out);
		launches += 3;
		VGB_CUDA(cudaGetLastError());
		return 0;
	}

	/* read back a few bytes (synchronises the stream) */
	int readback(void* dst, const void* src, size_t bytes, std::string* err)
	{
		VGB_CUDA(cudaMemcpyAsync(hPinned, src, bytes, cudaMemcpyDeviceToHost, st));
		VGB_CUDA(cudaStreamSynchronize(st));
		memcpy(dst, hPinned, bytes);
		return 0;
	}

	/* the emit kernel is persistent: as many blocks as the device holds at once (occupancy x SMs x VGB_EMIT_WAVES,
	 * default 1), each warp striding over the groups */
	int numSMs = 0;
	template<class IDX, uint32_t OPS, int JT, int CT, bool BIGI>
	void launchEmit(const PieceArgs& pa, uint32_t gridBlocks)
	{
		static int perSM = 0;
		static int waves = 0;
		if (perSM == 0) {
			/* a persistent grid must be fully resident from the start: ask for the shared-memory carve-out its
			 * 8-9 blocks per SM need (left to the driver's heuristic it settled, after a few launches, on a
			 * smaller one, and half of the grid only started once the other half had taken every ticket) */
			cudaFuncSetAttribute(k_group_emit<IDX, OPS, JT, CT, BIGI>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
			if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_group_emit<IDX, OPS, JT, CT, BIGI>, (int)PPB, 0) != cudaSuccess || perSM < 1) perSM = 8;
			const char* e = getenv("VGB_EMIT_WAVES");
			waves = e ? atoi(e) : 1;
			if (waves < 1) waves = 1;
			const char* f = getenv("VGB_EMIT_PER_SM"); /* tuning / diagnosis: resident blocks per SM the grid is sized for */
			if (f && atoi(f) > 0) perSM = atoi(f);
		}
		if (numSMs == 0) {
			if (cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || numSMs < 1) numSMs = 148;
		}
		const uint32_t cap = (uint32_t)perSM * (uint32_t)numSMs * (uint32_t)waves;
		if (getenv("VGB_DEBUG")) fprintf(stderr, "[vgb] emit grid: perSM %d numSMs %d waves %d blocks %u of %u\n", perSM, numSMs, waves, gridBlocks < cap ? gridBlocks : cap, gridBlocks);
		/* large batches: a ticket is a chunk of 4 consecutive groups (a quarter of the same-address atomics);
		 * small ones keep one group per ticket so that every warp has work at once */
		PieceArgs q = pa;
		q.chunkShift = gridBlocks >= 8 * cap ? 2u : 0u;
		const uint32_t need = (gridBlocks + (1u << q.chunkShift) - 1) >> q.chunkShift;
		k_group_emit<IDX, OPS, JT, CT, BIGI><<<need < cap ? need : cap, PPB, 0, st>>>(q);
	}

	template<class IDX, uint32_t OPS, int JT, int CT, bool BIGI>
	void launchEmitSp(const PieceArgs& pa, uint32_t gridBlocks)
	{
		static int perSM = 0;
		if (perSM == 0) {
			cudaFuncSetAttribute(k_group_emit_sp<IDX, OPS, JT, CT, BIGI>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
			if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_group_emit_sp<IDX, OPS, JT, CT, BIGI>, (int)PPB, 0) != cudaSuccess || perSM < 1) perSM = 8;
			const char* f = getenv("VGB_EMIT_PER_SM");
			if (f && atoi(f) > 0) perSM = atoi(f);
		}
		if (numSMs == 0) {
			if (cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || numSMs < 1) numSMs = 148;
		}
		const uint32_t cap = (uint32_t)perSM * (uint32_t)numSMs;
		/* one group per ticket: a decoupled look-back only stays decoupled if every group publishes its totals a
		 * (roughly) fixed time after its ticket was drawn.  A warp walking the 4 consecutive groups of one ticket
		 * publishes the 4th three groups later than the 1st, every later ticket waits for it, and the waits chain
		 * up over the whole launch (measured: 2.9 s instead of 1 ms on the headline workload) */
		PieceArgs q = pa;
		static int spLate = -1;
		if (spLate < 0) {
			const char* e = getenv("VGB_SP_LATE_TICKET"); /* 1: draw a group's ticket when the group starts */
			spLate = e ? atoi(e) : 0;
		}
		q.chunkShift = spLate ? 1u : 0u; /* the single-pass kernel reads it as the ticket policy */
		const uint32_t need = gridBlocks;
		if (getenv("VGB_DEBUG")) fprintf(stderr, "[vgb] single-pass emit grid: perSM %d blocks %u of %u\n", perSM, need < cap ? need : cap, gridBlocks);
		k_group_emit_sp<IDX, OPS, JT, CT, BIGI><<<need < cap ? need : cap, PPB, 0, st>>>(q);
	}

	int run(uint32_t flags, uint64_t vertexBase, std::string* err, const MergeHook* merge = nullptr) override
	{
		if (!uploaded) {
			if (err) *err = "run before upload";
			return VGB_ERR_STATE;
		}
		VGB_CUDA(cudaSetDevice(dev));
		launches = 0;
		idxBytes = (flags & VGB_RUN_INDEX_U32) ? 4 : 2;
		{
			const bool sse = (flags & VGB_RUN_FILL_ORDER_SSE) != 0;
			if (sse != sseFill) sizesKnown = false; /* the piece decomposition of fills differs between the two orders */
			sseFill = sse;
		}
		memset(&sz, 0, sizeof(sz));
		sz.num_subpaths = numSub;
		sz.num_meshes = numJobs;
		VGB_CUDA(cudaEventRecord(ev[0], st));

		/* ---- flatten ---- */
		uint32_t numPts = 0;
		if (rawMode) {
			numPts = ht.numRawPts;
		} else if (numSub) {
			if (dSubNum.ensure((size_t)numSub * 4, err) || dSubFirst.ensure((size_t)(numSub + 1) * 4, err) || dSubClosed.ensure(numSub, err))
				return VGB_ERR_CUDA;
			FlattenArgs fa;
			fa.subRec = dSubRec.as<SubRec>();
			fa.numSub = numSub;
			fa.cmd = dCmd.as<uint8_t>();
			fa.data = dData.as<float>();
			fa.pathScale = dScale.as<float>();
			fa.pathTol = dTol.as<float>();
			fa.pathXform = hasXform ? dXform.as<float>() : nullptr;
			fa.subNum = dSubNum.as<uint32_t>();
			fa.subFirst = dSubFirst.as<uint32_t>();
			fa.subClosed = dSubClosed.as<uint8_t>();
			fa.pts = nullptr;
			fa.subMode = nullptr;
			const uint32_t fb = (numSub + 127) / 128;
			slotsUsed = false;
			if (slotMode) {
				const uint32_t arenaCap = totalSlots / 2 + (1u << 20);
				if (dPts.ensure(((size_t)totalSlots + arenaCap) * 8 + 16, err) || dCounters.ensure(16, err)) return VGB_ERR_CUDA;
				VGB_CUDA(cudaMemsetAsync(dCounters.p, 0, 16, st));
				SlotArgs sa;
				fa.pts = dPts.as<float2>();
				sa.f = fa;
				sa.slotFirst = dSlotFirst.as<uint32_t>();
				sa.subFirstOut = dSubFirst.as<uint32_t>();
				sa.arenaBase = totalSlots;
				sa.arenaCap = arenaCap;
				sa.counters = dCounters.as<uint32_t>();
				k_flatten_slots<<<fb, 128, 0, st>>>(sa);
				++launches;
				VGB_CUDA(cudaGetLastError());
				slotsUsed = true;
				numPts = 0xFFFFFFFFu; /* read back later together with the piece count */
			} else {
				CmdArgs ca;
				const uint32_t numCmd = ht.numCmd;
				const uint32_t cb = (numCmd + 127) / 128;
				if (cmdMode) {
					if (dCmdSub.ensure((size_t)numCmd * 4, err) || dCmdData.ensure((size_t)numCmd * 4, err) || dCmdCount.ensure((size_t)numCmd * 4, err) ||
					    dCmdOff.ensure((size_t)numCmd * 4, err) || dCmdOk.ensure(numCmd, err) || dSubMode.ensure(numSub, err))
						return VGB_ERR_CUDA;
					fa.subMode = dSubMode.as<uint8_t>();
					ca.f = fa;
					ca.numCmd = numCmd;
					ca.cmdSub = dCmdSub.as<uint32_t>();
					ca.cmdData = dCmdData.as<uint32_t>();
					ca.cmdCount = dCmdCount.as<uint32_t>();
					ca.cmdOff = dCmdOff.as<uint32_t>();
					ca.cmdOk = dCmdOk.as<uint8_t>();
					ca.subModeW = dSubMode.as<uint8_t>();
					k_cmd_index<<<fb, 128, 0, st>>>(ca);
					k_cmd_flatten<false><<<cb, 128, 0, st>>>(ca);
					k_cmd_sub_sizes<<<fb, 128, 0, st>>>(ca);
					launches += 3;
				}
				k_flatten<false><<<fb, 128, 0, st>>>(fa);
				++launches;
				if (scan<ScanOps<uint32_t>>(dSubNum.as<uint32_t>(), numSub, dSubFirst.as<uint32_t>(), err)) return VGB_ERR_CUDA;
				if (sizesKnown) {
					numPts = kNumPts;
				} else if (readback(&numPts, dSubFirst.as<uint32_t>() + numSub, 4, err)) {
					return VGB_ERR_CUDA;
				}
				if (dPts.ensure((size_t)numPts * 8 + 16, err)) return VGB_ERR_CUDA;
				fa.pts = dPts.as<float2>();
				if (cmdMode) {
					ca.f = fa;
					k_cmd_flatten<true><<<cb, 128, 0, st>>>(ca);
					++launches;
				}
				k_flatten<true><<<fb, 128, 0, st>>>(fa);
				++launches;
				VGB_CUDA(cudaGetLastError());
			}
		}
		sz.num_points = numPts;
		VGB_CUDA(cudaEventRecord(ev[1], st));

		/* ---- draws, jobs, pieces ---- */
		const bool fillOnly = opsMask == OPS_FILL_AA && getenv("VGB_GENERIC_FILL") == nullptr;
		uint32_t numPieces = 0;
		if (numJobs) {
			if (dDD.ensure((size_t)numDraws * sizeof(DrawDerived), err) || dJobPieces.ensure((size_t)numJobs * 4, err) ||
			    dPieceStart.ensure((size_t)(numJobs + 1) * 4, err))
				return VGB_ERR_CUDA;
			k_draw_prep<<<(numDraws + 127) / 128, 128, 0, st>>>(dDraws.as<DrawRec>(), numDraws, dScale.as<float>(), dTol.as<float>(), ht.fringe,
			                                                     dDD.as<DrawDerived>());
			if (fillOnly) {
				if (dMeshTotals.ensure((size_t)numJobs * 8, err)) return VGB_ERR_CUDA;
				k_fill_job_sizes<<<(numJobs + 255) / 256, 256, 0, st>>>(dJobSub.as<uint32_t>(), numJobs, dSubNum.as<uint32_t>(),
				                                                         dJobPieces.as<uint32_t>(), dMeshTotals.as<uint2>());
			} else {
				k_job_pieces<<<(numJobs + 255) / 256, 256, 0, st>>>(dJobSub.as<uint32_t>(), dJobDraw.as<uint32_t>(), numJobs, dSubNum.as<uint32_t>(),
				                                                     dDraws.as<DrawRec>(), dJobPieces.as<uint32_t>(), sseFill ? 1u : 0u);
			}
			launches += 2;
			if (scan<ScanOps<uint32_t>>(dJobPieces.as<uint32_t>(), numJobs, dPieceStart.as<uint32_t>(), err)) return VGB_ERR_CUDA;
			if (sizesKnown) {
				numPieces = kNumPieces;
			}
		}
		if (!sizesKnown) {
			/* first run of an upload: the piece count, the bulk input's error flag and the slot-flatten counters come
			 * back in ONE read (three small copies, one stream drain) */
			uint32_t* h = reinterpret_cast<uint32_t*>(hPinned);
			memset(h, 0, 32);
			if (numJobs) VGB_CUDA(cudaMemcpyAsync(h, dPieceStart.as<uint32_t>() + numJobs, 4, cudaMemcpyDeviceToHost, st));
			if (bulkMode) VGB_CUDA(cudaMemcpyAsync(h + 1, dBulkErr.p, 4, cudaMemcpyDeviceToHost, st));
			if (slotsUsed) VGB_CUDA(cudaMemcpyAsync(h + 4, dCounters.p, 16, cudaMemcpyDeviceToHost, st));
			VGB_CUDA(cudaStreamSynchronize(st));
			numPieces = h[0];
			if (bulkMode && h[1]) {
				if (err) *err = "non-finite coordinate in the bulk input";
				return VGB_ERR_NONFINITE;
			}
			if (slotsUsed) {
				if (h[5]) {
					/* the overflow arena was too small: redo this run with the two-pass flatten */
					slotMode = false;
					const int rc = run(flags, vertexBase, err, merge);
					return rc;
				}
				numPts = h[6];
				sz.num_points = numPts;
			}
		} else if (slotsUsed) {
			numPts = kNumPts;
			sz.num_points = numPts;
		}
		bool useSp = false;
		const uint32_t numBlocks = (numPieces + GRP - 1) / GRP;          /* offset groups (warps or blocks) */
		const uint32_t gridBlocks = (numPieces + PPB - 1) / PPB;         /* 128-thread launch blocks      */
		uint64_t totals[2] = {0, 0};
		if (dMeshOff.ensure((size_t)(numJobs + 1) * 16, err)) return VGB_ERR_CUDA;
		PieceArgs pa;
		memset(&pa, 0, sizeof(pa));
		if (numPieces) {
			if (dJobRec.ensure((size_t)numJobs * sizeof(JobRec), err) || dPieceJob.ensure((size_t)numPieces * 8, err) ||
			    dBlockTotals.ensure((size_t)numBlocks * 8, err) || dBlockBase.ensure((size_t)(numBlocks + 1) * 16, err) ||
			    dMeshTotals.ensure((size_t)numJobs * 8, err) || dBlockTail.ensure((size_t)numBlocks * 4, err))
				return VGB_ERR_CUDA;
			k_job_prep<<<(numJobs + 255) / 256, 256, 0, st>>>(dJobSub.as<uint32_t>(), dJobDraw.as<uint32_t>(), numJobs, dSubFirst.as<uint32_t>(),
			                                                   dSubNum.as<uint32_t>(), dSubClosed.as<uint8_t>(), dDD.as<DrawDerived>(),
			                                                   dPieceStart.as<uint32_t>(), dJobRec.as<JobRec>(), dPieceJob.as<uint2>(), sseFill ? 1u : 0u);
			++launches;
			pa.numPieces = numPieces;
			pa.numJobs = numJobs;
			pa.pieceRec = dPieceJob.as<uint2>();
			pa.jobRec = dJobRec.as<JobRec>();
			pa.pts = dPts.as<float2>();
			pa.blockTail = dBlockTail.as<uint32_t>();
			pa.blockTotals = dBlockTotals.as<uint2>();
			pa.meshTotals = dMeshTotals.as<uint2>();
			pa.blockBase = dBlockBase.as<ulonglong2>();
			pa.meshOff = dMeshOff.as<ulonglong2>();
			pa.vertexBase = vertexBase;
			if (dTicket.ensure(16, err)) return VGB_ERR_CUDA;
			pa.ticket = dTicket.as<uint32_t>();
			/* single-pass emit (k_group_emit_sp): no count pass; needs output buffers before the sizes are known, so it
			 * runs when a previous run left buffers behind (or the sizes of this upload are known), and never with the
			 * final layout or a merged run, whose tables need the totals before the first vertex is written */
			if (spMode < 0) {
				const char* e = getenv("VGB_SINGLE_PASS");
				spMode = e ? atoi(e) : VGB_SINGLE_PASS_DEFAULT;
			}
			useSp = spMode > 0 && !fillOnly && !finalOn && merge == nullptr &&
			        (sizesKnown || (dPos.cap > 64 && dCol.cap > 64 && dIdx.cap > 64));
			if (useSp) {
				/* nothing to do here: classification happens in the emit kernel */
			} else if (fillOnly) {
				/* closed-form sizes: the mesh table is a plain scan of (2N, 9N-6); no count kernel */
				if (scan<ScanOps<uint2>>(dMeshTotals.as<uint2>(), numJobs, dMeshOff.as<ulonglong2>(), err)) return VGB_ERR_CUDA;
			} else {
#define VGB_COUNT_K k_group_count
			const uint32_t countBlocks = (numBlocks + 4 * GC - 1) / (4 * GC); /* 4 warps per block, GC groups per warp */
			if (opsMask == OPS_STROKE_AA && uniJoin == VGB_JOIN_ROUND && uniCap == VGB_CAP_ROUND) VGB_COUNT_K<OPS_STROKE_AA, 1, 1><<<countBlocks, PPB, 0, st>>>(pa);
			else if (opsMask == OPS_STROKE_AA && uniJoin == VGB_JOIN_MITER && uniCap == VGB_CAP_BUTT) VGB_COUNT_K<OPS_STROKE_AA, 0, 0><<<countBlocks, PPB, 0, st>>>(pa);
			else if (opsMask == OPS_STROKE_AA) VGB_COUNT_K<OPS_STROKE_AA, 3, 3><<<countBlocks, PPB, 0, st>>>(pa);
			else if (opsMask == OPS_FILL_AA) VGB_COUNT_K<OPS_FILL_AA, 3, 3><<<countBlocks, PPB, 0, st>>>(pa);
			else VGB_COUNT_K<OPS_ALL, 3, 3><<<countBlocks, PPB, 0, st>>>(pa);
#undef VGB_COUNT_K
			++launches;
			if (scan<ScanOps<uint2>>(dBlockTotals.as<uint2>(), numBlocks, dBlockBase.as<ulonglong2>(), err)) return VGB_ERR_CUDA;
			k_mesh_off<<<(numJobs + 1 + 255) / 256, 256, 0, st>>>(dPieceStart.as<uint32_t>(), dPieceJob.as<uint2>(), numJobs, numPieces, numBlocks,
			                                                       dBlockBase.as<ulonglong2>(), dMeshTotals.as<uint2>(), dMeshOff.as<ulonglong2>(), dTicket.as<uint32_t>());
			++launches;
			} /* !fillOnly */
			if (sizesKnown) {
				totals[0] = kTotals[0];
				totals[1] = kTotals[1];
			} else if (!useSp && readback(totals, dMeshOff.as<ulonglong2>() + numJobs, 16, err)) {
				return VGB_ERR_CUDA;
			}
		} else {
			VGB_CUDA(cudaMemsetAsync(dMeshOff.p, 0, (size_t)(numJobs + 1) * 16, st));
		}
		sz.num_vertices = totals[0];
		sz.num_indices = totals[1];
		finalRan = false;
		if (finalOn && numPieces) {
			if (idxBytes != 2) {
				if (err) *err = "the final layout has uint16 indices: run with VGB_RUN_INDEX_U16";
				return VGB_ERR_INVALID;
			}
			/* two consecutive segments always hold more than max vertices together, so there are at most
			 * 2V/max + 2 of them: K doubling rounds with 2^K above that bound */
			const uint64_t segBound = 2 * totals[0] / fcfg.max_vb_vertices + 3;
			uint32_t K = 0;
			while (((uint64_t)1 << K) < segBound) ++K;
			const uint32_t numStarts = 1u << K;
			const bool split = numScissorBreaks > 0;
			finalCmdCap = numStarts + numScissorBreaks;
			if (dSegNext[0].ensure((size_t)(numJobs + 1) * 4, err) || dSegNext[1].ensure((size_t)(numJobs + 1) * 4, err) ||
			    dSegStarts.ensure((size_t)numStarts * 4, err) || dSegHdr.ensure(sizeof(FinalHdr), err) ||
			    dSegCmds.ensure((size_t)finalCmdCap * sizeof(vgb_draw_cmd), err) || dMeshBase.ensure((size_t)numJobs * 4, err))
				return VGB_ERR_CUDA;
			VGB_CUDA(cudaMemsetAsync(dSegHdr.p, 0, sizeof(FinalHdr), st));
			const uint32_t jb = (numJobs + 1 + 255) / 256;
			k_seg_next<<<jb, 256, 0, st>>>(dMeshOff.as<ulonglong2>(), numJobs, fcfg.max_vb_vertices, fcfg.vb_id, fcfg.vb_count, fcfg.cmd_vertices,
			                               dSegNext[0].as<uint32_t>(), dSegStarts.as<uint32_t>(), dSegHdr.as<FinalHdr>());
			++launches;
			for (uint32_t k = 0; k < K; ++k) {
				const uint32_t half = 1u << k, last = k + 1 == K;
				const uint32_t nb = last ? (half + 255) / 256 : jb > (half + 255) / 256 ? jb : (half + 255) / 256;
				k_seg_double<<<nb, 256, 0, st>>>(dSegNext[k & 1].as<uint32_t>(), dSegNext[(k + 1) & 1].as<uint32_t>(), numJobs, dSegStarts.as<uint32_t>(), half,
				                                 last ? 0u : 1u);
				++launches;
			}
			if (!split) {
				k_seg_cmds<<<(numStarts + 255) / 256, 256, 0, st>>>(dSegStarts.as<uint32_t>(), numStarts, dMeshOff.as<ulonglong2>(), numJobs, fcfg.ib_count,
				                                                     dJobDraw.as<uint32_t>(), haveScissor ? dDrawScissor.as<uint32_t>() : nullptr,
				                                                     dSegHdr.as<FinalHdr>(), dSegCmds.as<vgb_draw_cmd>());
				k_seg_mesh_base<<<(numJobs + 255) / 256, 256, 0, st>>>(dSegStarts.as<uint32_t>(), numStarts, dMeshOff.as<ulonglong2>(), numJobs,
				                                                        dSegHdr.as<FinalHdr>(), dMeshBase.as<uint32_t>());
				launches += 2;
			} else {
				/* scissor changes open draw commands inside a vertex buffer */
				if (dSplitIn.ensure((size_t)numJobs * 8, err) || dSplitX.ensure((size_t)(numJobs + 1) * 16, err) ||
				    dSplitY.ensure((size_t)(numJobs + 1) * 16, err) || dMarkPos.ensure((size_t)(numScissorBreaks + 1) * 4, err) ||
				    dCmdMesh.ensure((size_t)(finalCmdCap + 1) * 4, err))
					return VGB_ERR_CUDA;
				const uint32_t mb = (numJobs + 255) / 256;
				k_split_marks<<<mb, 256, 0, st>>>(dMeshOff.as<ulonglong2>(), numJobs, dJobDraw.as<uint32_t>(), dDrawScissor.as<uint32_t>(), dSplitIn.as<uint2>());
				if (scan<ScanOps<uint2>>(dSplitIn.as<uint2>(), numJobs, dSplitX.as<ulonglong2>(), err)) return VGB_ERR_CUDA;
				k_split_pos<<<mb, 256, 0, st>>>(dSplitIn.as<uint2>(), dSplitX.as<ulonglong2>(), numJobs, dMarkPos.as<uint32_t>());
				k_split_flags<<<mb, 256, 0, st>>>(dSplitIn.as<uint2>(), dSplitX.as<ulonglong2>(), numJobs, dMarkPos.as<uint32_t>(), dSegStarts.as<uint32_t>(),
				                                   numStarts);
				if (scan<ScanOps<uint2>>(dSplitIn.as<uint2>(), numJobs, dSplitY.as<ulonglong2>(), err)) return VGB_ERR_CUDA;
				k_split_cmds<<<mb, 256, 0, st>>>(dSplitIn.as<uint2>(), dSplitY.as<ulonglong2>(), numJobs, dSegStarts.as<uint32_t>(), dMeshOff.as<ulonglong2>(),
				                                  fcfg.ib_count, dJobDraw.as<uint32_t>(), dDrawScissor.as<uint32_t>(), dSegHdr.as<FinalHdr>(),
				                                  dSegCmds.as<vgb_draw_cmd>(), dCmdMesh.as<uint32_t>());
				const uint32_t fb = ((numJobs > finalCmdCap ? numJobs : finalCmdCap) + 255) / 256;
				k_split_fin<<<fb, 256, 0, st>>>(dSplitIn.as<uint2>(), dSplitY.as<ulonglong2>(), numJobs, dMeshOff.as<ulonglong2>(), dSegHdr.as<FinalHdr>(),
				                                 dCmdMesh.as<uint32_t>(), dSegCmds.as<vgb_draw_cmd>(), dMeshBase.as<uint32_t>());
				launches += 5;
			}
			VGB_CUDA(cudaGetLastError());
			pa.meshBase = dMeshBase.as<uint32_t>();
			finalRan = true;
		}
		mergedRan = false;
		if (merge != nullptr && merge->allGatherU64 != nullptr) {
			/* the one exchange step: all-gather of every rank's (V, I) on this stream, exclusive base on the device */
			if (dRankTotals.ensure((size_t)merge->nranks * 16 + 16, err) || dRankBase.ensure(32, err)) return VGB_ERR_CUDA;
			const int rc = merge->allGatherU64(merge->user, dMeshOff.as<ulonglong2>() + numJobs, dRankTotals.p, 2, (void*)st, err);
			if (rc) return rc;
			k_rank_base<<<1, 32, 0, st>>>(dRankTotals.as<uint64_t>(), (uint32_t)merge->rank, (uint32_t)merge->nranks, dRankBase.as<uint64_t>());
			++launches;
			VGB_CUDA(cudaGetLastError());
			pa.vertexBasePtr = dRankBase.as<uint64_t>();
			mergedRan = true;
		}
		VGB_CUDA(cudaEventRecord(ev[2], st));

		/* ---- emit ---- */
		if (numPieces && useSp) {
			if (dSpDesc.ensure((size_t)numBlocks * 32, err)) return VGB_ERR_CUDA;
			pa.spA = dSpDesc.as<uint4>();
			pa.spP = dSpDesc.as<uint4>() + numBlocks;
			pa.spFlags = dTicket.as<uint32_t>() + 1;
			pa.pieceStart = dPieceStart.as<uint32_t>();
			pa.meshOffW = dMeshOff.as<ulonglong2>();
			const bool bigI = idxBytes == 4 && sizesKnown && numBlocks > 0 && totals[1] / numBlocks > (uint64_t)(CIW_COUNT / 2) * 3 / 4;
			for (int attempt = 0;; ++attempt) {
				if (sizesKnown || attempt > 0) {
					if (dPos.ensure(totals[0] * 8 + 32, err) || dCol.ensure(totals[0] * 4 + 32, err) || dIdx.ensure(totals[1] * idxBytes + 32, err))
						return VGB_ERR_CUDA;
				}
				pa.pos = dPos.as<float2>();
				pa.col = dCol.as<uint32_t>();
				pa.idx = dIdx.p;
				const uint64_t cv0 = (dPos.cap - 32) / 8, cv1 = (dCol.cap - 32) / 4;
				pa.capV = cv0 < cv1 ? cv0 : cv1;
				pa.capI = (dIdx.cap - 32) / idxBytes;
				VGB_CUDA(cudaMemsetAsync(dSpDesc.p, 0, (size_t)numBlocks * 32, st));
				VGB_CUDA(cudaMemsetAsync(dTicket.p, 0, 16, st));
#define VGB_EMIT_K(T, O, J, C)                                                   \
	do {                                                                         \
		if (bigI) launchEmitSp<T, O, J, C, true>(pa, gridBlocks);                 \
		else launchEmitSp<T, O, J, C, false>(pa, gridBlocks);                     \
	} while (0)
#define VGB_LAUNCH_EMIT(T)                                                                                  \
	do {                                                                                                    \
		if (opsMask == OPS_STROKE_AA && uniJoin == VGB_JOIN_ROUND && uniCap == VGB_CAP_ROUND) VGB_EMIT_K(T, OPS_STROKE_AA, 1, 1); \
		else if (opsMask == OPS_STROKE_AA && uniJoin == VGB_JOIN_MITER && uniCap == VGB_CAP_BUTT) VGB_EMIT_K(T, OPS_STROKE_AA, 0, 0); \
		else if (opsMask == OPS_STROKE_AA) VGB_EMIT_K(T, OPS_STROKE_AA, 3, 3); \
		else if (opsMask == OPS_FILL_AA) VGB_EMIT_K(T, OPS_FILL_AA, 3, 3); \
		else VGB_EMIT_K(T, OPS_ALL, 3, 3);                                 \
	} while (0)
				if (idxBytes == 2) {
					if (bigI) return VGB_ERR_STATE; /* uint16 + BIGI is the final layout: never single-pass */
					VGB_LAUNCH_EMIT(uint16_t);
				} else {
					VGB_LAUNCH_EMIT(uint32_t);
				}
#undef VGB_LAUNCH_EMIT
#undef VGB_EMIT_K
				++launches;
				VGB_CUDA(cudaGetLastError());
				if (sizesKnown) {
					/* the buffers hold exactly this upload's output: nothing can overflow, and the flags are looked at
					 * where the stream is drained anyway (sync / download) instead of draining it here */
					spUnchecked = true;
					break;
				}
				/* a new upload: the totals the scan produced, and whether everything fitted */
				uint64_t* h = reinterpret_cast<uint64_t*>(hPinned);
				VGB_CUDA(cudaMemcpyAsync(h, dMeshOff.as<ulonglong2>() + numJobs, 16, cudaMemcpyDeviceToHost, st));
				VGB_CUDA(cudaMemcpyAsync(h + 2, dTicket.as<uint32_t>() + 1, 4, cudaMemcpyDeviceToHost, st));
				VGB_CUDA(cudaStreamSynchronize(st));
				const uint32_t spf = *reinterpret_cast<uint32_t*>(h + 2);
				if ((spf & 2u) || attempt > 1) {
					/* the look-back gave up (never expected): this batch object goes back to the two-pass pipeline */
					spMode = 0;
					return run(flags, vertexBase, err, merge);
				}
				totals[0] = h[0];
				totals[1] = h[1];
				if (!(spf & 1u)) break;
				/* some groups did not fit: grow to the totals (now known) and emit again */
			}
			sz.num_vertices = totals[0];
			sz.num_indices = totals[1];
		} else if (numPieces) {
			if (dPos.ensure(totals[0] * 8 + 32, err) || dCol.ensure(totals[0] * 4 + 32, err) || dIdx.ensure(totals[1] * idxBytes + 32, err))
				return VGB_ERR_CUDA;
			pa.pos = dPos.as<float2>();
			pa.col = dCol.as<uint32_t>();
			pa.idx = dIdx.p;
			/* uint32 indices: pick the index-image size from the batch's average indices per group */
			const bool bigI = idxBytes == 4 ? (numBlocks > 0 && totals[1] / numBlocks > (uint64_t)(CIW_COUNT / 2) * 3 / 4) : finalRan;
#define VGB_EMIT_K(T, O, J, C)                                                   \
	do {                                                                         \
		if (bigI) launchEmit<T, O, J, C, true>(pa, gridBlocks);                   \
		else launchEmit<T, O, J, C, false>(pa, gridBlocks);                       \
	} while (0)
#define VGB_LAUNCH_EMIT(T)                                                                                  \
	do {                                                                                                    \
		if (opsMask == OPS_STROKE_AA && uniJoin == VGB_JOIN_ROUND && uniCap == VGB_CAP_ROUND) VGB_EMIT_K(T, OPS_STROKE_AA, 1, 1); \
		else if (opsMask == OPS_STROKE_AA && uniJoin == VGB_JOIN_MITER && uniCap == VGB_CAP_BUTT) VGB_EMIT_K(T, OPS_STROKE_AA, 0, 0); \
		else if (opsMask == OPS_STROKE_AA) VGB_EMIT_K(T, OPS_STROKE_AA, 3, 3); \
		else if (opsMask == OPS_FILL_AA) VGB_EMIT_K(T, OPS_FILL_AA, 3, 3); \
		else VGB_EMIT_K(T, OPS_ALL, 3, 3);                                 \
	} while (0)
			if (fillOnly) {
				const uint32_t fb = (numPieces + 255) / 256;
				if (idxBytes == 2 && sseFill) k_fill_emit<uint16_t, true><<<fb, 256, 0, st>>>(pa);
				else if (idxBytes == 2) k_fill_emit<uint16_t, false><<<fb, 256, 0, st>>>(pa);
				else if (sseFill) k_fill_emit<uint32_t, true><<<fb, 256, 0, st>>>(pa);
				else k_fill_emit<uint32_t, false><<<fb, 256, 0, st>>>(pa);
			} else if (idxBytes == 2) {
				VGB_LAUNCH_EMIT(uint16_t);
			} else {
				VGB_LAUNCH_EMIT(uint32_t);
			}
#undef VGB_LAUNCH_EMIT
#undef VGB_EMIT_K
			++launches;
			VGB_CUDA(cudaGetLastError());
			if (finalRan) {
				if (dUV.ensure(totals[0] * 4 + 32, err)) return VGB_ERR_CUDA;
				const uint32_t uvw = (uint32_t)(uint16_t)fcfg.uv[0] | ((uint32_t)(uint16_t)fcfg.uv[1] << 16);
				k_fill_u32<<<148 * 8, 256, 0, st>>>(dUV.as<uint32_t>(), totals[0], uvw);
				++launches;
				VGB_CUDA(cudaGetLastError());
			}
		}
		VGB_CUDA(cudaEventRecord(ev[3], st));
		sz.kernel_launches = launches;
		ran = true;
		kNumPts = numPts;
		kNumPieces = numPieces;
		kTotals[0] = totals[0];
		kTotals[1] = totals[1];
		sizesKnown = true;
		return 0;
	}

	int mergedInfo(uint64_t out[4], void** rankTotalsDev, std::string* err) override
	{
		if (!ran || !mergedRan) {
			if (err) *err = "no merged run yet";
			return VGB_ERR_STATE;
		}
		VGB_CUDA(cudaSetDevice(dev));
		if (readback(out, dRankBase.p, 32, err)) return VGB_ERR_CUDA;
		if (rankTotalsDev) *rankTotalsDev = dRankTotals.p;
		return 0;
	}

	/* single-pass runs of known sizes: did the kernel raise a flag (never expected)? */
	int checkSp(std::string* err)
	{
		if (!spUnchecked) return 0;
		spUnchecked = false;
		uint32_t f = 0;
		if (readback(&f, dTicket.as<uint32_t>() + 1, 4, err)) return VGB_ERR_CUDA;
		if (f) {
			spMode = 0; /* this batch object goes back to the two-pass pipeline */
			ran = false;
			if (err) *err = "single-pass emit did not complete (flags " + std::to_string(f) + "): run again";
			return VGB_ERR_CUDA;
		}
		return 0;
	}

	int sync(std::string* err) override
	{
		VGB_CUDA(cudaSetDevice(dev));
		VGB_CUDA(cudaStreamSynchronize(st));
		if (const int rc = checkSp(err)) return rc;
		if (ran) {
			cudaEventElapsedTime(&rt.total, ev[0], ev[3]);
			cudaEventElapsedTime(&rt.flatten, ev[0], ev[1]);
			cudaEventElapsedTime(&rt.strokeCount, ev[1], ev[2]);
			cudaEventElapsedTime(&rt.emit, ev[2], ev[3]);
		}
		return 0;
	}

	int mark(uint32_t slot, std::string* err) override
	{
		if (slot >= 8) {
			if (err) *err = "event slot out of range";
			return VGB_ERR_INVALID;
		}
		VGB_CUDA(cudaSetDevice(dev));
		VGB_CUDA(cudaEventRecord(userEv[slot], st));
		return 0;
	}

	int elapsed(uint32_t a, uint32_t b, float* ms, std::string* err) override
	{
		if (a >= 8 || b >= 8) {
			if (err) *err = "event slot out of range";
			return VGB_ERR_INVALID;
		}
		VGB_CUDA(cudaSetDevice(dev));
		VGB_CUDA(cudaEventSynchronize(userEv[b]));
		VGB_CUDA(cudaEventElapsedTime(ms, userEv[a], userEv[b]));
		return 0;
	}

	int setFinal(const vgb_final_cfg* cfg, std::string* err) override
	{
		if (!cfg) {
			finalOn = false;
			return 0;
		}
		if (cfg->max_vb_vertices < 2 || cfg->max_vb_vertices > 65536 || cfg->vb_count > cfg->max_vb_vertices || cfg->cmd_vertices > cfg->vb_count) {
			if (err) *err = "final layout: need 2 <= max_vb_vertices <= 65536 and cmd_vertices <= vb_count <= max_vb_vertices";
			return VGB_ERR_INVALID;
		}
		fcfg = *cfg;
		finalOn = true;
		return 0;
	}

	int downloadFinal(int16_t* uv, uint64_t vcap, vgb_draw_cmd* cmds, uint32_t cmdCap, uint32_t* numCmds, void** uvDev, std::string* err) override
	{
		if (!ran || (!finalRan && sz.num_vertices)) {
			if (err) *err = "download_final needs a run with the final layout set";
			return VGB_ERR_STATE;
		}
		VGB_CUDA(cudaSetDevice(dev));
		if (uvDev) *uvDev = dUV.p;
		if (numCmds) *numCmds = 0;
		if (!sz.num_vertices) return 0;
		FinalHdr h;
		if (readback(&h, dSegHdr.p, sizeof(h), err)) return VGB_ERR_CUDA;
		if (h.oversize) {
			if (err) *err = "a single mesh has max_vb_vertices vertices or more (vg.cpp:5323)";
			return VGB_ERR_CAPACITY;
		}
		/* a leading command that received only empty meshes carries nothing: drop it (ids stay as they are) */
		std::vector<vgb_draw_cmd> tmp(h.numCmds);
		if (h.numCmds) VGB_CUDA(cudaMemcpyAsync(tmp.data(), dSegCmds.p, (size_t)h.numCmds * sizeof(vgb_draw_cmd), cudaMemcpyDeviceToHost, st));
		VGB_CUDA(cudaStreamSynchronize(st));
		uint32_t n = 0;
		for (uint32_t i = 0; i < h.numCmds; ++i) {
			if (tmp[i].num_vertices == 0) continue;
			if (cmds) {
				if (n >= cmdCap) {
					if (err) *err = "draw-command capacity too small";
					return VGB_ERR_CAPACITY;
				}
				cmds[n] = tmp[i];
			}
			++n;
		}
		if (numCmds) *numCmds = n;
		if (uv) {
			if (vcap < sz.num_vertices) {
				if (err) *err = "vertex capacity too small";
				return VGB_ERR_CAPACITY;
			}
			VGB_CUDA(cudaMemcpyAsync(uv, dUV.p, (size_t)sz.num_vertices * 4, cudaMemcpyDeviceToHost, st));
		}
		VGB_CUDA(cudaStreamSynchronize(st));
		return 0;
	}

	int replay(const float* xforms, uint32_t n, const float* inv, std::string* err) override
	{
		if (!ran) {
			if (err) *err = "replay before run";
			return VGB_ERR_STATE;
		}
		if (finalRan) {
			if (err) *err = "replay of a final-layout run is not supported: run with plain uint16 or uint32 indices";
			return VGB_ERR_STATE;
		}
		VGB_CUDA(cudaSetDevice(dev));
		const uint64_t V = sz.num_vertices, I = sz.num_indices;
		replayN = n;
		replayRan = true;
		if (!n || !V) return 0;
		if (dRXf.ensure(((size_t)n + 1) * 24, err) || dRPos.ensure(V * n * 8 + 32, err) || dRCol.ensure(V * n * 4 + 32, err) ||
		    dRIdx.ensure(I * n * idxBytes + 32, err))
			return VGB_ERR_CUDA;
		VGB_CUDA(cudaMemcpyAsync(dRXf.p, xforms, (size_t)n * 24, cudaMemcpyHostToDevice, st));
		const float* dInv = nullptr;
		if (inv) {
			VGB_CUDA(cudaMemcpyAsync(dRXf.as<float>() + 6 * (size_t)n, inv, 24, cudaMemcpyHostToDevice, st));
			dInv = dRXf.as<float>() + 6 * (size_t)n;
		}
		const uint32_t rows = (n + REPLAY_CHUNK - 1) / REPLAY_CHUNK;
		dim3 grid((unsigned)((V + 255) / 256), rows < 65535u ? rows : 65535u);
		k_replay_vertices<<<grid, 256, 0, st>>>(dPos.as<float2>(), dCol.as<uint32_t>(), V, dRXf.as<float>(), n, dInv, dRPos.as<float2>(), dRCol.as<uint32_t>());
		if (I) {
			const unsigned gy = rows < 65535u ? rows : 65535u;
			if (idxBytes == 4) {
				k_replay_indices<uint32_t, true><<<dim3((unsigned)((I + 255) / 256), gy), 256, 0, st>>>(dIdx.as<uint32_t>(), I, V, n, dRIdx.as<uint32_t>());
			} else if (I % 2 == 0) {
				k_replay_indices<uint32_t, false><<<dim3((unsigned)((I / 2 + 255) / 256), gy), 256, 0, st>>>(dIdx.as<uint32_t>(), I / 2, V, n, dRIdx.as<uint32_t>());
			} else {
				k_replay_indices<uint16_t, false><<<dim3((unsigned)((I + 255) / 256), gy), 256, 0, st>>>(dIdx.as<uint16_t>(), I, V, n, dRIdx.as<uint16_t>());
			}
		}
		VGB_CUDA(cudaGetLastError());
		return 0;
	}

	int replaySizes(vgb_sizes* out, std::string* err) override
	{
		if (!replayRan) {
			if (err) *err = "no replay yet";
			return VGB_ERR_STATE;
		}
		*out = sz;
		out->num_vertices = sz.num_vertices * replayN;
		out->num_indices = sz.num_indices * replayN;
		out->num_meshes = sz.num_meshes * replayN;
		out->kernel_launches = sz.num_indices ? 2 : 1;
		return 0;
	}

	void replayPtrs(void** pos, void** color, void** index) override
	{
		if (pos) *pos = dRPos.p;
		if (color) *color = dRCol.p;
		if (index) *index = dRIdx.p;
	}

	int downloadReplay(float* pos, uint32_t* color, void* index, uint64_t vcap, uint64_t icap, std::string* err) override
	{
		if (!replayRan) {
			if (err) *err = "no replay yet";
			return VGB_ERR_STATE;
		}
		VGB_CUDA(cudaSetDevice(dev));
		const uint64_t V = sz.num_vertices * replayN, I = sz.num_indices * replayN;
		if (((pos || color) && vcap < V) || (index && icap < I)) {
			if (err) *err = "capacity too small";
			return VGB_ERR_CAPACITY;
		}
		if (pos && V) VGB_CUDA(cudaMemcpyAsync(pos, dRPos.p, V * 8, cudaMemcpyDeviceToHost, st));
		if (color && V) VGB_CUDA(cudaMemcpyAsync(color, dRCol.p, V * 4, cudaMemcpyDeviceToHost, st));
		if (index && I) VGB_CUDA(cudaMemcpyAsync(index, dRIdx.p, I * idxBytes, cudaMemcpyDeviceToHost, st));
		VGB_CUDA(cudaStreamSynchronize(st));
		return 0;
	}

	void sizes(vgb_sizes* out) override { *out = sz; }
	RunTimes times() override { return rt; }
	void* stream() override { return (void*)st; }

	void devicePtrs(void** pos, void** color, void** index, void** mvo, void** mio, void** pts, void** subs) override
	{
		if (pos) *pos = dPos.p;
		if (color) *color = dCol.p;
		if (index) *index = dIdx.p;
		if (mvo) *mvo = dMeshOff.p; /* interleaved {vtx, idx} u64 pairs */
		if (mio) *mio = dMeshOff.p ? (char*)dMeshOff.p + 8 : nullptr;
		if (pts) *pts = dPts.p;
		if (subs) *subs = dSubFirst.p;
	}

	int download(float* pos, uint32_t* color, void* index, uint64_t vcap, uint64_t icap, uint64_t* mvo, uint64_t* mio, std::string* err) override
	{
		if (!ran) {
			if (err) *err = "download before run";
			return VGB_ERR_STATE;
		}
		VGB_CUDA(cudaSetDevice(dev));
		if (const int rc = checkSp(err)) return rc;
		if ((pos || color) && vcap < sz.num_vertices) {
			if (err) *err = "vertex capacity too small";
			return VGB_ERR_CAPACITY;
		}
		if (index && icap < sz.num_indices) {
			if (err) *err = "index capacity too small";
			return VGB_ERR_CAPACITY;
		}
		if (pos && sz.num_vertices) VGB_CUDA(cudaMemcpyAsync(pos, dPos.p, sz.num_vertices * 8, cudaMemcpyDeviceToHost, st));
		if (color && sz.num_vertices) VGB_CUDA(cudaMemcpyAsync(color, dCol.p, sz.num_vertices * 4, cudaMemcpyDeviceToHost, st));
		if (index && sz.num_indices) VGB_CUDA(cudaMemcpyAsync(index, dIdx.p, sz.num_indices * idxBytes, cudaMemcpyDeviceToHost, st));
		if (mvo || mio) {
			const uint32_t n = numJobs + 1;
			if (dMeshSplit.ensure((size_t)n * 16, err)) return VGB_ERR_CUDA;
			uint64_t* dv = dMeshSplit.as<uint64_t>();
			uint64_t* di = dv + n;
			k_split_mesh_off<<<(n + 255) / 256, 256, 0, st>>>(dMeshOff.as<ulonglong2>(), n, dv, di);
			VGB_CUDA(cudaGetLastError());
			if (mvo) VGB_CUDA(cudaMemcpyAsync(mvo, dv, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
			if (mio) VGB_CUDA(cudaMemcpyAsync(mio, di, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
		}
		VGB_CUDA(cudaStreamSynchronize(st));
		return 0;
	}

	int downloadPaths(float* pts, uint64_t ptsCap, vgb_subpath* subs, uint32_t subCap, std::string* err) override
	{
		if (!ran) {
			if (err) *err = "download before run";
			return VGB_ERR_STATE;
		}
		VGB_CUDA(cudaSetDevice(dev));
		if ((pts && ptsCap < sz.num_points) || (subs && subCap < numSub)) {
			if (err) *err = "capacity too small";
			return VGB_ERR_CAPACITY;
		}
		const void* srcPts = dPts.p;
		const void* srcFirst = dSubFirst.p;
		if (slotsUsed && numSub) {
			/* the run used the slotted (non-dense) point buffer: hand out a dense copy, in sub-path order */
			if (dDenseFirst.ensure((size_t)(numSub + 1) * 4, err) || dDensePts.ensure(sz.num_points * 8 + 16, err)) return VGB_ERR_CUDA;
			if (scan<ScanOps<uint32_t>>(dSubNum.as<uint32_t>(), numSub, dDenseFirst.as<uint32_t>(), err)) return VGB_ERR_CUDA;
			k_gather_points<<<(numSub + 127) / 128, 128, 0, st>>>(dPts.as<float2>(), dSubFirst.as<uint32_t>(), dSubNum.as<uint32_t>(),
			                                                       dDenseFirst.as<uint32_t>(), numSub, dDensePts.as<float2>());
			VGB_CUDA(cudaGetLastError());
			srcPts = dDensePts.p;
			srcFirst = dDenseFirst.p;
		}
		if (pts && sz.num_points) VGB_CUDA(cudaMemcpyAsync(pts, srcPts, sz.num_points * 8, cudaMemcpyDeviceToHost, st));
		if (subs && numSub) {
			std::vector<uint32_t> f(numSub), n(numSub);
			std::vector<uint8_t> c(numSub);
			VGB_CUDA(cudaMemcpyAsync(f.data(), srcFirst, (size_t)numSub * 4, cudaMemcpyDeviceToHost, st));
			VGB_CUDA(cudaMemcpyAsync(n.data(), dSubNum.p, (size_t)numSub * 4, cudaMemcpyDeviceToHost, st));
			VGB_CUDA(cudaMemcpyAsync(c.data(), dSubClosed.p, (size_t)numSub, cudaMemcpyDeviceToHost, st));
			VGB_CUDA(cudaStreamSynchronize(st));
			for (uint32_t s = 0; s < numSub; ++s) {
				subs[s].first_vertex = f[s];
				subs[s].num_vertices = n[s];
				subs[s].is_closed = c[s];
			}
		}
		VGB_CUDA(cudaStreamSynchronize(st));
		return 0;
	}
};

Pipeline* Pipeline::create(int device, std::string* err)
{
	PipelineImpl* p = new PipelineImpl();
	if (p->init(device, err) != 0) {
		delete p;
		return nullptr;
	}
	return p;
}

int deviceCount()
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

void* hostAlloc(size_t bytes)
{
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
	return p;
}

void hostFree(void* p)
{
	if (p) cudaFreeHost(p);
}

int copyToHost(void* dst, const void* srcDev, size_t bytes)
{
	return cudaMemcpy(dst, srcDev, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : VGB_ERR_CUDA;
}
} // namespace vgb
