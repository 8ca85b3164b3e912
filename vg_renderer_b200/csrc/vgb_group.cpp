/*
 * vgb_group.cpp -- several GPUs, one merged vertex/index space (the layout contract is
 * vg.cpp:5207-5244: indices of a mesh are rebased by the vertices that precede it).
 *
 * One process per GPU.  Paths shard across ranks by range and every rank runs the whole pipeline on its
 * own paths with no communication -- except when ONE merged buffer is requested.  Then there is exactly
 * one exchange step per run: an all-gather of each rank's (vertex count, index count), 16 bytes per rank,
 * whose exclusive prefix is the rank's base (vgb_batch_run_merged: ncclAllGather on the batch's stream
 * between the count stage and the emit kernel; the base never visits the host).  Optionally the payload
 * is gathered into one owner's buffer at those bases (vgb_group_gather: ncclSend / ncclRecv, i.e. peer
 * copies over NVLink / NVSwitch); that step is NVLink-bound, not HBM-bound, and is timed separately.
 *
 * NCCL is resolved at run time (dlopen "libnccl.so.2"): single-GPU users of libvgb200.so need no NCCL,
 * and inside a torch.distributed process the already-loaded NCCL is reused.
 */
#include "vgb_internal.h"
#include "vgb.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <string.h>
#include <string>
#include <vector>

#if __has_include(<nccl.h>)
#include <nccl.h>
#else
/* the handful of declarations this file needs (NCCL 2.x ABI) */
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
typedef int ncclDataType_t;
#define ncclSuccess 0
#define ncclUint8 1
#define ncclUint64 5
#endif

namespace
{
struct Nccl
{
	void* lib = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	std::string err;

	bool load()
	{
		if (lib) return true;
		const char* names[] = {"libnccl.so.2", "libnccl.so"};
		for (const char* n : names) {
			lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
			if (lib) break;
		}
		if (!lib) {
			err = "NCCL not found (dlopen libnccl.so.2): the merged multi-GPU buffer needs it";
			return false;
		}
#define VGB_SYM(field, name)                                  \
	field = (decltype(field))dlsym(lib, name);                \
	if (!field) {                                             \
		err = std::string("NCCL symbol missing: ") + name;    \
		lib = nullptr;                                        \
		return false;                                         \
	}
		VGB_SYM(GetUniqueId, "ncclGetUniqueId")
		VGB_SYM(CommInitRank, "ncclCommInitRank")
		VGB_SYM(CommDestroy, "ncclCommDestroy")
		VGB_SYM(AllGather, "ncclAllGather")
		VGB_SYM(Send, "ncclSend")
		VGB_SYM(Recv, "ncclRecv")
		VGB_SYM(GroupStart, "ncclGroupStart")
		VGB_SYM(GroupEnd, "ncclGroupEnd")
		VGB_SYM(GetErrorString, "ncclGetErrorString")
#undef VGB_SYM
		return true;
	}
};

Nccl& nccl()
{
	static Nccl n;
	return n;
}
}

struct vgb_group
{
	ncclComm_t comm = nullptr;
	int device = 0, rank = 0, nranks = 1;
	vgb::MergeHook hook;
	/* the owner's merged buffers (vgb_group_gather) */
	void* gPos = nullptr;
	void* gCol = nullptr;
	void* gIdx = nullptr;
	size_t capPos = 0, capCol = 0, capIdx = 0;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace vgb
{
std::string g_groupErr;

static int groupAllGather(void* user, const void* send, void* recv, size_t count, void* stream, std::string* err)
{
	vgb_group* g = (vgb_group*)user;
	const ncclResult_t r = nccl().AllGather(send, recv, count, ncclUint64, g->comm, (cudaStream_t)stream);
	if (r != ncclSuccess) {
		if (err) *err = std::string("ncclAllGather: ") + nccl().GetErrorString(r);
		return VGB_ERR_CUDA;
	}
	return 0;
}

int groupUniqueId(void* id128, std::string* err)
{
	if (!id128) {
		if (err) *err = "null id";
		return VGB_ERR_INVALID;
	}
	if (!nccl().load()) {
		if (err) *err = nccl().err;
		return VGB_ERR_STATE;
	}
	ncclUniqueId id;
	const ncclResult_t r = nccl().GetUniqueId(&id);
	if (r != ncclSuccess) {
		if (err) *err = std::string("ncclGetUniqueId: ") + nccl().GetErrorString(r);
		return VGB_ERR_CUDA;
	}
	static_assert(sizeof(id) == VGB_GROUP_ID_BYTES, "NCCL unique id size");
	memcpy(id128, &id, sizeof(id));
	return 0;
}

int groupCreate(vgb_group** out, int device, int rank, int nranks, const void* id128, std::string* err)
{
	if (!out || !id128 || nranks < 1 || rank < 0 || rank >= nranks) {
		if (err) *err = "group: need 0 <= rank < nranks and the shared 128-byte id";
		return VGB_ERR_INVALID;
	}
	if (!nccl().load()) {
		if (err) *err = nccl().err;
		return VGB_ERR_STATE;
	}
	if (cudaSetDevice(device) != cudaSuccess) {
		if (err) *err = "group: cudaSetDevice failed";
		return VGB_ERR_CUDA;
	}
	vgb_group* g = new vgb_group();
	g->device = device;
	g->rank = rank;
	g->nranks = nranks;
	ncclUniqueId id;
	memcpy(&id, id128, sizeof(id));
	const ncclResult_t r = nccl().CommInitRank(&g->comm, nranks, id, rank);
	if (r != ncclSuccess) {
		if (err) *err = std::string("ncclCommInitRank: ") + nccl().GetErrorString(r);
		delete g;
		return VGB_ERR_CUDA;
	}
	g->hook.rank = rank;
	g->hook.nranks = nranks;
	g->hook.user = g;
	g->hook.allGatherU64 = groupAllGather;
	cudaEventCreate(&g->ev0);
	cudaEventCreate(&g->ev1);
	*out = g;
	return 0;
}

void groupDestroy(vgb_group* g)
{
	if (!g) return;
	cudaSetDevice(g->device);
	if (g->gPos) cudaFree(g->gPos);
	if (g->gCol) cudaFree(g->gCol);
	if (g->gIdx) cudaFree(g->gIdx);
	if (g->ev0) cudaEventDestroy(g->ev0);
	if (g->ev1) cudaEventDestroy(g->ev1);
	if (g->comm) nccl().CommDestroy(g->comm);
	delete g;
}

const MergeHook* groupHook(vgb_group* g) { return g ? &g->hook : nullptr; }

static int ensureBuf(void** p, size_t* cap, size_t bytes, std::string* err)
{
	if (bytes <= *cap) return 0;
	if (*p) cudaFree(*p);
	*p = nullptr;
	*cap = 0;
	if (cudaMalloc(p, bytes + 256) != cudaSuccess) {
		if (err) *err = "group gather: cudaMalloc of the merged buffer failed";
		return VGB_ERR_CUDA;
	}
	*cap = bytes + 256;
	return 0;
}

/* Payload placement: every rank's pos / colour / index arrays land in the owner's buffers at the rank's
 * base.  Sizes come from the all-gathered totals of the merged run (read back once here: the host has to
 * post the receives).  One NCCL group of point-to-point transfers; the owner's own share is a local copy. */
int groupGather(vgb_group* g, Pipeline* pipe, uint32_t idxBytes, int owner, void** pos, void** color, void** index, uint64_t info[4], float* ms,
                std::string* err)
{
	if (!g || !pipe || owner < 0 || owner >= g->nranks) {
		if (err) *err = "group gather: bad arguments";
		return VGB_ERR_INVALID;
	}
	if (cudaSetDevice(g->device) != cudaSuccess) return VGB_ERR_CUDA;
	uint64_t base[4];
	void* dTotals = nullptr;
	int rc = pipe->mergedInfo(base, &dTotals, err);
	if (rc) return rc;
	std::vector<uint64_t> tot(2 * (size_t)g->nranks);
	if (cudaMemcpy(tot.data(), dTotals, tot.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
		if (err) *err = "group gather: reading the rank totals failed";
		return VGB_ERR_CUDA;
	}
	void *lp = nullptr, *lc = nullptr, *li = nullptr;
	pipe->devicePtrs(&lp, &lc, &li, nullptr, nullptr, nullptr, nullptr);
	cudaStream_t st = (cudaStream_t)pipe->stream();
	const uint64_t V = base[2], I = base[3];
	if (g->rank == owner) {
		if (ensureBuf(&g->gPos, &g->capPos, V * 8, err) || ensureBuf(&g->gCol, &g->capCol, V * 4, err) || ensureBuf(&g->gIdx, &g->capIdx, I * idxBytes, err))
			return VGB_ERR_CUDA;
	}
	cudaEventRecord(g->ev0, st);
	ncclResult_t r = nccl().GroupStart();
	uint64_t bv = 0, bi = 0;
	for (int k = 0; k < g->nranks && r == ncclSuccess; ++k) {
		const uint64_t v = tot[2 * k], i = tot[2 * k + 1];
		if (g->rank == owner) {
			char* dp = (char*)g->gPos + bv * 8;
			char* dc = (char*)g->gCol + bv * 4;
			char* di = (char*)g->gIdx + bi * idxBytes;
			if (k == owner) {
				if (v) {
					cudaMemcpyAsync(dp, lp, v * 8, cudaMemcpyDeviceToDevice, st);
					cudaMemcpyAsync(dc, lc, v * 4, cudaMemcpyDeviceToDevice, st);
				}
				if (i) cudaMemcpyAsync(di, li, i * idxBytes, cudaMemcpyDeviceToDevice, st);
			} else {
				if (v) {
					r = nccl().Recv(dp, v * 8, ncclUint8, k, g->comm, st);
					if (r == ncclSuccess) r = nccl().Recv(dc, v * 4, ncclUint8, k, g->comm, st);
				}
				if (i && r == ncclSuccess) r = nccl().Recv(di, i * idxBytes, ncclUint8, k, g->comm, st);
			}
		} else if (k == g->rank) {
			if (v) {
				r = nccl().Send(lp, v * 8, ncclUint8, owner, g->comm, st);
				if (r == ncclSuccess) r = nccl().Send(lc, v * 4, ncclUint8, owner, g->comm, st);
			}
			if (i && r == ncclSuccess) r = nccl().Send(li, i * idxBytes, ncclUint8, owner, g->comm, st);
		}
		bv += v;
		bi += i;
	}
	const ncclResult_t r2 = nccl().GroupEnd();
	if (r == ncclSuccess) r = r2;
	if (r != ncclSuccess) {
		if (err) *err = std::string("group gather: ") + nccl().GetErrorString(r);
		return VGB_ERR_CUDA;
	}
	cudaEventRecord(g->ev1, st);
	if (cudaStreamSynchronize(st) != cudaSuccess) {
		if (err) *err = "group gather: stream error";
		return VGB_ERR_CUDA;
	}
	if (ms) cudaEventElapsedTime(ms, g->ev0, g->ev1);
	if (info) memcpy(info, base, 32);
	if (pos) *pos = g->rank == owner ? g->gPos : nullptr;
	if (color) *color = g->rank == owner ? g->gCol : nullptr;
	if (index) *index = g->rank == owner ? g->gIdx : nullptr;
	return 0;
}
} // namespace vgb
