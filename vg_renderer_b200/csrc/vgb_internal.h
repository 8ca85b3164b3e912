/*
 * vgb_internal.h -- interface between the host-side recorder / C ABI (vgb_batch.cpp,
 * vgb_objects.cpp) and the device pipeline (vgb_device.cu).  Not installed.
 */
#ifndef VGB_INTERNAL_H
#define VGB_INTERNAL_H

#include "vgb.h"
#include <stddef.h>
#include <stdint.h>
#include <string>
#include <vector>

namespace vgb
{
/* one sub-path's slice of the command stream: {cmdBegin, cmdEnd, dataBegin, pathId} */
struct SubRec
{
	uint32_t cmdBegin, cmdEnd, dataBegin, pathId;
};

/* one draw = one strokerXXX configuration applied to every sub-path of one path */
struct DrawRec
{
	uint32_t ocj;   /* op | cap << 8 | join << 16 */
	uint32_t color;
	float width;
	uint32_t pathId;
};

/* per-draw constants derived on the device by k_draw_prep (stroker.cpp:1393-1399 etc.) */
struct DrawDerived
{
	float hsw, hsw_aa, da, fringe;
	uint32_t nHC, color, c0, ocj;
};

/* one job (mesh) as the piece kernels read it: 48 bytes, three 16-byte loads */
struct JobRec
{
	uint32_t first, N, flags, pieceStart; /* flags = closed | op << 8 | cap << 16 | join << 24 */
	float hsw, hsw_aa, da, fringe;
	uint32_t nHC, color, c0, pad;
};

/* everything the host hands to a run (all host pointers; sizes in elements) */
struct HostTables
{
	/* flatten stage input */
	const SubRec* subRec = nullptr;
	uint32_t numSub = 0;
	const uint8_t* cmd = nullptr;
	uint32_t numCmd = 0;
	const float* data = nullptr;
	uint32_t numData = 0;
	const float* pathScale = nullptr;
	const float* pathTol = nullptr;
	const float* pathXform = nullptr; /* nullable: identity for every path */
	uint32_t numPaths = 0;
	/* single-pass flatten: sub-path s owns point slots [subSlotFirst[s], subSlotFirst[s+1]) sized from
	 * its commands by the recorder (numSub+1 entries; nullptr = always use the two-pass flatten) */
	const uint32_t* subSlotFirst = nullptr;
	/* raw-polyline mode (strokerXXX on caller polylines): replaces the flatten stage */
	const float* rawPts = nullptr;
	uint32_t numRawPts = 0;
	const uint32_t* rawFirst = nullptr;
	const uint32_t* rawNum = nullptr;
	const uint8_t* rawClosed = nullptr;
	/* bulk mode (vgb_batch_set_cubic_strokes): n single-cubic paths, tables generated on the device */
	const float* bulkP8 = nullptr;
	uint32_t bulkN = 0;
	float bulkScale = 1.0f, bulkTol = 0.25f;
	/* draws and jobs */
	const DrawRec* draws = nullptr;
	uint32_t numDraws = 0;
	/* per draw: scissor id | force-new-draw-command << 31 (nullptr: one scissor, no forced commands) */
	const uint32_t* drawScissor = nullptr;
	uint32_t numScissorBreaks = 0; /* draws carrying the force flag (bounds the draw-command table) */
	const uint32_t* jobSub = nullptr;
	const uint32_t* jobDraw = nullptr;
	uint32_t numJobs = 0;
	float fringe = 1.0f;
};

struct RunTimes
{
	float total = 0, flatten = 0, strokeCount = 0, emit = 0;
};

/* merged vertex/index space over several GPUs: the one exchange step of the path.  After the
 * count stage a rank knows its totals (V, I) -- on the device; the hook all-gathers the 16-byte pair of every
 * rank into rankTotals on the pipeline's own stream (vgb_group.cpp: ncclAllGather), a one-warp kernel turns
 * it into this rank's exclusive base, and the emit kernel adds that base to every index.  Nothing comes back
 * to the host in between. */
struct MergeHook
{
	int rank = 0, nranks = 1;
	void* user = nullptr;
	/* all-gather `count` uint64 from send (device) into recv (device, count * nranks) on `stream` */
	int (*allGatherU64)(void* user, const void* send, void* recv, size_t count, void* stream, std::string* err) = nullptr;
};

/* device pipeline: owns the stream, the device buffers and the kernels */
class Pipeline
{
public:
	static Pipeline* create(int device, std::string* err);
	virtual ~Pipeline() {}
	virtual int upload(const HostTables& t, std::string* err) = 0;
	virtual int run(uint32_t flags, uint64_t vertexBase, std::string* err, const MergeHook* merge = nullptr) = 0;
	/* after a merged run: {vertex base, index base, total vertices, total indices} of the group, and the device
	 * array of every rank's (V, I) (2 * nranks uint64); synchronises the stream */
	virtual int mergedInfo(uint64_t out[4], void** rankTotalsDev, std::string* err) = 0;
	virtual int sync(std::string* err) = 0;
	virtual void sizes(vgb_sizes* out) = 0;
	virtual void devicePtrs(void** pos, void** color, void** index, void** mvo, void** mio, void** pts, void** subs) = 0;
	virtual int download(float* pos, uint32_t* color, void* index, uint64_t vcap, uint64_t icap,
	                     uint64_t* mvo, uint64_t* mio, std::string* err) = 0;
	virtual int downloadPaths(float* pts, uint64_t ptsCap, vgb_subpath* subs, uint32_t subCap, std::string* err) = 0;
	virtual int setFinal(const vgb_final_cfg* cfg, std::string* err) = 0;
	virtual int downloadFinal(int16_t* uv, uint64_t vcap, vgb_draw_cmd* cmds, uint32_t cmdCap, uint32_t* numCmds, void** uvDev, std::string* err) = 0;
	virtual int replay(const float* xforms, uint32_t n, const float* inv, std::string* err) = 0;
	virtual int replaySizes(vgb_sizes* out, std::string* err) = 0;
	virtual void replayPtrs(void** pos, void** color, void** index) = 0;
	virtual int downloadReplay(float* pos, uint32_t* color, void* index, uint64_t vcap, uint64_t icap, std::string* err) = 0;
	virtual RunTimes times() = 0;
	virtual void* stream() = 0;
	virtual int mark(uint32_t slot, std::string* err) = 0;
	virtual int elapsed(uint32_t a, uint32_t b, float* ms, std::string* err) = 0;
};

int deviceCount();
/* vgb_group.cpp */
int groupUniqueId(void* id128, std::string* err);
int groupCreate(vgb_group** out, int device, int rank, int nranks, const void* id128, std::string* err);
void groupDestroy(vgb_group* g);
const MergeHook* groupHook(vgb_group* g);
int groupGather(vgb_group* g, Pipeline* pipe, uint32_t idxBytes, int owner, void** pos, void** color, void** index, uint64_t info[4], float* ms,
                std::string* err);
void* hostAlloc(size_t bytes);
void hostFree(void* p);
int copyToHost(void* dst, const void* srcDev, size_t bytes);
} // namespace vgb

#endif
