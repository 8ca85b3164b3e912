/*
 * vg_oracle.h -- CPU restatement of the reference's path-flatten + stroke/fill algorithm.
 * TEST INFRASTRUCTURE ONLY (see vg_oracle.c).  Same entry points as oracle/ref_capi.cpp
 * (prefix vgo_ instead of vgref_), so oracle/pyref.py drives both identically.
 */
#ifndef VG_ORACLE_H
#define VG_ORACLE_H

#include <stdint.h>
#include "vgb.h" /* command / op ids and vgb_subpath */

#ifdef __cplusplus
extern "C" {
#endif

uint32_t vgo_variant(void);

int vgo_flatten(const uint8_t* cmd, uint32_t ncmd, const float* data, float scale, float tol,
                float* pts, uint32_t ptsCap, vgb_subpath* sp, uint32_t spCap, uint32_t* numPts, uint32_t* numSp);

int vgo_mesh(uint32_t op, const float* vtx, uint32_t n, int closed, uint32_t color, float width, uint32_t cap, uint32_t join,
             float scale, float tol, float fringe,
             float* pos, uint32_t* col, uint16_t* idx, uint32_t vCap, uint32_t iCap, uint32_t* nv, uint32_t* ni, int* hasColor);

void vgo_math_eval(uint32_t op, const float* x, const float* y, float* out, uint32_t n);

int vgo_run_batch(uint32_t pathBegin, uint32_t pathEnd,
                  const uint32_t* pathCmdOff, const uint32_t* pathDataOff, const uint8_t* cmd, const float* data,
                  const float* pathScale, const float* pathTol, const float* pathXform, float fringe,
                  const uint32_t* drawBegin, const uint32_t* drawOp, const uint32_t* drawColor, const float* drawWidth,
                  const uint8_t* drawCap, const uint8_t* drawJoin,
                  int store, float* pos, uint32_t* col, uint16_t* idx, uint64_t vCap, uint64_t iCap,
                  uint64_t* meshVtxOff, uint64_t* meshIdxOff, uint32_t meshCap, uint64_t* totals);

int vgo_final_layout(const uint64_t* meshVtxOff, const uint64_t* meshIdxOff, uint32_t numMeshes, const uint16_t* idxLocal,
                     const vgb_final_cfg* cfg, uint16_t* idxOut, int16_t* uvOut, vgb_draw_cmd* cmds, uint32_t cmdCap, uint32_t* numCmds);

int vgo_final_layout_sc(const uint64_t* meshVtxOff, const uint64_t* meshIdxOff, uint32_t numMeshes, const uint16_t* idxLocal,
                        const vgb_final_cfg* cfg, const uint32_t* meshScissor, uint16_t* idxOut, int16_t* uvOut, vgb_draw_cmd* cmds,
                        uint32_t cmdCap, uint32_t* numCmds);

void vgo_transform_positions(const float* pos, uint32_t n, const float* mtx, float* out);

#ifdef __cplusplus
}
#endif
#endif
