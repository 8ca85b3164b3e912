/*
 * Stand-in for <bx/math.h>.  TEST INFRASTRUCTURE ONLY (oracle).
 * Two numerics variants:
 *   default          : vg_renderer_b200/csrc/vgb_math.h -- explicit fp32 routines, the SAME
 *                      code the CUDA kernels use (parity by construction).
 *   -DBX_SHIM_LIBM=1 : glibc libm (sinf/cosf/acosf/atan2f/tanf) -- used only to quantify how
 *                      sensitive vertex counts / positions are to the choice of math library.
 * Neither is bx's own code: bx is absent and unpinned, hence "parity unpinned" w.r.t. bx.
 */
#ifndef BX_SHIM_MATH_H
#define BX_SHIM_MATH_H

#include "bx.h"
#include <math.h>
#include "vgb_math.h"

namespace bx
{
constexpr float kPi = 3.1415926535897932384626433832795f;
constexpr float kPi2 = 6.2831853071795864769252867665590f;
constexpr float kPiHalf = 1.5707963267948966192313216916398f;
constexpr float kInvPi = 1.0f / 3.1415926535897932384626433832795f;

inline float abs(float a) { return ::fabsf(a); }
inline float sqrt(float a) { return ::sqrtf(a); }
inline float rsqrt(float a) { return 1.0f / ::sqrtf(a); }
inline float ceil(float a) { return ::ceilf(a); }
inline float floor(float a) { return ::floorf(a); }
inline float sign(float a) { return (float)((a > 0.0f) - (a < 0.0f)); }
inline float square(float a) { return a * a; }
inline float mod(float a, float b) { return ::fmodf(a, b); }
#if defined(BX_SHIM_LIBM) && BX_SHIM_LIBM
inline float sin(float a) { return ::sinf(a); }
inline float cos(float a) { return ::cosf(a); }
inline float tan(float a) { return ::tanf(a); }
inline float acos(float a) { return ::acosf(a); }
inline float atan2(float y, float x) { return ::atan2f(y, x); }
#else
inline float sin(float a) { return vgb_sin(a); }
inline float cos(float a) { return vgb_cos(a); }
inline float tan(float a) { return vgb_tan(a); }
inline float acos(float a) { return vgb_acos(a); }
inline float atan2(float y, float x) { return vgb_atan2(y, x); }
#endif

inline float exp(float a) { return ::expf(a); }
inline float pow(float a, float b) { return ::powf(a, b); }
inline float log(float a) { return ::logf(a); }
inline float round(float a) { return ::roundf(a); }
inline float trunc(float a) { return ::truncf(a); }
inline float fract(float a) { return a - ::truncf(a); }
inline float toRad(float d) { return d * kPi / 180.0f; }
/* column-major 4x4, as bgfx expects (only used for the view/projection of vg::end, which the stub ignores) */
inline void mtxIdentity(float* m)
{
	memset(m, 0, sizeof(float) * 16);
	m[0] = m[5] = m[10] = m[15] = 1.0f;
}
inline void mtxOrtho(float* m, float l, float r, float b, float t, float n, float f, float offset, bool homogeneousNdc)
{
	const float aa = 2.0f / (r - l), bb = 2.0f / (t - b), cc = (homogeneousNdc ? 2.0f : 1.0f) / (f - n);
	const float dd = (l + r) / (l - r), ee = (t + b) / (b - t), ff = homogeneousNdc ? (n + f) / (n - f) : n / (n - f);
	memset(m, 0, sizeof(float) * 16);
	m[0] = aa; m[5] = bb; m[10] = cc; m[12] = dd + offset; m[13] = ee; m[14] = ff; m[15] = 1.0f;
}

template<typename T> inline T min(const T& a, const T& b) { return a < b ? a : b; }
template<typename T> inline T max(const T& a, const T& b) { return a > b ? a : b; }
template<typename T> inline T min(const T& a, const T& b, const T& c) { return min(min(a, b), c); }
template<typename T> inline T max(const T& a, const T& b, const T& c) { return max(max(a, b), c); }
template<typename T> inline T clamp(const T& a, const T& lo, const T& hi) { return min(max(a, lo), hi); } /* bx order */
}

#endif
