/*
 * vgb_math.h -- explicit fp32 scalar math shared bit-for-bit by
 *   (1) the CUDA kernels (device, __f*_rn intrinsics: never FMA-contracted),
 *   (2) the host side of libvgb200 (per-call constants of the vg:: wrappers),
 *   (3) the oracle's bx shim and C restatement (oracle/), compiled with -ffp-contract=off.
 *
 * Why this file exists: the reference takes cos/sin/acos/atan2/tan/rsqrt from bkaradzic/bx
 * (reference call sites: src/path.cpp:232,245-246,260-268,307-313,602-613,654-661;
 * src/stroker.cpp:36,1398-1399,1476-1481,1588-1594,1626,1744-1750,1919-1925).  bx is neither
 * vendored nor pinned, and these functions feed INTEGER decisions (ceil(pi/da),
 * (uint32)(dAngle/da)) that set vertex counts and therefore index buffers.  This build
 * therefore defines the functions itself, using only IEEE-754 binary32 add/sub/mul/div/sqrt
 * plus explicit fma in a fixed order, so CPU oracle and GPU kernels agree to the last bit by construction.
 * Coefficients: tools/fit_math.py (near-minimax fits; accuracy checked in tests/test_math.py).
 *
 * C99 / C++ / CUDA compatible.  The only fused operations are the EXPLICIT fma steps of the
 * polynomial kernels below (IEEE fused multiply-add: exactly defined, so still reproducible);
 * nothing is ever contracted implicitly.
 */
#ifndef VGB_MATH_H
#define VGB_MATH_H

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VGB_HD __host__ __device__ __forceinline__
#else
#define VGB_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define VGB_MUL(a, b) __fmul_rn((a), (b))
#define VGB_ADD(a, b) __fadd_rn((a), (b))
#define VGB_SUB(a, b) __fsub_rn((a), (b))
#define VGB_DIV(a, b) __fdiv_rn((a), (b))
#define VGB_SQRT(a) __fsqrt_rn((a))
#define VGB_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
/* host: plain operators; translation units including this header must be built with
 * -ffp-contract=off (and no -ffast-math) so nothing is fused or re-associated behind our back.
 * fmaf() is the exactly-rounded fused multiply-add of C99 (one vfmadd with -mfma). */
#define VGB_MUL(a, b) ((a) * (b))
#define VGB_ADD(a, b) ((a) + (b))
#define VGB_SUB(a, b) ((a) - (b))
#define VGB_DIV(a, b) ((a) / (b))
#define VGB_SQRT(a) sqrtf((a))
#define VGB_FMA(a, b, c) fmaf((a), (b), (c))
#endif

#define VGB_PI 3.1415926535897932384626433832795f
#define VGB_PI2 6.2831853071795864769252867665590f
#define VGB_PI_HALF 1.5707963267948966192313216916398f
#define VGB_EPSILON 1e-5f /* VG_EPSILON, include/vg/vg.h:100 */

/* ---- helpers with exact semantics -------------------------------------------------- */
VGB_HD float vgb_abs(float a) { return fabsf(a); }
VGB_HD float vgb_sign(float a) { return (float)((a > 0.0f) - (a < 0.0f)); }
VGB_HD float vgb_rsqrt(float a) { return VGB_DIV(1.0f, VGB_SQRT(a)); }
VGB_HD float vgb_ceil(float a) { return ceilf(a); }
VGB_HD float vgb_floor(float a) { return floorf(a); }
VGB_HD uint32_t vgb_umax(uint32_t a, uint32_t b) { return a > b ? a : b; }

/* float -> uint32 as the reference's `(uint32_t)x` behaves in its x86-64 build (cvttss2si to a 64-bit
 * register, low half kept): truncation for |x| < 2^63, and 0 for NaN, infinities and anything larger (the
 * "integer indefinite" 0x8000000000000000).  Out-of-range conversions are undefined in C and SATURATE on
 * the GPU (inf -> 0xFFFFFFFF), so a degenerate stroke (width below the fringe: da = 0, pi/da = inf) would
 * otherwise get 2 cap points on the CPU and 4 billion on the device.  Every float -> count conversion of
 * the path goes through here, on both sides. */
VGB_HD uint32_t vgb_f2u(float x)
{
	if (!(fabsf(x) < 9.2233720368547758e18f)) {
		return 0u;
	}
	return (uint32_t)(unsigned long long)(long long)x;
}

/* ---- sin / cos ---------------------------------------------------------------------
 * Polynomials are Horner chains of explicit fused multiply-adds: each step is ONE IEEE-754
 * fma, so the value is a pure function of the inputs on any conforming machine. */
/* |r| <= pi/4 kernels */
VGB_HD float vgb__sin_k(float r)
{
	const float z = VGB_MUL(r, r);
	float p = VGB_FMA(-0x1.9aaefcp-13f, z, 0x1.110bf6p-7f);
	p = VGB_FMA(p, z, -0x1.555552p-3f);
	return VGB_FMA(VGB_MUL(r, z), p, r);
}

VGB_HD float vgb__cos_k(float r)
{
	const float z = VGB_MUL(r, r);
	float p = VGB_FMA(0x1.9bc328p-16f, z, -0x1.6c12aap-10f);
	p = VGB_FMA(p, z, 0x1.555554p-5f);
	return VGB_FMA(VGB_MUL(z, z), p, VGB_FMA(-0.5f, z, 1.0f));
}

/* Cody-Waite reduction by pi/2 = C1 + C2 + C3, three fused steps; quadrant in *q.
 * Accurate for |x| up to a few thousand (the path only produces |x| < 10). */
VGB_HD float vgb__reduce(float x, int* q)
{
	const float kf = rintf(VGB_MUL(x, 0.63661977236758134308f));
	float r = VGB_FMA(kf, -1.5703125f, x);
	r = VGB_FMA(kf, -4.837512969970703125e-4f, r);
	r = VGB_FMA(kf, -7.54978995489188216e-8f, r);
	*q = (int)kf;
	return r;
}

VGB_HD float vgb_sin(float x)
{
	int q;
	const float r = vgb__reduce(x, &q);
	const float v = (q & 1) ? vgb__cos_k(r) : vgb__sin_k(r);
	return (q & 2) ? -v : v;
}

VGB_HD float vgb_cos(float x)
{
	int q;
	const float r = vgb__reduce(x, &q);
	const float v = (q & 1) ? vgb__sin_k(r) : vgb__cos_k(r);
	return ((q + 1) & 2) ? -v : v;
}

/* both at once (one reduction); bit-identical to vgb_sin / vgb_cos */
VGB_HD void vgb_sincos(float x, float* s, float* c)
{
	int q;
	const float r = vgb__reduce(x, &q);
	const float sk = vgb__sin_k(r);
	const float ck = vgb__cos_k(r);
	const float sv = (q & 1) ? ck : sk;
	const float cv = (q & 1) ? sk : ck;
	*s = (q & 2) ? -sv : sv;
	*c = ((q + 1) & 2) ? -cv : cv;
}

VGB_HD float vgb_tan(float x) { return VGB_DIV(vgb_sin(x), vgb_cos(x)); }

/* ---- atan2 -------------------------------------------------------------------------
 * atan(t) on the whole first octant t in [0,1] with one odd polynomial (|err| < 2e-8), so
 * atan2 costs a single division and no data-dependent branch. */
VGB_HD float vgb__atan01(float t)
{
	const float z = VGB_MUL(t, t);
	float p = VGB_FMA(-0x1.404d1cp-9f, z, 0x1.c21b1cp-7f);
	p = VGB_FMA(p, z, -0x1.28ec28p-5f);
	p = VGB_FMA(p, z, 0x1.01503p-4f);
	p = VGB_FMA(p, z, -0x1.633408p-4f);
	p = VGB_FMA(p, z, 0x1.c419eep-4f);
	p = VGB_FMA(p, z, -0x1.246f5cp-3f);
	p = VGB_FMA(p, z, 0x1.99985ap-3f);
	p = VGB_FMA(p, z, -0x1.555554p-2f);
	return VGB_FMA(VGB_MUL(t, z), p, t);
}

/* atan2(y, x) in (-pi, pi]; atan2(0, 0) is defined as 0 regardless of zero signs. */
VGB_HD float vgb_atan2(float y, float x)
{
	const float ax = fabsf(x);
	const float ay = fabsf(y);
	const float mx = ax > ay ? ax : ay;
	const float mn = ax > ay ? ay : ax;
	if (mx == 0.0f) {
		return 0.0f;
	}
	float a = vgb__atan01(VGB_DIV(mn, mx));
	if (ay > ax) {
		a = VGB_SUB(VGB_PI_HALF, a);
	}
	if (x < 0.0f) {
		a = VGB_SUB(VGB_PI, a);
	}
	return y < 0.0f ? -a : a;
}

/* ---- acos -------------------------------------------------------------------------- */
/* asin(x) for |x| <= 0.5 */
VGB_HD float vgb__asin_k(float x)
{
	const float z = VGB_MUL(x, x);
	float p = VGB_FMA(0x1.3cffcp-5f, z, 0x1.ae779ep-6f);
	p = VGB_FMA(p, z, 0x1.70fb0ap-5f);
	p = VGB_FMA(p, z, 0x1.33253ep-4f);
	p = VGB_FMA(p, z, 0x1.55556p-3f);
	return VGB_FMA(VGB_MUL(x, z), p, x);
}

/* acos(x), x clamped to [-1, 1] */
VGB_HD float vgb_acos(float x)
{
	const float ax = fabsf(x);
	if (ax >= 1.0f) {
		return x > 0.0f ? 0.0f : VGB_PI;
	}
	if (ax <= 0.5f) {
		return VGB_SUB(VGB_PI_HALF, vgb__asin_k(x));
	}
	/* acos(|x|) = 2 asin(sqrt((1-|x|)/2)) */
	const float s = VGB_SQRT(VGB_MUL(0.5f, VGB_SUB(1.0f, ax)));
	const float h = VGB_MUL(2.0f, vgb__asin_k(s));
	return x > 0.0f ? h : VGB_SUB(VGB_PI, h);
}

#endif /* VGB_MATH_H */
