/*
 * Minimal stand-in for bkaradzic/bx's <bx/bx.h>, sufficient to compile the reference's
 * src/path.cpp, src/stroker.cpp, src/vg_util.cpp and src/libtess2/ unmodified, from where
 * they lie under /root/reference (see oracle/Makefile).  TEST INFRASTRUCTURE ONLY.
 * bx itself is not vendored in the reference and not pinned; nothing here is
 * copied from bx -- only the names the reference uses are provided.
 */
#ifndef BX_SHIM_BX_H
#define BX_SHIM_BX_H

#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define BX_CPU_X86 1
#define BX_CONFIG_SUPPORTS_THREADING 0
#define BX_FORCE_INLINE inline __attribute__((always_inline))
#define BX_ALIGN_DECL(_align, _decl) _decl __attribute__((aligned(_align)))
#define BX_STRINGIZE_(_x) #_x
#define BX_STRINGIZE(_x) BX_STRINGIZE_(_x)
#define BX_FILE_LINE_LITERAL "" __FILE__ "(" BX_STRINGIZE(__LINE__) "): "
#define BX_PRAGMA_DIAGNOSTIC_IGNORED_MSVC(_x)
#define BX_PRAGMA_DIAGNOSTIC_IGNORED_CLANG_GCC(_x)
#define BX_PRAGMA_DIAGNOSTIC_IGNORED_GCC(_x)
#define BX_PRAGMA_DIAGNOSTIC_IGNORED_CLANG(_x)

namespace bx
{
template<typename... Ts> inline void unusedArgs(Ts&&...) {}
inline void memCopy(void* dst, const void* src, size_t n) { ::memcpy(dst, src, n); }
inline void memSet(void* dst, uint8_t v, size_t n) { ::memset(dst, v, n); }
inline void memMove(void* dst, const void* src, size_t n) { ::memmove(dst, src, n); }
inline uint32_t uint32_max(uint32_t a, uint32_t b) { return a > b ? a : b; }
inline uint32_t uint32_min(uint32_t a, uint32_t b) { return a < b ? a : b; }
inline int32_t memCmp(const void* a, const void* b, size_t n) { return ::memcmp(a, b, n); }
template<typename T> inline bool isPowerOf2(T a) { return a && !(a & (a - 1)); }
/* strided copy (vg.cpp: text atlas upload) */
inline void gather(void* dst, const void* src, uint32_t srcStride, uint32_t stride, uint32_t num)
{
	uint8_t* d = (uint8_t*)dst;
	const uint8_t* s = (const uint8_t*)src;
	for (uint32_t i = 0; i < num; ++i, d += stride, s += srcStride) ::memcpy(d, s, stride);
}
}

#define BX_UNUSED(...) ::bx::unusedArgs(__VA_ARGS__)

/* the few simd names src/vg_util.cpp's text-quad routine mentions (never called by the oracle;
 * present only so the VG_CONFIG_ENABLE_SIMD=1 build links) */
#include <xmmintrin.h>
#include <emmintrin.h>
namespace bx
{
typedef __m128 simd128_t;
inline simd128_t simd_splat(float a) { return _mm_set1_ps(a); }
inline simd128_t simd_ld(const void* p) { return _mm_loadu_ps((const float*)p); }
inline void simd_st(void* p, simd128_t a) { _mm_storeu_ps((float*)p, a); }
inline simd128_t simd_add(simd128_t a, simd128_t b) { return _mm_add_ps(a, b); }
inline simd128_t simd_mul(simd128_t a, simd128_t b) { return _mm_mul_ps(a, b); }
inline simd128_t simd_shuf_xyAB(simd128_t a, simd128_t b) { return _mm_movelh_ps(a, b); }
inline simd128_t simd_shuf_zwCD(simd128_t a, simd128_t b) { return _mm_movehl_ps(b, a); }
#define BX_SHIM_SWIZ(_name, _x, _y, _z, _w) \
	inline simd128_t simd_swiz_##_name(simd128_t a) { return _mm_shuffle_ps(a, a, _MM_SHUFFLE(_w, _z, _y, _x)); }
BX_SHIM_SWIZ(xxyy, 0, 0, 1, 1)
BX_SHIM_SWIZ(xyyx, 0, 1, 1, 0)
BX_SHIM_SWIZ(xzxz, 0, 2, 0, 2)
BX_SHIM_SWIZ(xzyw, 0, 2, 1, 3)
BX_SHIM_SWIZ(ywyw, 1, 3, 1, 3)
BX_SHIM_SWIZ(zwwz, 2, 3, 3, 2)
BX_SHIM_SWIZ(zzww, 2, 2, 3, 3)
#undef BX_SHIM_SWIZ
}

#endif
