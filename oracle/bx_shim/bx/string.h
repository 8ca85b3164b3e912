/* stand-in for <bx/string.h>: vg.cpp only needs strLen (text entry points) */
#ifndef BX_SHIM_STRING_H
#define BX_SHIM_STRING_H
#include "bx.h"
#include <stdio.h>
#include <stdarg.h>
namespace bx
{
inline int32_t strLen(const char* s, int32_t max = INT32_MAX)
{
	if (!s) return 0;
	int32_t n = 0;
	while (n < max && s[n]) ++n;
	return n;
}
inline int32_t snprintf(char* out, int32_t max, const char* fmt, ...)
{
	va_list a;
	va_start(a, fmt);
	const int32_t n = ::vsnprintf(out, max, fmt, a);
	va_end(a);
	return n;
}
}
#endif
