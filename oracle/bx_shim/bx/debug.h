#ifndef BX_SHIM_DEBUG_H
#define BX_SHIM_DEBUG_H
#include <stdio.h>
#include <stdarg.h>
namespace bx
{
inline void debugPrintf(const char* fmt, ...) { va_list a; va_start(a, fmt); vfprintf(stderr, fmt, a); va_end(a); }
inline void debugBreak() { __builtin_trap(); }
}
#endif
