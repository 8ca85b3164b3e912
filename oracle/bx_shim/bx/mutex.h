/* stand-in for <bx/mutex.h> (vg.cpp:383-385, 5040-5042: the data-pool mutex shared with bgfx's render thread;
 * the oracle is single-threaded) */
#ifndef BX_SHIM_MUTEX_H
#define BX_SHIM_MUTEX_H
#include "bx.h"
namespace bx
{
class Mutex
{
public:
	void lock() {}
	void unlock() {}
};
class MutexScope
{
public:
	explicit MutexScope(Mutex&) {}
};
}
#endif
