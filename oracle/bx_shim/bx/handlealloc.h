/* stand-in for <bx/handlealloc.h>: the dense id allocator vg.cpp uses for images, fonts and command lists
 * (vg.cpp:375-381, 741-760).  alloc() hands out the lowest free id, like bx::HandleAlloc's free list start. */
#ifndef BX_SHIM_HANDLEALLOC_H
#define BX_SHIM_HANDLEALLOC_H
#include "bx.h"
#include "allocator.h"
#include <new>
namespace bx
{
constexpr uint16_t kInvalidHandle = UINT16_MAX;
class HandleAlloc
{
public:
	explicit HandleAlloc(uint16_t maxHandles) : m_max(maxHandles), m_num(0), m_used(nullptr) {}
	uint16_t alloc()
	{
		for (uint16_t i = 0; i < m_max; ++i) {
			if (!m_used[i]) {
				m_used[i] = 1;
				++m_num;
				return i;
			}
		}
		return kInvalidHandle;
	}
	bool isValid(uint16_t h) const { return h < m_max && m_used[h] != 0; }
	void free(uint16_t h)
	{
		if (isValid(h)) {
			m_used[h] = 0;
			--m_num;
		}
	}
	uint16_t getNumHandles() const { return m_num; }
	uint16_t getMaxHandles() const { return m_max; }
	void reset()
	{
		memset(m_used, 0, m_max);
		m_num = 0;
	}
	uint16_t m_max, m_num;
	uint8_t* m_used;
};
inline HandleAlloc* createHandleAlloc(AllocatorI* a, uint16_t maxHandles)
{
	uint8_t* mem = (uint8_t*)a->realloc(nullptr, sizeof(HandleAlloc) + maxHandles, 0, nullptr, 0);
	HandleAlloc* h = ::new (mem) HandleAlloc(maxHandles);
	h->m_used = mem + sizeof(HandleAlloc);
	h->reset();
	return h;
}
inline void destroyHandleAlloc(AllocatorI* a, HandleAlloc* h)
{
	if (h) a->realloc(h, 0, 0, nullptr, 0);
}
}
#endif
