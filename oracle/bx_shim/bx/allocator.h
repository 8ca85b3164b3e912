/* Stand-in for <bx/allocator.h>: one-virtual allocator interface + the free functions the
 * reference calls (path.cpp:25-41,752; stroker.cpp:196-229,820,2325-2339). TEST INFRASTRUCTURE. */
#ifndef BX_SHIM_ALLOCATOR_H
#define BX_SHIM_ALLOCATOR_H

#include "bx.h"
#include <stdlib.h>

namespace bx
{
struct AllocatorI
{
	virtual ~AllocatorI() {}
	/* ptr==0: allocate; size==0: free; else resize. align==0 means natural alignment. */
	virtual void* realloc(void* ptr, size_t size, size_t align, const char* file, uint32_t line) = 0;
};

inline void* alloc(AllocatorI* a, size_t size, size_t align = 0) { return a->realloc(nullptr, size, align, nullptr, 0); }
inline void free(AllocatorI* a, void* ptr, size_t align = 0) { if (ptr) a->realloc(ptr, 0, align, nullptr, 0); }
inline void* realloc(AllocatorI* a, void* ptr, size_t size, size_t align = 0) { return a->realloc(ptr, size, align, nullptr, 0); }

/* aligned variants: over-allocate and stash the offset in the 4 bytes before the block */
inline void* alignedAlloc(AllocatorI* a, size_t size, size_t align)
{
	const size_t total = size + align + sizeof(uint32_t);
	uint8_t* base = (uint8_t*)a->realloc(nullptr, total, 0, nullptr, 0);
	uintptr_t p = ((uintptr_t)base + sizeof(uint32_t) + (align - 1)) & ~(uintptr_t)(align - 1);
	((uint32_t*)p)[-1] = (uint32_t)(p - (uintptr_t)base);
	return (void*)p;
}
inline void alignedFree(AllocatorI* a, void* ptr, size_t /*align*/)
{
	if (!ptr) return;
	uint8_t* base = (uint8_t*)ptr - ((uint32_t*)ptr)[-1];
	a->realloc(base, 0, 0, nullptr, 0);
}
inline void* alignedRealloc(AllocatorI* a, void* ptr, size_t size, size_t align)
{
	if (!ptr) return alignedAlloc(a, size, align);
	uint8_t* base = (uint8_t*)ptr - ((uint32_t*)ptr)[-1];
	const uint32_t oldOff = ((uint32_t*)ptr)[-1];
	const size_t total = size + align + sizeof(uint32_t);
	uint8_t* nb = (uint8_t*)a->realloc(base, total, 0, nullptr, 0);
	uintptr_t p = ((uintptr_t)nb + sizeof(uint32_t) + (align - 1)) & ~(uintptr_t)(align - 1);
	const uint32_t newOff = (uint32_t)(p - (uintptr_t)nb);
	if (newOff != oldOff) memMove((void*)p, nb + oldOff, size);
	((uint32_t*)p)[-1] = newOff;
	return (void*)p;
}

/* placement-new helpers vg.cpp uses for its Context (vg.cpp:735, 1015) */
struct PlacementNewTag {};
}
inline void* operator new(size_t, bx::PlacementNewTag, void* p) { return p; }
inline void operator delete(void*, bx::PlacementNewTag, void*) throw() {}
#define BX_NEW(_allocator, _type) ::new (bx::PlacementNewTag(), bx::alloc(_allocator, sizeof(_type))) _type
namespace bx
{
template<typename T> inline void deleteObject(AllocatorI* a, T* obj, size_t align = 0)
{
	if (obj) {
		obj->~T();
		free(a, obj, align);
	}
}

/* malloc-backed default, used by the oracle's C wrapper */
struct DefaultAllocator : AllocatorI
{
	void* realloc(void* ptr, size_t size, size_t, const char*, uint32_t) override
	{
		if (size == 0) { ::free(ptr); return nullptr; }
		return ::realloc(ptr, size);
	}
};
}

#endif
