// capi_handle.cpp — libStringSearchLib_h.so: the AS-SHIPPED handle-keyed C ABI (include/StringSearchLib_handle.h)
// over the engine core. Replaces /root/reference/nGramSearch/dllmain.cpp:22-151 one export at a time.
// The registry lock is NOT held while the GPU builds (the reference blocks every search for the whole build,
// dllmain.cpp:39); searches on one index are serialised by the engine, different indexes run concurrently.
#include "../../include/StringSearchLib_handle.h"
#include "../../include/ssb_engine.h"
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <unordered_map>

namespace {
struct IndexDeleter { void operator()(ssb_index* p) const { ssb_index_destroy(p); } };
std::shared_mutex g_lock;                                                 // dllmain.cpp:22
std::unordered_map<uint32_t, std::shared_ptr<ssb_index>> g_indexed;       // dllmain.cpp:24

std::shared_ptr<ssb_index> find(uint32_t handle)
{
	std::shared_lock<std::shared_mutex> lk(g_lock);
	auto it = g_indexed.find(handle);
	return it == g_indexed.end() ? nullptr : it->second;
}

uint32_t run(uint32_t handle, const char* query, char*** results, float** scores, float threshold, uint32_t limit)
{
	auto ix = find(handle);
	if (!ix || !query || !results) return 0;                              // dllmain.cpp:64-69: unknown handle => 0, outputs untouched
	uint64_t info[24];
	ssb_index_info(ix.get(), info);
	if (!info[12]) return 0;                                              // not indexed (nGramSearch.hpp:417,451)
	const void* q = query;
	ssb_batch* b = ssb_search_one(ix.get(), q, threshold, limit);      // concurrent callers coalesce into one GPU batch
	if (!b) return 0;
	const uint32_t n = b->offsets[1];
	// nGramSearch.hpp:428-436 / 462-468: arrays from new[], strings alias index storage
	char** out = new char*[n];
	for (uint32_t i = 0; i < n; ++i) out[i] = const_cast<char*>(static_cast<const char*>(b->strings[i]));
	*results = out;
	if (scores) {
		float* sc = new float[n];
		memcpy(sc, b->scores, n * sizeof(float));
		*scores = sc;
	}
	ssb_batch_free(b);
	return n;
}
}   // namespace

#define DLLEXP extern "C" __attribute__((visibility("default")))

DLLEXP uint32_t indexN(char** const words, const uint64_t size, const uint16_t rowSize, float* const weight)
{
	try {
		std::shared_ptr<ssb_index> ix(ssb_index_create(reinterpret_cast<const void* const*>(words), size, rowSize, weight, 3, 0, 0),
		                              IndexDeleter());
		if (!ix) return 0;
		std::unique_lock<std::shared_mutex> lk(g_lock);
		uint32_t handle = 1;                                              // 0 is reserved (dllmain.cpp:40-46)
		const uint32_t maxVal = std::numeric_limits<uint32_t>::max();
		while (g_indexed.find(handle) != g_indexed.end() && handle < maxVal) handle++;
		if (handle == maxVal) return 0;
		g_indexed.emplace(handle, std::move(ix));
		return handle;
	} catch (...) { return 0; }
}
DLLEXP uint32_t search(uint32_t handle, const char* query, char*** results, float threshold, uint32_t limit)
{
	try { return run(handle, query, results, nullptr, threshold, limit); } catch (...) { return 0; }
}
DLLEXP uint32_t score(uint32_t handle, const char* query, char*** results, float** scores, float threshold, uint32_t limit)
{
	try { return scores ? run(handle, query, results, scores, threshold, limit) : 0; } catch (...) { return 0; }   // dllmain.cpp:82-90 (UB there on unknown handle; 0 here)
}
DLLEXP void release(uint32_t, char** results, float* scores)
{
	delete[] results;                                                     // nGramSearch.hpp:477-483 (freed even after dispose: no leak)
	delete[] scores;
}
DLLEXP void dispose(uint32_t handle)
{
	std::shared_ptr<ssb_index> victim;
	{
		std::unique_lock<std::shared_mutex> lk(g_lock);
		auto it = g_indexed.find(handle);
		if (it == g_indexed.end()) return;                                // dllmain.cpp:113: unknown handle ignored
		victim = std::move(it->second);
		g_indexed.erase(it);
	}
}
DLLEXP uint64_t getSize(uint32_t handle)
{
	auto ix = find(handle);
	return ix ? ssb_index_size(ix.get()) : 0;
}
DLLEXP uint64_t getLibSize(uint32_t handle)
{
	auto ix = find(handle);
	return ix ? ssb_index_lib_size(ix.get()) : 0;
}
DLLEXP void setValidChar(uint32_t handle, char* const characters, int n)
{
	auto ix = find(handle);
	if (ix && characters && n >= 0) ssb_index_set_valid_char(ix.get(), characters, n);
}
