// capi_guid.cpp — libStringSearchLib.so: the README guid-keyed C ABI (include/StringSearchLib.h) over the engine
// core: index/indexW/index2D/index2DW, search/searchW, release/releaseW, dispose/disposeW, getSize[W]/getLibSize[W].
// Citations: /root/reference/Readme.md:<line>. The snapshot has no source for this surface; behaviour where the
// README is silent: re-index of an existing guid replaces it; narrow and wide libraries share one guid namespace;
// result strings are copies owned by the caller until release[W]().
#include "../../include/StringSearchLib.h"
#include "../../include/ssb_engine.h"
#include <cstring>
#include <cwchar>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace {
struct IndexDeleter { void operator()(ssb_index* p) const { ssb_index_destroy(p); } };
std::shared_mutex g_lock;
std::unordered_map<std::string, std::shared_ptr<ssb_index>> g_libs;

std::shared_ptr<ssb_index> find(const char* guid)
{
	if (!guid) return nullptr;
	std::shared_lock<std::shared_mutex> lk(g_lock);
	auto it = g_libs.find(guid);
	return it == g_libs.end() ? nullptr : it->second;
}
void publish(const char* guid, ssb_index* raw)
{
	if (!guid) { ssb_index_destroy(raw); return; }
	std::shared_ptr<ssb_index> ix(raw, IndexDeleter());
	std::shared_ptr<ssb_index> old;
	std::unique_lock<std::shared_mutex> lk(g_lock);
	if (!raw) return;                       // CUDA failure: the guid keeps whatever it had
	auto& slot = g_libs[guid];
	old = std::move(slot);
	slot = std::move(ix);
}
template <class C>
void do_index(const char* guid, C** const key, uint64_t size, uint16_t rowSize, float* const weight, uint16_t gSize)
{
	try {
		publish(guid, ssb_index_create(reinterpret_cast<const void* const*>(key), size, rowSize, weight, gSize, sizeof(C) == 4, 0));
	} catch (...) {}
}
// 2D input (Readme.md:31,47): key[r][c], weight[r][c]; size = number of rows. Flattened, then the 1D path.
template <class C>
void do_index2d(const char* guid, C*** const key, uint64_t size, uint16_t rowSize, float** const weight, uint16_t gSize)
{
	try {
		if (!key || rowSize == 0) { publish(guid, ssb_index_create(nullptr, 0, rowSize, nullptr, gSize, sizeof(C) == 4, 0)); return; }
		std::vector<const void*> flat(size * rowSize, nullptr);
		std::vector<float> w;
		if (weight) w.assign(size * rowSize, 1.0f);
		for (uint64_t r = 0; r < size; ++r) {
			if (!key[r]) continue;
			for (uint16_t c = 0; c < rowSize; ++c) {
				flat[r * rowSize + c] = key[r][c];
				if (weight && weight[r]) w[r * rowSize + c] = weight[r][c];
			}
		}
		publish(guid, ssb_index_create(flat.data(), flat.size(), rowSize, weight ? w.data() : nullptr, gSize, sizeof(C) == 4, 0));
	} catch (...) {}
}
template <class C>
void do_search(const char* guid, const C* query, C*** results, uint32_t* size, float threshold, uint32_t limit)
{
	if (size) *size = 0;
	try {
		auto ix = find(guid);
		if (!ix || !query || !results || !size) return;
		uint64_t info[24];
		ssb_index_info(ix.get(), info);
		if (!info[12] || info[11] != sizeof(C)) return;
		const void* q = query;
		ssb_batch* b = ssb_search_one(ix.get(), q, threshold, limit);      // concurrent callers coalesce into one GPU batch
		if (!b) return;
		const uint32_t n = b->offsets[1];
		C** out = new C*[n];
		for (uint32_t i = 0; i < n; ++i) {
			const C* s = static_cast<const C*>(b->strings[i]);
			size_t len = 0;
			while (s[len]) ++len;
			out[i] = new C[len + 1];
			memcpy(out[i], s, (len + 1) * sizeof(C));
		}
		ssb_batch_free(b);
		*results = out;
		*size = n;
	} catch (...) {}
}
template <class C> void do_release(C*** results, uint64_t size)
{
	if (!results || !*results) return;
	for (uint64_t i = 0; i < size; ++i) delete[] (*results)[i];
	delete[] *results;
	*results = nullptr;
}
void do_dispose(const char* guid)
{
	if (!guid) return;
	std::shared_ptr<ssb_index> victim;
	std::unique_lock<std::shared_mutex> lk(g_lock);
	auto it = g_libs.find(guid);
	if (it == g_libs.end()) return;         // Readme.md:180: unknown library ignored
	victim = std::move(it->second);
	g_libs.erase(it);
}
}   // namespace

#define DLLEXP extern "C" __attribute__((visibility("default")))

DLLEXP void index2D(char* const guid, char*** const key, const uint64_t size, const uint16_t rowSize, float** const weight, const uint16_t gSize)
{ do_index2d<char>(guid, key, size, rowSize, weight, gSize); }
DLLEXP void index2DW(char* const guid, wchar_t*** const key, const uint64_t size, const uint16_t rowSize, float** const weight, const uint16_t gSize)
{ do_index2d<wchar_t>(guid, key, size, rowSize, weight, gSize); }
DLLEXP void StringSearchLib_index(char* const guid, char** const key, const uint64_t size, const uint16_t rowSize, float* const weight, const uint16_t gSize)
{ do_index<char>(guid, key, size, rowSize, weight, gSize); }   // exported as symbol "index" (see the header)
DLLEXP void indexW(char* const guid, wchar_t** const key, const uint64_t size, const uint16_t rowSize, float* const weight, const uint16_t gSize)
{ do_index<wchar_t>(guid, key, size, rowSize, weight, gSize); }
DLLEXP void search(char* const guid, const char* query, char*** results, uint32_t* size, const float threshold, uint32_t limit)
{ do_search<char>(guid, query, results, size, threshold, limit); }
DLLEXP void searchW(char* const guid, const wchar_t* query, wchar_t*** results, uint32_t* size, const float threshold, uint32_t limit)
{ do_search<wchar_t>(guid, query, results, size, threshold, limit); }
DLLEXP void release(char* const, char*** results, uint64_t size) { do_release<char>(results, size); }
DLLEXP void releaseW(char* const, wchar_t*** results, uint64_t size) { do_release<wchar_t>(results, size); }
DLLEXP void dispose(char* const guid) { do_dispose(guid); }
DLLEXP void disposeW(char* const guid) { do_dispose(guid); }
DLLEXP uint64_t getSize(char* const guid) { auto ix = find(guid); return ix ? ssb_index_size(ix.get()) : 0; }
DLLEXP uint64_t getSizeW(char* const guid) { auto ix = find(guid); return ix ? ssb_index_size(ix.get()) : 0; }
DLLEXP uint64_t getLibSize(char* const guid) { auto ix = find(guid); return ix ? ssb_index_lib_size(ix.get()) : 0; }
DLLEXP uint64_t getLibSizeW(char* const guid) { auto ix = find(guid); return ix ? ssb_index_lib_size(ix.get()) : 0; }
