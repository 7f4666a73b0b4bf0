// engine.cu — host side of the engine core: packing of the caller's char**/wchar_t** into pinned CSR buffers,
// index objects, result marshalling, and the extern "C" API of include/ssb_engine.h.
// Mirrors the role of dllmain.cpp:22-151 + StringIndex::search/score/release (nGramSearch.hpp:415-483).
#include "../../include/ssb_engine.h"
#include "index.h"
#include "search.h"
#include "digest.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <cwchar>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

using namespace ssb;

namespace {

// trivially destructible on purpose: thread-local destructors run at the very start of exit(), BEFORE the static registries of
// the shim libraries destroy their remaining indexes through guarded() — a std::string here would be written to after its
// buffer was freed (glibc: "corrupted double-linked list" at interpreter exit)
thread_local char g_lastError[512] = {0};
inline void set_last_error(const char* msg)
{
	strncpy(g_lastError, msg ? msg : "", sizeof(g_lastError) - 1);
	g_lastError[sizeof(g_lastError) - 1] = 0;
}

struct EngineGlobals {
	std::once_flag once;
	int device = -1;
	void init()
	{
		std::call_once(once, [&] {
			if (device < 0) SSB_CUDA(cudaGetDevice(&device));
			SSB_CUDA(cudaSetDevice(device));
			cudaMemPool_t pool;
			SSB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
			uint64_t keep = ~0ull;   // keep freed blocks cached in the pool: no OS round trips between batches
			SSB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
			std::atexit([] { g_exiting.store(true); });
		});
		SSB_CUDA(cudaSetDevice(device));
	}
} g_engine;

unsigned host_threads()
{
	unsigned n = std::thread::hardware_concurrency();
	return std::max(1u, std::min(n ? n : 1u, 32u));
}

// f(chunk index, begin, end) over nt contiguous chunks of [0, n); returns the number of chunks
template <class F> unsigned parallel_for(uint64_t n, unsigned nt, F f)
{
	if (nt <= 1) { f(0u, (uint64_t)0, n); return 1; }
	std::vector<std::thread> th;
	uint64_t chunk = (n + nt - 1) / nt;
	unsigned used = 0;
	for (unsigned t = 0; t < nt; ++t) {
		uint64_t b = t * chunk, e = std::min(n, b + chunk);
		if (b >= e) break;
		th.emplace_back([=] { f(t, b, e); });
		++used;
	}
	for (auto& t : th) t.join();
	return used;
}

// Process-wide pinned staging for index(): reused across builds so that steady-state builds do not pay
// cudaMallocHost for a few hundred MB each time. Guarded by a mutex: one packing at a time.
struct PackStaging {
	std::mutex mu;
	PinBuf<uint64_t> off;
	PinBuf<uint8_t> isNull;
	PinBuf<uint8_t> chars;
	PinBuf<float> wts;
} g_staging;

inline size_t unit_len(const char* s) { return strlen(s); }
inline size_t unit_len(const wchar_t* s) { return wcslen(s); }

}   // namespace

// Pinned result slabs: the arrays an ssb_batch exposes are the D2H destinations themselves (no host-side copies or
// per-result gathers); ssb_batch_free() returns the slab to its pool.
struct Slab {
	void* base = nullptr;
	size_t bytes = 0;
	HostResult hr;
};
struct SlabPool {
	std::mutex mu;
	std::vector<Slab*> free;
	~SlabPool()
	{
		for (Slab* s : free) { if (!g_exiting.load()) cudaFreeHost(s->base); delete s; }
	}
	static size_t al(size_t x) { return (x + 63) & ~size_t(63); }
	Slab* get(size_t capQ, size_t capResults)
	{
		{
			std::lock_guard<std::mutex> lk(mu);
			for (size_t i = 0; i < free.size(); ++i)
				if (free[i]->hr.capQ >= capQ && free[i]->hr.capResults >= capResults) {
					Slab* s = free[i];
					free.erase(free.begin() + i);
					return s;
				}
		}
		auto* s = new Slab();
		capResults = std::max<size_t>(capResults, 1);
		const size_t oOff = 0, oKid = oOff + al((capQ + 1) * 4), oSc = oKid + al(capResults * 4), oRow = oSc + al(capResults * 4),
		             oStr = oRow + al(capResults * 8);
		s->bytes = oStr + al(capResults * 8);
		SSB_CUDA(cudaMallocHost(&s->base, s->bytes));
		auto* b = static_cast<uint8_t*>(s->base);
		s->hr.offsets = reinterpret_cast<uint32_t*>(b + oOff); s->hr.kids = reinterpret_cast<uint32_t*>(b + oKid);
		s->hr.scores = reinterpret_cast<float*>(b + oSc); s->hr.rows = reinterpret_cast<uint64_t*>(b + oRow);
		s->hr.strs = reinterpret_cast<uint64_t*>(b + oStr);
		s->hr.capQ = capQ; s->hr.capResults = capResults;
		return s;
	}
	void put(Slab* s)
	{
		std::lock_guard<std::mutex> lk(mu);
		if (free.size() >= 4) { cudaFreeHost(s->base); delete s; return; }
		free.push_back(s);
	}
};
struct BatchImpl {
	ssb_batch pub;                       // must stay the first member: ssb_batch* <-> BatchImpl*
	std::shared_ptr<SlabPool> pool;
	Slab* slab = nullptr;
};

struct ssb_index {
	std::shared_ptr<SlabPool> pool = std::make_shared<SlabPool>();
	DeviceIndex dix;
	HostKeys hk;
	SearchWorkspace ws;
	BuildStats bst;
	double msPack = 0;
	uint8_t validLut[256];
	cudaStream_t stream = nullptr;
	std::mutex mu;            // one batch at a time per index (the workspace is shared)
	PinBuf<uint8_t> pinQ;     // pinned staging of packed queries
	PinBuf<uint32_t> pinQOff;
	SearchStats lastStats;
	// coalescing front end of ssb_search_one: callers that arrive while a batch is running are answered together by the next one
	struct OneReq { const void* q; float th; uint32_t limit; ssb_batch* out = nullptr; bool done = false; };
	std::mutex oneMu;
	std::condition_variable oneCv;
	std::vector<OneReq*> onePending;
	bool oneBusy = false;
	std::atomic<uint64_t> oneCalls{0}, oneBatches{0};
	~ssb_index()
	{
		if (stream && !g_exiting.load()) { cudaStreamSynchronize(stream); }
		// device buffers are released by their destructors (stream-ordered); the stream itself is destroyed last
	}
};

namespace {

void default_valid(uint8_t lut[256])
{
	memset(lut, 0, 256);
	// nGramSearch.h:307-313
	for (const char* p = ".%$ @0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"; *p; ++p) lut[(unsigned char)*p] = 1;
}

template <class C>
ssb_index* create_from_pointers(const C* const* words, uint64_t size, uint16_t rowSize, const float* weight, uint16_t gSize,
                                uint64_t rowBase)
{
	g_engine.init();
	// (ssb_index_destroy, not delete, if the build throws: the stream has to be synchronised and destroyed too)
	std::unique_ptr<ssb_index, void (*)(ssb_index*)> ix(new ssb_index(), &ssb_index_destroy);
	default_valid(ix->validLut);
	SSB_CUDA(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
	auto t0 = std::chrono::steady_clock::now();
	PackedWords pw;
	DevBuf<uint8_t> dChars, dNull;                      // the words on the device (uploaded while they are packed); freed in stream order
	DevBuf<uint64_t> dOff;
	std::lock_guard<std::mutex> stagingLock(g_staging.mu);
	auto& off = g_staging.off;
	auto& isNull = g_staging.isNull;
	auto& charsBytes = g_staging.chars;
	auto& wts = g_staging.wts;
	if (words && size >= 2) {
		off.reserve(size + 1);
		isNull.reserve(size);
		uint64_t* o = off.get();
		uint8_t* nu = isNull.get();
		// pass 1: lengths (parallel strlen over the caller's pointers) as chunk-local inclusive prefix sums
		const unsigned nt = (unsigned)std::min<uint64_t>(host_threads(), std::max<uint64_t>(1, size / 65536));
		std::vector<uint64_t> chunkSum(nt + 1, 0);
		parallel_for(size, nt, [&](unsigned t, uint64_t b, uint64_t e) {
			uint64_t run = 0;
			for (uint64_t i = b; i < e; ++i) {
				const C* w = words[i];
				nu[i] = w == nullptr;
				run += w ? unit_len(w) : 0;
				o[i + 1] = run;
			}
			chunkSum[t + 1] = run;
		});
		o[0] = 0;
		for (unsigned t = 0; t < nt; ++t) chunkSum[t + 1] += chunkSum[t];
		const uint64_t total = chunkSum[nt];
		charsBytes.reserve((std::max<uint64_t>(total, 1) + size) * sizeof(C));   // + terminators: doubles as the key-table bounce buffer
		C* c = reinterpret_cast<C*>(charsBytes.get());
		// pass 2: globalise the offsets and gather the code units into one pinned CSR buffer. Every packer thread uploads its
		// part in a few pieces as it goes (all on the build's stream, in any order: the build's first kernel comes after all of
		// them), so the H2D copy of the words hides behind the packing instead of following it
		dChars.alloc(std::max<uint64_t>(total, 1) * sizeof(C), ix->stream);
		dOff.alloc(size + 1, ix->stream);
		dNull.alloc(size, ix->stream);
		SSB_CUDA(cudaMemcpyAsync(dNull.get(), nu, size, cudaMemcpyHostToDevice, ix->stream));
		SSB_CUDA(cudaMemcpyAsync(dOff.get(), o, sizeof(uint64_t), cudaMemcpyHostToDevice, ix->stream));      // o[0] = 0
		int dev = 0;
		SSB_CUDA(cudaGetDevice(&dev));
		std::vector<cudaError_t> upErr(nt, cudaSuccess);
		cudaStream_t up = ix->stream;
		uint8_t* dC = dChars.get();
		uint64_t* dO = dOff.get();
		parallel_for(size, nt, [&, dev, up, dC, dO](unsigned t, uint64_t b, uint64_t e) {
			cudaSetDevice(dev);
			const uint64_t base = chunkSum[t];
			uint64_t prev = base;      // == global o[b]
			const uint64_t piece = std::max<uint64_t>((e - b + 3) / 4, 1);
			for (uint64_t pb = b; pb < e; pb += piece) {
				const uint64_t pe = std::min(e, pb + piece), first = prev;
				for (uint64_t i = pb; i < pe; ++i) {
					const uint64_t end = o[i + 1] + base;
					if (words[i]) memcpy(c + prev, words[i], (end - prev) * sizeof(C));
					o[i + 1] = end;
					prev = end;
				}
				cudaError_t er = cudaSuccess;
				if (prev > first) er = cudaMemcpyAsync(dC + first * sizeof(C), c + first, (prev - first) * sizeof(C), cudaMemcpyHostToDevice, up);
				if (er == cudaSuccess) er = cudaMemcpyAsync(dO + pb + 1, o + pb + 1, (pe - pb) * sizeof(uint64_t), cudaMemcpyHostToDevice, up);
				if (er != cudaSuccess) upErr[t] = er;
			}
		});
		for (cudaError_t er : upErr) SSB_CUDA(er);
		pw.d_chars = dChars.get(); pw.d_off = dOff.get(); pw.d_isNull = dNull.get();
		if (weight) {
			wts.reserve(size);
			memcpy(wts.get(), weight, size * sizeof(float));
		}
		pw.chars = c; pw.off = o; pw.isNull = nu; pw.weight = weight ? wts.get() : nullptr; pw.n = size; pw.totalUnits = total;
		pw.pinnedScratch = charsBytes.get(); pw.pinnedScratchBytes = charsBytes.size();
	}
	ix->msPack = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	build_index(ix->dix, ix->hk, pw, rowSize, gSize, (int)sizeof(C), rowBase, ix->validLut, ix->stream, &ix->bst);
	return ix.release();
}

ssb_index* create_from_packed(const void* units, const uint64_t* offsets, const uint8_t* isNull, uint64_t size, uint16_t rowSize,
                              const float* weight, uint16_t gSize, int unitBytes, uint64_t rowBase)
{
	g_engine.init();
	// (ssb_index_destroy, not delete, if the build throws: the stream has to be synchronised and destroyed too)
	std::unique_ptr<ssb_index, void (*)(ssb_index*)> ix(new ssb_index(), &ssb_index_destroy);
	default_valid(ix->validLut);
	SSB_CUDA(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
	PackedWords pw;
	std::vector<uint8_t> zeros;
	if (units && offsets && size >= 2) {
		if (!isNull) { zeros.assign(size, 0); isNull = zeros.data(); }
		pw.chars = units; pw.off = offsets; pw.isNull = isNull; pw.weight = weight; pw.n = size; pw.totalUnits = offsets[size];
	}
	build_index(ix->dix, ix->hk, pw, rowSize, gSize, unitBytes, rowBase, ix->validLut, ix->stream, &ix->bst);
	return ix.release();
}

ssb_batch* marshal(ssb_index* ix, const BatchOutput& out)
{
	auto* impl = new BatchImpl();
	ssb_batch* b = &impl->pub;
	b->nq = out.nq;
	b->offsets = new uint32_t[out.nq + 1];
	uint64_t total = 0;
	for (uint32_t q = 0; q < out.nq; ++q) { b->offsets[q] = (uint32_t)total; total += out.count[q]; }
	b->offsets[out.nq] = (uint32_t)total;
	b->kids = new uint32_t[total ? total : 1];
	b->scores = new float[total ? total : 1];
	b->rows = new uint64_t[total ? total : 1];
	b->strings = new const void*[total ? total : 1];
	const uint8_t* chars = ix->hk.chars;
	if (total && ix->hk.off.empty()) {       // out-of-band path only: fetch the key offsets / rows once
		ix->hk.off.resize(ix->dix.K + 1);
		ix->hk.row.resize(ix->dix.K);
		SSB_CUDA(cudaMemcpyAsync(ix->hk.off.data(), ix->dix.keyOff.get(), (size_t)(ix->dix.K + 1) * 4, cudaMemcpyDeviceToHost, ix->stream));
		SSB_CUDA(cudaMemcpyAsync(ix->hk.row.data(), ix->dix.keyRow.get(), (size_t)ix->dix.K * 8, cudaMemcpyDeviceToHost, ix->stream));
		SSB_CUDA(cudaStreamSynchronize(ix->stream));
	}
	const int unit = ix->dix.unit;
	for (uint32_t q = 0; q < out.nq; ++q) {
		uint32_t n = out.count[q], o = b->offsets[q];
		if (n <= out.stride) {
			memcpy(b->kids + o, out.kid.data() + (size_t)q * out.stride, (size_t)n * 4);
			memcpy(b->scores + o, out.score.data() + (size_t)q * out.stride, (size_t)n * 4);
		}
	}
	for (auto& big : out.big) {
		uint32_t o = b->offsets[big.first];
		for (size_t i = 0; i < big.second.size(); ++i) { b->kids[o + i] = big.second[i].first; b->scores[o + i] = big.second[i].second; }
	}
	for (uint64_t i = 0; i < total; ++i) {
		uint32_t kid = b->kids[i];
		b->rows[i] = ix->hk.row[kid];
		b->strings[i] = chars + (size_t)ix->hk.off[kid] * unit;
	}
	return b;
}

// one batch: pinned slab as the D2H destination; falls back to heap marshalling when a result list exceeds the stride
ssb_batch* run_batch(ssb_index* ix, const void* units, const uint32_t* offsets, uint32_t nq, float threshold, uint32_t limit)
{
	const uint32_t lim = limit == 0 ? 0x7FFFFFFFu : limit;
	const size_t stride = std::min<uint32_t>(lim, FAST_KLIM);
	Slab* slab = ix->pool->get(nq, (size_t)nq * stride);
	slab->hr.strBase = reinterpret_cast<uint64_t>(ix->hk.chars);
	BatchOutput out;
	try {
		search_batch(ix->dix, ix->ws, units, offsets, nq, threshold, limit, out, ix->stream, &ix->lastStats, &slab->hr);
	} catch (...) {
		ix->pool->put(slab);
		throw;
	}
	if (!out.csr) {
		ix->pool->put(slab);
		if (!ix->dix.indexed || nq == 0 || ix->dix.K == 0) {
			out.nq = nq;
			out.count.assign(nq, 0);
		}
		return marshal(ix, out);
	}
	auto* impl = new BatchImpl();
	impl->pool = ix->pool;
	impl->slab = slab;
	impl->pub.nq = nq;
	impl->pub.offsets = slab->hr.offsets;
	impl->pub.kids = slab->hr.kids;
	impl->pub.scores = slab->hr.scores;
	impl->pub.rows = slab->hr.rows;
	impl->pub.strings = reinterpret_cast<const void**>(slab->hr.strs);
	return &impl->pub;
}

template <class F> auto guarded(F f, decltype(f()) fail) -> decltype(f())
{
	try {
		g_lastError[0] = 0;
		return f();
	} catch (const std::exception& e) {
		set_last_error(e.what());
		fprintf(stderr, "[ssb] %s\n", e.what());
	} catch (...) {
		set_last_error("unknown error");
	}
	return fail;
}

}   // namespace

// ---- serialisation (no reference equivalent: the reference rebuilds its hash containers from words[] every time).
// File = header (sizes, flags, validChar) + every device array as (u64 element count, raw bytes) + the host key table.
namespace {
struct IndexFileHeader {
	char magic[8];
	uint32_t version, headerBytes;
	int32_t g, unit, symBits;
	uint32_t indexed, K, T, E, G, P, nShort, maxTermLen, maxKeyLen, termUnits, keyUnits, A;
	uint32_t identity, uniformW, multiEdgeKeys, packed;
	float wmax;
	uint64_t rowBase, hostKeyBytes;
	uint8_t validLut[256];
};
constexpr char INDEX_MAGIC[8] = {'S', 'S', 'B', 'I', 'D', 'X', '1', 0};

struct FileCloser { void operator()(FILE* f) const { if (f) fclose(f); } };
using FilePtr = std::unique_ptr<FILE, FileCloser>;

void write_all(FILE* f, const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) throw std::runtime_error("ssb_index_save: write failed"); }
void read_all(FILE* f, void* p, size_t n) { if (n && fread(p, 1, n, f) != n) throw std::runtime_error("ssb_index_load: truncated file"); }

// device array <-> file through a pinned bounce buffer
template <class T> void save_buf(FILE* f, const DevBuf<T>& b, size_t count, PinBuf<uint8_t>& bounce, cudaStream_t s)
{
	const uint64_t n = b.get() ? count : 0;
	write_all(f, &n, 8);
	const size_t chunk = bounce.size();
	for (size_t off = 0; off < n * sizeof(T); off += chunk) {
		const size_t len = std::min(chunk, n * sizeof(T) - off);
		SSB_CUDA(cudaMemcpyAsync(bounce.get(), reinterpret_cast<const uint8_t*>(b.get()) + off, len, cudaMemcpyDeviceToHost, s));
		SSB_CUDA(cudaStreamSynchronize(s));
		write_all(f, bounce.get(), len);
	}
}
// expect: the element count the header implies, or SIZE_MAX for arrays whose length only the file knows (then the bytes left
// in the file bound it): a truncated or corrupt file must not turn into a huge allocation or into kernels that index past
// the end of a short array
template <class T> void load_buf(FILE* f, DevBuf<T>& b, size_t expect, uint64_t fileBytes, PinBuf<uint8_t>& bounce, cudaStream_t s)
{
	uint64_t n = 0;
	read_all(f, &n, 8);
	const long pos = ftell(f);
	if (pos < 0 || n > (1ull << 36) || n * sizeof(T) > fileBytes - std::min<uint64_t>(fileBytes, (uint64_t)pos) ||
	    (expect != SIZE_MAX && n != expect))
		throw std::runtime_error("ssb_index_load: corrupt or truncated index file (array size)");
	b.alloc(n, s);
	const size_t chunk = bounce.size();
	for (size_t off = 0; off < n * sizeof(T); off += chunk) {
		const size_t len = std::min(chunk, n * sizeof(T) - off);
		read_all(f, bounce.get(), len);
		SSB_CUDA(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(b.get()) + off, bounce.get(), len, cudaMemcpyHostToDevice, s));
		SSB_CUDA(cudaStreamSynchronize(s));
	}
}
// one visitor keeps the order of the arrays identical for save and load; the third argument tells whether the count follows
// from the header (true) or is the array's own size (false: unknown before the array has been read)
template <class F> void for_each_array(DeviceIndex& d, F f)
{
	f(d.symTable, d.symTable.size(), false); f(d.validLut, d.validLut.size(), false);
	f(d.gramCodes, (size_t)d.G, true); f(d.gramOff, (size_t)d.G + 1, true); f(d.postings, d.postings.size(), false);
	f(d.termOff, (size_t)d.T + 1, true); f(d.termChars, d.termChars.size(), false); f(d.termEdgeOff, (size_t)d.T + 1, true);
	f(d.edgeKid, (size_t)d.E, true); f(d.edgeW, (size_t)d.E, true); f(d.shortTerms, (size_t)d.nShort, true);
	f(d.keyOff, (size_t)d.K + 1, true); f(d.keyChars, d.keyChars.size(), false); f(d.keyEdgeOff, (size_t)d.K + 1, true);
	f(d.keyEdgeTerm, (size_t)d.E, true); f(d.keyEdgeW, (size_t)d.E, true); f(d.keyWmax, (size_t)d.K, true); f(d.keyRow, (size_t)d.K, true);
	f(d.keyHash, (size_t)d.K, true);
}
}   // namespace

extern "C" {

#define SSB_API __attribute__((visibility("default")))

SSB_API const char* ssb_last_error(void) { return g_lastError; }
SSB_API uint64_t ssb_kernel_launches(void) { return g_kernel_launches.load(); }
SSB_API int ssb_set_search_path(int path)
{
	if (path < -1 || path > 1) return -1;
	g_search_path.store(path);
	return 0;
}
SSB_API int ssb_set_device(int device)
{
	return guarded([&] { g_engine.device = device; g_engine.init(); return 0; }, -1);
}

SSB_API ssb_index* ssb_index_create(const void* const* words, uint64_t size, uint16_t rowSize, const float* weight, uint16_t gSize,
                                    int wide, uint64_t rowBase)
{
	return guarded([&]() -> ssb_index* {
		if (wide) return create_from_pointers<wchar_t>(reinterpret_cast<const wchar_t* const*>(words), size, rowSize, weight, gSize, rowBase);
		return create_from_pointers<char>(reinterpret_cast<const char* const*>(words), size, rowSize, weight, gSize, rowBase);
	}, nullptr);
}
SSB_API ssb_index* ssb_index_create_packed(const void* units, const uint64_t* offsets, const uint8_t* isNull, uint64_t size,
                                           uint16_t rowSize, const float* weight, uint16_t gSize, int wide, uint64_t rowBase)
{
	return guarded([&]() -> ssb_index* { return create_from_packed(units, offsets, isNull, size, rowSize, weight, gSize, wide ? 4 : 1, rowBase); },
	               nullptr);
}
SSB_API int ssb_index_save(ssb_index* ix, const char* path)
{
	if (!ix || !path) return -1;
	return guarded([&] {
		g_engine.init();
		std::lock_guard<std::mutex> lk(ix->mu);
		DeviceIndex& d = ix->dix;
		FilePtr f(fopen(path, "wb"));
		if (!f) throw std::runtime_error(std::string("ssb_index_save: cannot open ") + path);
		IndexFileHeader h{};
		memcpy(h.magic, INDEX_MAGIC, 8);
		h.version = 1; h.headerBytes = sizeof(IndexFileHeader);
		h.g = d.g; h.unit = d.unit; h.symBits = d.symBits;
		h.indexed = d.indexed; h.K = d.K; h.T = d.T; h.E = d.E; h.G = d.G; h.P = d.P; h.nShort = d.nShort;
		h.maxTermLen = d.maxTermLen; h.maxKeyLen = d.maxKeyLen; h.termUnits = d.termUnits; h.keyUnits = d.keyUnits; h.A = d.A;
		h.identity = d.identity; h.uniformW = d.uniformW; h.multiEdgeKeys = d.multiEdgeKeys; h.packed = d.packed;
		h.wmax = d.wmax; h.rowBase = d.rowBase; h.hostKeyBytes = ix->hk.charBytes;
		memcpy(h.validLut, ix->validLut, 256);
		write_all(f.get(), &h, sizeof(h));
		PinBuf<uint8_t> bounce;
		bounce.alloc_exact(16u << 20);
		for_each_array(d, [&](auto& b, size_t count, bool) { save_buf(f.get(), b, count, bounce, ix->stream); });
		write_all(f.get(), ix->hk.chars, ix->hk.charBytes);
		if (fflush(f.get()) != 0) throw std::runtime_error("ssb_index_save: flush failed");
		return 0;
	}, -1);
}

SSB_API ssb_index* ssb_index_load(const char* path)
{
	if (!path) return nullptr;
	return guarded([&]() -> ssb_index* {
		g_engine.init();
		FilePtr f(fopen(path, "rb"));
		if (!f) throw std::runtime_error(std::string("ssb_index_load: cannot open ") + path);
		IndexFileHeader h{};
		read_all(f.get(), &h, sizeof(h));
		if (memcmp(h.magic, INDEX_MAGIC, 8) != 0 || h.version != 1 || h.headerBytes != sizeof(IndexFileHeader))
			throw std::runtime_error("ssb_index_load: not an index file of this engine version");
		if ((h.unit != 1 && h.unit != 4) || h.g < 1 || h.g > MAX_GSIZE) throw std::runtime_error("ssb_index_load: corrupt header");
		if (fseek(f.get(), 0, SEEK_END) != 0) throw std::runtime_error("ssb_index_load: cannot seek");
		const long endPos = ftell(f.get());
		if (endPos < 0 || fseek(f.get(), (long)sizeof(h), SEEK_SET) != 0) throw std::runtime_error("ssb_index_load: cannot seek");
		const uint64_t fileBytes = (uint64_t)endPos;
		// (ssb_index_destroy, not delete: the stream has to be synchronised and destroyed too)
		std::unique_ptr<ssb_index, void (*)(ssb_index*)> ix(new ssb_index(), &ssb_index_destroy);
		SSB_CUDA(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
		DeviceIndex& d = ix->dix;
		d.g = h.g; d.unit = h.unit; d.symBits = h.symBits; d.indexed = h.indexed != 0;
		d.K = h.K; d.T = h.T; d.E = h.E; d.G = h.G; d.P = h.P; d.nShort = h.nShort;
		d.maxTermLen = h.maxTermLen; d.maxKeyLen = h.maxKeyLen; d.termUnits = h.termUnits; d.keyUnits = h.keyUnits; d.A = h.A;
		d.identity = h.identity != 0; d.uniformW = h.uniformW != 0; d.multiEdgeKeys = h.multiEdgeKeys != 0; d.packed = h.packed != 0;
		d.wmax = h.wmax; d.rowBase = h.rowBase;
		memcpy(ix->validLut, h.validLut, 256);
		PinBuf<uint8_t> bounce;
		bounce.alloc_exact(16u << 20);
		for_each_array(d, [&](auto& b, size_t count, bool fromHeader) { load_buf(f.get(), b, fromHeader ? count : SIZE_MAX, fileBytes, bounce, ix->stream); });
		if (d.postings.size() < (size_t)d.P || d.termChars.size() < (size_t)d.termUnits * d.unit || d.keyChars.size() < (size_t)d.keyUnits * d.unit ||
		    d.validLut.size() < 256 || h.hostKeyBytes > (uint64_t)d.keyChars.size())
			throw std::runtime_error("ssb_index_load: corrupt or truncated index file (table sizes)");
		ix->hk.reset();
		ix->hk.K = h.K;
		ix->hk.charBytes = h.hostKeyBytes;
		if (h.hostKeyBytes) {
			ix->hk.chars = static_cast<uint8_t*>(malloc(h.hostKeyBytes));
			if (!ix->hk.chars) throw std::bad_alloc();
			read_all(f.get(), ix->hk.chars, h.hostKeyBytes);
		}
		build_sieve(d, ix->stream);                               // derived data: not in the file
		SSB_CUDA(cudaStreamSynchronize(ix->stream));
		return ix.release();
	}, nullptr);
}

SSB_API void ssb_index_destroy(ssb_index* ix)
{
	if (!ix) return;
	guarded([&] {
		cudaStream_t s = ix->stream;
		delete ix;
		if (s && !g_exiting.load()) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
		return 0;
	}, 0);
}
SSB_API uint64_t ssb_index_size(const ssb_index* ix) { return ix && ix->dix.indexed ? ix->dix.T : 0; }
SSB_API uint64_t ssb_index_lib_size(const ssb_index* ix) { return ix && ix->dix.indexed ? ix->dix.G : 0; }
SSB_API void ssb_index_set_valid_char(ssb_index* ix, const char* chars, int n)
{
	if (!ix) return;
	guarded([&] {
		std::lock_guard<std::mutex> lk(ix->mu);
		memset(ix->validLut, 0, 256);
		for (int i = 0; i < n; ++i) ix->validLut[(unsigned char)chars[i]] = 1;
		if (ix->dix.validLut.get()) {
			SSB_CUDA(cudaMemcpyAsync(ix->dix.validLut.get(), ix->validLut, 256, cudaMemcpyHostToDevice, ix->stream));
			SSB_CUDA(cudaStreamSynchronize(ix->stream));
		}
		return 0;
	}, 0);
}
SSB_API void ssb_index_info(const ssb_index* ix, uint64_t info[24])
{
	memset(info, 0, 24 * sizeof(uint64_t));
	if (!ix) return;
	const DeviceIndex& d = ix->dix;
	info[0] = d.K; info[1] = d.T; info[2] = d.E; info[3] = d.G; info[4] = d.P; info[5] = d.nShort; info[6] = d.symBits; info[7] = d.A;
	info[8] = d.identity; info[9] = d.uniformW; info[10] = d.g; info[11] = d.unit; info[12] = d.indexed;
	info[13] = d.termUnits; info[14] = d.keyUnits;
	info[15] = d.device_bytes(); info[16] = ix->bst.windows; info[17] = d.packed; info[18] = d.multiEdgeKeys;
	info[19] = ix->bst.hashRetries;
}
SSB_API void ssb_index_build_ms(const ssb_index* ix, double ms[8])
{
	memset(ms, 0, 8 * sizeof(double));
	if (!ix) return;
	ms[0] = ix->bst.msTotal + ix->msPack; ms[1] = ix->msPack; ms[2] = ix->bst.msH2D; ms[3] = ix->bst.msIntern; ms[4] = ix->bst.msEdges;
	ms[5] = ix->bst.msGrams; ms[6] = ix->bst.msD2H; ms[7] = ix->bst.msSieve;
}
SSB_API int ssb_index_copy(const ssb_index* ix, int which, void* dst, uint64_t bytes)
{
	if (!ix) return -1;
	return guarded([&] {
		const DeviceIndex& d = ix->dix;
		const void* src = nullptr;
		uint64_t have = 0;
		auto pick = [&](const void* p, uint64_t b) { src = p; have = b; };
		switch (which) {
		case 0: pick(d.gramCodes.get(), (uint64_t)d.G * 8); break;
		case 1: pick(d.gramOff.get(), (uint64_t)(d.G + 1) * 4); break;
		case 2: pick(d.postings.get(), (uint64_t)d.P * 4); break;
		case 3: pick(d.termOff.get(), (uint64_t)(d.T + 1) * 4); break;
		case 4: pick(d.termChars.get(), (uint64_t)d.termUnits * d.unit); break;
		case 5: pick(d.termEdgeOff.get(), (uint64_t)(d.T + 1) * 4); break;
		case 6: pick(d.edgeKid.get(), (uint64_t)d.E * 4); break;
		case 7: pick(d.edgeW.get(), (uint64_t)d.E * 4); break;
		case 8: pick(d.keyOff.get(), (uint64_t)(d.K + 1) * 4); break;
		case 9: pick(d.keyChars.get(), (uint64_t)d.keyUnits * d.unit); break;
		case 10: pick(d.keyRow.get(), (uint64_t)d.K * 8); break;
		case 11: pick(d.symTable.get(), d.unit == 1 ? 1024 : (uint64_t)d.A * 4); break;
		case 12: pick(d.shortTerms.get(), (uint64_t)d.nShort * 4); break;
		default: throw std::runtime_error("ssb_index_copy: unknown table");
		}
		if (have != bytes) throw std::runtime_error("ssb_index_copy: size mismatch");
		if (bytes) {
			SSB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ix->stream));
			SSB_CUDA(cudaStreamSynchronize(ix->stream));
		}
		return 0;
	}, -1);
}

SSB_API ssb_batch* ssb_search_batch_packed(ssb_index* ix, const void* units, const uint32_t* offsets, uint32_t nq, float threshold,
                                           uint32_t limit)
{
	if (!ix) return nullptr;
	return guarded([&]() -> ssb_batch* {
		g_engine.init();
		std::lock_guard<std::mutex> lk(ix->mu);
		return run_batch(ix, units, offsets, nq, threshold, limit);
	}, nullptr);
}
// packs NUL-terminated query strings into the index's pinned staging; returns the offsets array
static const uint32_t* pack_queries(ssb_index* ix, const void* const* queries, uint32_t nq)
{
	const int unit = ix->dix.unit;
	ix->pinQOff.reserve(nq + 1);
	uint32_t* off = ix->pinQOff.get();
	uint64_t total = 0;
	for (uint32_t q = 0; q < nq; ++q) {
		off[q] = (uint32_t)total;
		const void* p = queries[q];
		total += !p ? 0 : (unit == 1 ? strlen((const char*)p) : wcslen((const wchar_t*)p));
	}
	off[nq] = (uint32_t)total;
	ix->pinQ.reserve(std::max<uint64_t>(total * unit, 16));
	for (uint32_t q = 0; q < nq; ++q)
		if (queries[q]) memcpy(ix->pinQ.get() + (size_t)off[q] * unit, queries[q], (size_t)(off[q + 1] - off[q]) * unit);
	return off;
}

SSB_API int ssb_search_batch_device(ssb_index* ix, const void* const* queries, uint32_t nq, float threshold, uint32_t limit)
{
	if (!ix) return -1;
	return guarded([&] {
		g_engine.init();
		std::lock_guard<std::mutex> lk(ix->mu);
		const uint32_t* off = pack_queries(ix, queries, nq);
		BatchOutput out;
		search_batch(ix->dix, ix->ws, ix->pinQ.get(), off, nq, threshold, limit, out, ix->stream, &ix->lastStats, nullptr, true);
		return 0;
	}, -1);
}

SSB_API ssb_batch* ssb_search_batch(ssb_index* ix, const void* const* queries, uint32_t nq, float threshold, uint32_t limit)
{
	if (!ix) return nullptr;
	return guarded([&]() -> ssb_batch* {
		g_engine.init();
		std::lock_guard<std::mutex> lk(ix->mu);
		const int unit = ix->dix.unit;
		ix->pinQOff.reserve(nq + 1);
		uint32_t* off = ix->pinQOff.get();
		uint64_t total = 0;
		for (uint32_t q = 0; q < nq; ++q) {
			off[q] = (uint32_t)total;
			const void* p = queries[q];
			total += !p ? 0 : (unit == 1 ? strlen((const char*)p) : wcslen((const wchar_t*)p));
		}
		off[nq] = (uint32_t)total;
		ix->pinQ.reserve(std::max<uint64_t>(total * unit, 16));
		for (uint32_t q = 0; q < nq; ++q)
			if (queries[q]) memcpy(ix->pinQ.get() + (size_t)off[q] * unit, queries[q], (size_t)(off[q + 1] - off[q]) * unit);
		return run_batch(ix, ix->pinQ.get(), off, nq, threshold, limit);
	}, nullptr);
}
// Single query with coalescing (the reference's intended usage is many threads calling search() under a
// shared lock, dllmain.cpp:63). The first caller becomes the leader and runs a batch; callers that arrive meanwhile queue
// up and are answered together by the next leader in ONE batch (same threshold / limit), so concurrent single calls get
// batch throughput without any added latency when the index is idle.
SSB_API ssb_batch* ssb_search_one(ssb_index* ix, const void* query, float threshold, uint32_t limit)
{
	if (!ix) return nullptr;
	ssb_index::OneReq me{query, threshold, limit};
	ix->oneCalls.fetch_add(1, std::memory_order_relaxed);
	std::unique_lock<std::mutex> lk(ix->oneMu);
	ix->onePending.push_back(&me);
	while (!me.done) {
		if (ix->oneBusy) { ix->oneCv.wait(lk); continue; }
		// leader: everything pending with my parameters goes into one batch — my own request always (thresholds are compared
		// bit for bit: a NaN threshold still matches itself)
		ix->oneBusy = true;
		std::vector<ssb_index::OneReq*> mine, rest;
		for (auto* r : ix->onePending)
			((r == &me || (memcmp(&r->th, &me.th, sizeof(float)) == 0 && r->limit == me.limit)) ? mine : rest).push_back(r);
		ix->onePending.swap(rest);
		lk.unlock();
		// whatever happens below, every request of this batch is marked done and the index is released: a caller left
		// waiting would wait forever (out stays nullptr on failure, which is this function's error value)
		try {
			std::vector<const void*> qs(mine.size());
			for (size_t i = 0; i < mine.size(); ++i) qs[i] = mine[i]->q;
			ssb_batch* b = ssb_search_batch(ix, qs.data(), (uint32_t)qs.size(), threshold, limit);
			ix->oneBatches.fetch_add(1, std::memory_order_relaxed);
			if (b && mine.size() == 1) mine[0]->out = b;
			else if (b) {
				for (size_t i = 0; i < mine.size(); ++i) {             // hand every caller its own slice (heap arrays, like marshal())
					const uint32_t o = b->offsets[i], n = b->offsets[i + 1] - o;
					auto* impl = new BatchImpl();
					ssb_batch* s = &impl->pub;
					mine[i]->out = s;                                  // owned from here on: ssb_batch_free copes with partly filled slices
					s->nq = 1;
					s->offsets = new uint32_t[2]{0, n};
					s->kids = new uint32_t[n ? n : 1]; s->scores = new float[n ? n : 1];
					s->rows = new uint64_t[n ? n : 1]; s->strings = new const void*[n ? n : 1];
					memcpy(s->kids, b->kids + o, (size_t)n * 4); memcpy(s->scores, b->scores + o, (size_t)n * 4);
					memcpy(s->rows, b->rows + o, (size_t)n * 8); memcpy(s->strings, b->strings + o, (size_t)n * sizeof(void*));
				}
				ssb_batch_free(b);
			}
		} catch (const std::exception& e) {
			set_last_error(e.what());
		} catch (...) {
			set_last_error("ssb_search_one: unknown error");
		}
		lk.lock();
		for (auto* r : mine) r->done = true;
		ix->oneBusy = false;
		ix->oneCv.notify_all();
	}
	return me.out;
}
/* calls / batches of ssb_search_one since the index was created (how well concurrent callers coalesce) */
SSB_API void ssb_search_one_stats(const ssb_index* ix, uint64_t out[2])
{
	out[0] = ix ? ix->oneCalls.load() : 0;
	out[1] = ix ? ix->oneBatches.load() : 0;
}
SSB_API void ssb_batch_free(ssb_batch* b)
{
	if (!b) return;
	auto* impl = reinterpret_cast<BatchImpl*>(b);
	if (impl->slab) impl->pool->put(impl->slab);
	else { delete[] b->offsets; delete[] b->kids; delete[] b->scores; delete[] b->rows; delete[] b->strings; }
	delete impl;
}
SSB_API int ssb_search_resident(ssb_index* ix, uint32_t nq, float threshold, uint32_t limit, float ms[2], uint64_t stats[8])
{
	if (!ix) return -1;
	return guarded([&] {
		g_engine.init();
		std::lock_guard<std::mutex> lk(ix->mu);
		SearchStats st;
		search_batch_resident(ix->dix, ix->ws, nq, threshold, limit, ix->stream, &st);
		if (ms) { ms[0] = st.msDevice; ms[1] = st.msFastKernel; }
		if (stats) {
			stats[0] = st.nFast; stats[1] = st.nDense; stats[2] = st.nOverflow; stats[3] = st.postingsTouched;
			stats[4] = st.nSieve; stats[5] = st.nSieveHandover; stats[6] = st.nSieveGiveUp; stats[7] = st.sieveProbes;
		}
		return 0;
	}, -1);
}
SSB_API int ssb_export_records(ssb_index* ix, uint32_t nq, uint32_t limit, void* d_out)
{
	if (!ix) return -1;
	return guarded([&] {
		g_engine.init();
		std::lock_guard<std::mutex> lk(ix->mu);
		// the batch's results are complete (search calls synchronise their stream); everything of the exchange is only
		// ENQUEUED on the legacy default stream, in order with the caller's collectives
		export_records(ix->dix, ix->ws, nq, limit, static_cast<MergeRecord*>(d_out), 0);
		return 0;
	}, -1);
}
SSB_API int ssb_merge_records(const void* d_gathered, uint32_t nShards, uint32_t nq, uint32_t limit, void* d_merged, uint32_t* d_counts)
{
	return guarded([&] {
		g_engine.init();
		// asynchronous on the legacy default stream: ordered behind collectives the caller enqueued there (torch's default)
		merge_records(static_cast<const MergeRecord*>(d_gathered), nShards, nq, limit, static_cast<MergeRecord*>(d_merged), d_counts, 0);
		return 0;
	}, -1);
}

SSB_API uint64_t ssb_index_digest_count(const ssb_index* ix, int which)
{
	return ix && (which == 0 || which == 1) ? digest_count(ix->dix, which) : 0;
}
SSB_API int ssb_index_export_digests(ssb_index* ix, int which, void* d_out)
{
	if (!ix || (which != 0 && which != 1)) return -1;
	return guarded([&] {
		g_engine.init();
		std::lock_guard<std::mutex> lk(ix->mu);
		export_digests(ix->dix, which, d_out, ix->stream);
		return 0;
	}, -1);
}
SSB_API int ssb_count_distinct_digests(const void* d_digests, uint64_t n, uint32_t part, uint32_t nParts, uint64_t* out)
{
	if (!out) return -1;
	return guarded([&] {
		g_engine.init();
		*out = count_distinct_digests(d_digests, n, part, nParts, 0);
		return 0;
	}, -1);
}

SSB_API void* ssb_device_alloc(uint64_t bytes)
{
	return guarded([&]() -> void* {
		g_engine.init();
		void* p = nullptr;
		SSB_CUDA(cudaMalloc(&p, bytes));
		SSB_CUDA(cudaMemset(p, 0xFF, bytes));
		return p;
	}, nullptr);
}
SSB_API void ssb_device_free(void* d_ptr) { if (d_ptr) cudaFree(d_ptr); }
SSB_API int ssb_memcpy_d2h(void* dst, const void* d_src, uint64_t bytes)
{
	return guarded([&] { g_engine.init(); SSB_CUDA(cudaMemcpy(dst, d_src, bytes, cudaMemcpyDeviceToHost)); return 0; }, -1);
}
SSB_API int ssb_ipc_get_handle(void* d_ptr, unsigned char handle[64])
{
	return guarded([&] {
		g_engine.init();
		static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
		cudaIpcMemHandle_t h;
		SSB_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
		memcpy(handle, &h, 64);
		return 0;
	}, -1);
}
SSB_API void* ssb_ipc_open_handle(const unsigned char handle[64])
{
	return guarded([&]() -> void* {
		g_engine.init();
		cudaIpcMemHandle_t h;
		memcpy(&h, handle, 64);
		void* p = nullptr;
		SSB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
		return p;
	}, nullptr);
}
SSB_API int ssb_ipc_close_handle(void* d_ptr) { return d_ptr && cudaIpcCloseMemHandle(d_ptr) == cudaSuccess ? 0 : -1; }
SSB_API int ssb_export_records_p2p(ssb_index* ix, uint32_t nq, uint32_t limit, void* const* peerRecv, uint32_t nRanks, uint32_t myRank)
{
	if (!ix) return -1;
	return guarded([&] {
		g_engine.init();
		std::lock_guard<std::mutex> lk(ix->mu);
		export_records_p2p(ix->dix, ix->ws, nq, limit, reinterpret_cast<MergeRecord* const*>(peerRecv), nRanks, myRank, 0);
		return 0;
	}, -1);
}
SSB_API int ssb_export_records_p2p_signal(ssb_index* ix, uint32_t nq, uint32_t limit, void* const* peerRecv, uint32_t nRanks, uint32_t myRank,
                                          void* const* peerFlags, uint32_t epoch)
{
	if (!ix || !peerFlags) return -1;
	return guarded([&] {
		g_engine.init();
		std::lock_guard<std::mutex> lk(ix->mu);
		export_records_p2p(ix->dix, ix->ws, nq, limit, reinterpret_cast<MergeRecord* const*>(peerRecv), nRanks, myRank, 0,
		                   reinterpret_cast<uint32_t* const*>(peerFlags), epoch);
		return 0;
	}, -1);
}
SSB_API int ssb_merge_records_p2p_wait(const void* d_recv, uint32_t nShards, uint32_t nqLocal, uint32_t limit, void* const* peerMerged,
                                       void* const* peerCounts, uint32_t nOut, uint32_t myRank, const void* d_myFlags, uint32_t epoch)
{
	if (!d_myFlags) return -1;
	return guarded([&] {
		g_engine.init();
		merge_records_p2p(static_cast<const MergeRecord*>(d_recv), nShards, nqLocal, limit, reinterpret_cast<MergeRecord* const*>(peerMerged),
		                  reinterpret_cast<uint32_t* const*>(peerCounts), nOut, myRank, 0, static_cast<const uint32_t*>(d_myFlags), epoch);
		return 0;
	}, -1);
}
SSB_API int ssb_device_memset(void* d_ptr, int value, uint64_t bytes)
{
	return guarded([&] { g_engine.init(); SSB_CUDA(cudaMemset(d_ptr, value, bytes)); return 0; }, -1);
}
SSB_API int ssb_merge_records_p2p(const void* d_recv, uint32_t nShards, uint32_t nqLocal, uint32_t limit, void* const* peerMerged,
                                  void* const* peerCounts, uint32_t nRanks, uint32_t myRank)
{
	return guarded([&] {
		g_engine.init();
		merge_records_p2p(static_cast<const MergeRecord*>(d_recv), nShards, nqLocal, limit, reinterpret_cast<MergeRecord* const*>(peerMerged),
		                  reinterpret_cast<uint32_t* const*>(peerCounts), nRanks, myRank, 0);
		return 0;
	}, -1);
}

}   // extern "C"
