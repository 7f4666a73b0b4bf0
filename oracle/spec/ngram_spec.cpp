// oracle/spec/ngram_spec.cpp — C API over the CPU restatement (ngram_spec.hpp). TEST INFRASTRUCTURE ONLY:
// loaded through ctypes by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; never by the product.
#include "ngram_spec.hpp"
#include <thread>

namespace {
struct Handle {
	int wide;
	spec::Index<char>* n = nullptr;
	spec::Index<wchar_t>* w = nullptr;
};
template <class F> auto visit(void* h, F f)
{
	auto* H = static_cast<Handle*>(h);
	return H->wide ? f(*H->w) : f(*H->n);
}
template <class C> uint64_t chars_total(const std::vector<std::basic_string<C>>& v)
{
	uint64_t s = 0;
	for (auto& x : v) s += x.size();
	return s;
}
template <class C> void dump_strs(const std::vector<std::basic_string<C>>& v, uint64_t* off, void* chars)
{
	C* out = static_cast<C*>(chars);
	uint64_t p = 0;
	for (size_t i = 0; i < v.size(); ++i) {
		off[i] = p;
		memcpy(out + p, v[i].data(), v[i].size() * sizeof(C));
		p += v[i].size();
	}
	off[v.size()] = p;
}
}   // namespace

#define API extern "C" __attribute__((visibility("default")))

API void* spec_create(const void* const* words, uint64_t size, uint16_t rowSize, const float* weight, uint16_t gSize,
                      int wide, uint64_t rowBase)
{
	auto* H = new Handle{wide};
	if (wide) {
		H->w = new spec::Index<wchar_t>();
		H->w->build(reinterpret_cast<const wchar_t* const*>(words), size, rowSize, weight, gSize, rowBase);
	} else {
		H->n = new spec::Index<char>();
		H->n->build(reinterpret_cast<const char* const*>(words), size, rowSize, weight, gSize, rowBase);
	}
	return H;
}
API void spec_free(void* h)
{
	auto* H = static_cast<Handle*>(h);
	delete H->n; delete H->w; delete H;
}
API int spec_indexed(void* h) { return visit(h, [](auto& ix) { return (int)ix.indexed; }); }
API uint64_t spec_size(void* h) { return visit(h, [](auto& ix) { return ix.size(); }); }
API uint64_t spec_lib_size(void* h) { return visit(h, [](auto& ix) { return ix.lib_size(); }); }
API uint64_t spec_num_keys(void* h) { return visit(h, [](auto& ix) { return (uint64_t)ix.keyStr.size(); }); }
API uint64_t spec_num_edges(void* h) { return visit(h, [](auto& ix) { return (uint64_t)ix.edgeKid.size(); }); }
API uint64_t spec_num_postings(void* h) { return visit(h, [](auto& ix) { return (uint64_t)ix.postings.size(); }); }
API uint64_t spec_num_short(void* h) { return visit(h, [](auto& ix) { return (uint64_t)ix.shortTerms.size(); }); }
API uint64_t spec_longest(void* h) { return visit(h, [](auto& ix) { return (uint64_t)ix.longest; }); }
API uint64_t spec_key_chars_total(void* h) { return visit(h, [](auto& ix) { return chars_total(ix.keyStr); }); }
API uint64_t spec_term_chars_total(void* h) { return visit(h, [](auto& ix) { return chars_total(ix.termStr); }); }
// keys in ext-id order / terms in term-id order; chars are char or wchar_t units
API void spec_keys(void* h, uint64_t* off, void* chars, uint32_t* kidOfExt, uint64_t* firstRow)
{
	visit(h, [&](auto& ix) {
		dump_strs(ix.keyStr, off, chars);
		for (size_t i = 0; i < ix.keyStr.size(); ++i) { kidOfExt[i] = ix.kidOfExt[i]; firstRow[i] = ix.keyFirstRow[i]; }
		return 0;
	});
}
API void spec_terms(void* h, uint64_t* off, void* chars) { visit(h, [&](auto& ix) { dump_strs(ix.termStr, off, chars); return 0; }); }
API void spec_edges(void* h, uint32_t* off, uint32_t* kid, float* w)
{
	visit(h, [&](auto& ix) {
		std::copy(ix.termEdgeOff.begin(), ix.termEdgeOff.end(), off);
		std::copy(ix.edgeKid.begin(), ix.edgeKid.end(), kid);
		std::copy(ix.edgeW.begin(), ix.edgeW.end(), w);
		return 0;
	});
}
// grams: G*g code units (as u32), CSR offsets (G+1), postings (term ids)
API void spec_grams(void* h, uint32_t* units, uint64_t* off, uint32_t* postings)
{
	visit(h, [&](auto& ix) {
		using Ops = typename std::decay_t<decltype(ix)>::Ops;
		size_t p = 0;
		for (auto& gs : ix.gramStr) for (auto c : gs) units[p++] = Ops::unit(c);
		std::copy(ix.gramOff.begin(), ix.gramOff.end(), off);
		if (ix.gramOff.empty()) off[0] = 0;
		std::copy(ix.postings.begin(), ix.postings.end(), postings);
		return 0;
	});
}
// normalises the raw query exactly as search() does and returns (term id, count) pairs of the long path
API uint32_t spec_long_counts(void* h, const void* query, uint32_t* terms, uint32_t* counts, uint32_t cap)
{
	return visit(h, [&](auto& ix) {
		using Ix = std::decay_t<decltype(ix)>;
		typename Ix::Str q(reinterpret_cast<const typename Ix::Str::value_type*>(query));
		q = ix.normalise(q);
		std::vector<std::pair<uint32_t, uint32_t>> out;
		ix.long_counts(q, out);
		for (uint32_t i = 0; i < out.size() && i < cap; ++i) { terms[i] = out[i].first; counts[i] = out[i].second; }
		return (uint32_t)out.size();
	});
}
API uint32_t spec_search(void* h, const void* query, float threshold, uint32_t limit, uint32_t* outExt, float* outScore,
                         uint32_t cap)
{
	return visit(h, [&](auto& ix) {
		using Ix = std::decay_t<decltype(ix)>;
		auto r = ix.search(reinterpret_cast<const typename Ix::Str::value_type*>(query), threshold, limit);
		for (uint32_t i = 0; i < r.size() && i < cap; ++i) { outExt[i] = r[i].key; outScore[i] = r[i].score; }
		return (uint32_t)r.size();
	});
}
// batch of queries, results in fixed-stride rows; nthreads caller threads (queries striped)
API void spec_search_batch(void* h, const void* const* queries, uint32_t nq, float threshold, uint32_t limit,
                           uint32_t* counts, uint32_t* outExt, float* outScore, uint32_t stride, int nthreads)
{
	if (nthreads < 1) nthreads = 1;
	auto work = [&](int tid) {
		for (uint32_t i = tid; i < nq; i += nthreads)
			counts[i] = spec_search(h, queries[i], threshold, limit, outExt + (size_t)i * stride,
			                        outScore + (size_t)i * stride, stride);
	};
	std::vector<std::thread> th;
	for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
	work(0);
	for (auto& t : th) t.join();
}
API void spec_set_valid_char(void* h, const char* chars, int n)
{
	visit(h, [&](auto& ix) {
		memset(ix.valid, 0, sizeof ix.valid);
		for (int i = 0; i < n; ++i) ix.valid[(unsigned char)chars[i]] = 1;
		return 0;
	});
}
