// digest.cu — global getSize / getLibSize over key-range shards (SURVEY.md §8e "Global counters").
//
// size() and libSize() (nGramSearch.hpp:488-499) are DISTINCT counts — terms and grams — and the same term or gram can occur
// in several shards, so the global numbers are the size of a union, not a sum. Every shard exports one 16-byte digest per
// term / per gram; the digests of all shards are gathered (one all-gather, or peer copies), every rank counts the distinct
// digests of its own residue class (lo % nParts == part: equal digests land on the same rank) and the per-rank counts are
// summed (one 8-byte all-reduce). Nothing shard-local leaks into a digest: term ids and packed gram codes depend on the
// shard's own alphabet remap (index.h), so digests are computed from the code units themselves.
//   gram digest : the g code units packed exactly when they fit 128 bits (narrow: g <= 16, wide: g <= 4) — the count is then
//                 exact like the reference's injective gramHash (nGramSearch.h:147-150); two independent 64-bit hashes otherwise
//   term digest : two independent 64-bit hashes of (length, units) — a collision needs 128 equal bits (p < T^2 / 2^129:
//                 1.5e-23 at 10^8 terms); the multi-GPU merge already identifies equal key strings by a 64-bit hash (merge.cu)
#include "digest.h"
#include "primitives.cuh"
#include "textops.cuh"

namespace ssb {

namespace {

constexpr int TPB = 256;
inline unsigned grid_for(size_t n) { return (unsigned)((n + TPB - 1) / TPB ? (n + TPB - 1) / TPB : 1); }

struct Digest { uint64_t lo, hi; };

__device__ __forceinline__ uint64_t chain2_step(uint64_t h, uint32_t u) { return mix64(h + 0x9E3779B97F4A7C15ull + u); }

template <class U>
__global__ void k_term_digests(const uint32_t* __restrict__ termOff, const U* __restrict__ termChars, uint32_t T, Digest* __restrict__ out)
{
	const uint32_t t = blockIdx.x * TPB + threadIdx.x;
	if (t >= T) return;
	const uint32_t b = termOff[t], e = termOff[t + 1];
	uint64_t h1 = 0x7465726d73ull, h2 = 0x5EED5EED5EEDull;
	for (uint32_t i = b; i < e; ++i) { const uint32_t u = (uint32_t)termChars[i]; h1 = hash_step(h1, u); h2 = chain2_step(h2, u); }
	out[t].lo = hash_done(h1, e - b);
	out[t].hi = mix64(h2 ^ (uint64_t)(e - b));
}

// the units of gram i are read from a term that contains it: the first posting's term, the window whose packed code equals
// gramCodes[i] (the same packing as k_emit_grams, build.cu)
template <class U>
__global__ void k_gram_digests(const uint64_t* __restrict__ gramCodes, const uint32_t* __restrict__ gramOff, const uint32_t* __restrict__ postings,
                               const uint32_t* __restrict__ termOff, const U* __restrict__ termChars, const uint32_t* __restrict__ symTable,
                               uint32_t A, uint32_t symBits, uint32_t g, uint32_t G, Digest* __restrict__ out, uint32_t* __restrict__ bad)
{
	const uint32_t i = blockIdx.x * TPB + threadIdx.x;
	if (i >= G) return;
	const uint64_t want = gramCodes[i];
	const uint32_t t = postings[gramOff[i]];
	const uint32_t b = termOff[t], e = termOff[t + 1];
	const uint64_t mask = (g * symBits >= 64) ? ~0ull : ((1ull << (g * symBits)) - 1);
	uint64_t code = 0;
	uint32_t at = 0xFFFFFFFFu;
	for (uint32_t j = b; j < e; ++j) {
		code = ((code << symBits) | sym_of(termChars[j], symTable, A)) & mask;
		if (j + 1 >= b + g && code == want) { at = j + 1 - g; break; }
	}
	if (at == 0xFFFFFFFFu) { atomicAdd(bad, 1u); out[i].lo = ~0ull; out[i].hi = i; return; }
	constexpr uint32_t unitBits = sizeof(U) == 1 ? 8u : 32u;
	Digest d{0, 0};
	if (g * unitBits <= 128) {
		for (uint32_t j = 0; j < g; ++j) {
			const uint64_t u = (uint64_t)(uint32_t)termChars[at + j];
			const uint32_t sh = j * unitBits;
			if (sh < 64) d.lo |= u << sh; else d.hi |= u << (sh - 64);
		}
	} else {
		uint64_t h1 = 0x6772616d73ull, h2 = 0xC0FFEEC0FFEEull;
		for (uint32_t j = 0; j < g; ++j) { const uint32_t u = (uint32_t)termChars[at + j]; h1 = hash_step(h1, u); h2 = chain2_step(h2, u); }
		d.lo = hash_done(h1, g);
		d.hi = mix64(h2 ^ g);
	}
	out[i] = d;
}

// digests of one residue class, appended in any order (they are sorted next); the residue is taken of a mix of both halves so
// that exactly packed grams spread as evenly as hashes do
__global__ void k_select_part(const Digest* __restrict__ in, uint64_t n, uint32_t part, uint32_t nParts, uint64_t* __restrict__ lo,
                              uint32_t* __restrict__ src, uint32_t* __restrict__ count)
{
	const uint64_t i = (uint64_t)blockIdx.x * TPB + threadIdx.x;
	if (i >= n) return;
	const Digest d = in[i];
	if (mix64(d.lo ^ mix64(d.hi)) % nParts != part) return;
	const uint32_t at = atomicAdd(count, 1u);
	lo[at] = d.lo;
	src[at] = (uint32_t)i;
}

// after the sort by lo: one thread per run of equal lo counts the distinct hi values inside it (runs are the copies of one
// digest in the shards — at most nShards long — unless two different digests share 64 bits)
__global__ void k_count_runs(const uint64_t* __restrict__ lo, const uint32_t* __restrict__ src, const Digest* __restrict__ in, uint32_t n,
                             unsigned long long* __restrict__ distinct)
{
	const uint32_t i = blockIdx.x * TPB + threadIdx.x;
	uint32_t mine = 0;
	if (i < n && (i == 0 || lo[i - 1] != lo[i])) {
		const uint64_t l = lo[i];
		for (uint32_t a = i; a < n && lo[a] == l; ++a) {
			const uint64_t h = in[src[a]].hi;
			bool seen = false;
			for (uint32_t b = i; b < a && !seen; ++b) seen = in[src[b]].hi == h;
			mine += seen ? 0u : 1u;
		}
	}
	// one atomic per warp
	for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
	if ((threadIdx.x & 31) == 0 && mine) atomicAdd(distinct, (unsigned long long)mine);
}

}   // namespace

uint64_t digest_count(const DeviceIndex& ix, int which) { return !ix.indexed ? 0 : (which == 0 ? ix.T : ix.G); }

void export_digests(const DeviceIndex& ix, int which, void* d_out, cudaStream_t s)
{
	if (!ix.indexed) return;
	Digest* out = static_cast<Digest*>(d_out);
	if (which == 0) {
		if (!ix.T) return;
		if (ix.unit == 1) k_term_digests<uint8_t><<<grid_for(ix.T), TPB, 0, s>>>(ix.termOff.get(), ix.termChars.get(), ix.T, out);
		else k_term_digests<uint32_t><<<grid_for(ix.T), TPB, 0, s>>>(ix.termOff.get(), reinterpret_cast<const uint32_t*>(ix.termChars.get()), ix.T, out);
		SSB_LAUNCH_CHECK();
		SSB_CUDA(cudaStreamSynchronize(s));
		return;
	}
	if (!ix.G) return;
	DevBuf<uint32_t> bad;
	bad.alloc(1, s); bad.zero(s);
	if (ix.unit == 1)
		k_gram_digests<uint8_t><<<grid_for(ix.G), TPB, 0, s>>>(ix.gramCodes.get(), ix.gramOff.get(), ix.postings.get(), ix.termOff.get(), ix.termChars.get(),
		                                                       ix.symTable.get(), ix.A, (uint32_t)ix.symBits, (uint32_t)ix.g, ix.G, out, bad.get());
	else
		k_gram_digests<uint32_t><<<grid_for(ix.G), TPB, 0, s>>>(ix.gramCodes.get(), ix.gramOff.get(), ix.postings.get(), ix.termOff.get(),
		                                                        reinterpret_cast<const uint32_t*>(ix.termChars.get()), ix.symTable.get(), ix.A,
		                                                        (uint32_t)ix.symBits, (uint32_t)ix.g, ix.G, out, bad.get());
	SSB_LAUNCH_CHECK();
	uint32_t hBad = 0;
	SSB_CUDA(cudaMemcpyAsync(&hBad, bad.get(), 4, cudaMemcpyDeviceToHost, s));
	SSB_CUDA(cudaStreamSynchronize(s));
	bad.release();
	if (hBad) throw std::runtime_error("export_digests: a gram was not found in the term of its first posting");
}

uint64_t count_distinct_digests(const void* d_digests, uint64_t n, uint32_t part, uint32_t nParts, cudaStream_t s)
{
	if (!n) return 0;
	if (!nParts || part >= nParts) throw std::runtime_error("count_distinct_digests: part / nParts");
	if (n >= (1ull << 32) - 16) throw std::runtime_error("count_distinct_digests: more than 2^32 digests");
	const Digest* in = static_cast<const Digest*>(d_digests);
	DevBuf<uint64_t> lo, loAlt, acc;
	DevBuf<uint32_t> src, srcAlt, cnt;
	// (the residue class holds ~n / nParts digests; sized for the worst case of a single class)
	lo.alloc(n, s); loAlt.alloc(n, s); src.alloc(n, s); srcAlt.alloc(n, s); cnt.alloc(1, s); acc.alloc(1, s);
	cnt.zero(s); acc.zero(s);
	k_select_part<<<grid_for(n), TPB, 0, s>>>(in, n, part, nParts, lo.get(), src.get(), cnt.get());
	SSB_LAUNCH_CHECK();
	uint32_t m = 0;
	SSB_CUDA(cudaMemcpyAsync(&m, cnt.get(), 4, cudaMemcpyDeviceToHost, s));
	SSB_CUDA(cudaStreamSynchronize(s));
	uint64_t distinct = 0;
	if (m) {
		const bool alt = radix_sort_u64(lo.get(), loAlt.get(), src.get(), srcAlt.get(), m, 0, 64, s);
		k_count_runs<<<grid_for(m), TPB, 0, s>>>(alt ? loAlt.get() : lo.get(), alt ? srcAlt.get() : src.get(), in, m,
		                                         reinterpret_cast<unsigned long long*>(acc.get()));
		SSB_LAUNCH_CHECK();
		SSB_CUDA(cudaMemcpyAsync(&distinct, acc.get(), 8, cudaMemcpyDeviceToHost, s));
		SSB_CUDA(cudaStreamSynchronize(s));
	}
	lo.release(); loAlt.release(); src.release(); srcAlt.release(); cnt.release(); acc.release();
	return distinct;
}

}   // namespace ssb
