// sieve.cu — block bitmaps over the posting lists (the "sieve" that k_sieve_search reads instead of the postings).
//
// The reference counts matches by touching every posting of every query gram (nGramSearch.hpp:289-298). Once the running
// k-th best or the threshold (hpp:313-316) demands c matches, a term can only matter if it lies in a BLOCK of 2^shift
// consecutive term ids in which at least c of the query's lists have a posting. For every gram with a long enough list
// this file derives from the (bit-exact) CSR postings one array of GROUPS, 96 blocks each:
//   { first, bits0, bits1, bits2 }   first = absolute index of the list's first posting at or after the group's first block,
//                                    bits  = one bit per block: the list has a posting in the block
//   { 128 bits of slots }            (SSB_SIEVE_INLINE, index.h) per occupied block, in order: offset of its first posting inside
//                                    the block + "more postings follow in this block"
// so that a query streams K such arrays through a bit-sliced threshold counter and resolves the few blocks that pass it from
// the same 32-byte sector: block number + slot offset is the term id. Only blocks in which a list has several postings (or
// groups with more occupied blocks than slots) go back to the postings: first + popcount of the lower bits is the position
// of the block's first posting. Derived data: rebuilt after ssb_index_load, never serialised, never compared by the parity tests —
// postings stay the source of truth.
#include "index.h"
#include "primitives.cuh"
#include <algorithm>
#include <cmath>

namespace ssb {
namespace {

constexpr int TPB = 256;
inline unsigned grid_for(size_t n, int tpb = TPB) { return (unsigned)std::max<size_t>(1, (n + tpb - 1) / tpb); }

// [0] = sum of list lengths, [1] = sum of squared list lengths (query grams are drawn roughly in proportion to their
// frequency, so the length a query sees is the P-weighted mean sum(L^2) / sum(L))
__global__ void __launch_bounds__(TPB) k_sieve_stats(const uint32_t* __restrict__ gramOff, uint32_t G, unsigned long long* __restrict__ acc)
{
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long l = 0, l2 = 0;
	if (g < G) { l = gramOff[g + 1] - gramOff[g]; l2 = l * l; }
	for (int o = 16; o; o >>= 1) { l += __shfl_xor_sync(0xffffffffu, l, o); l2 += __shfl_xor_sync(0xffffffffu, l2, o); }
	if ((threadIdx.x & 31) == 0 && l) { atomicAdd(&acc[0], l); atomicAdd(&acc[1], l2); }
}

__global__ void __launch_bounds__(TPB) k_sieve_flags(const uint32_t* __restrict__ gramOff, uint32_t G, uint32_t minLen, uint8_t* __restrict__ flag)
{
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g < G) flag[g] = (gramOff[g + 1] - gramOff[g]) >= minLen;
}

__global__ void __launch_bounds__(TPB) k_sieve_slots(const uint8_t* __restrict__ flag, const uint32_t* __restrict__ pos, uint32_t G,
                                                     uint32_t* __restrict__ slotOf, uint32_t* __restrict__ gramOfSlot)
{
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= G) return;
	if (flag[g]) { slotOf[g] = pos[g]; gramOfSlot[pos[g]] = g; }
	else slotOf[g] = 0xFFFFFFFFu;
}

// one CTA per slot: bits from the postings, the groups' first-posting entries by binary search
template <bool INL>
__global__ void __launch_bounds__(TPB) k_sieve_fill(const uint32_t* __restrict__ gramOfSlot, const uint32_t* __restrict__ gramOff,
                                                    const uint32_t* __restrict__ postings, int shift, uint32_t nGroups,
                                                    uint32_t* __restrict__ groups)
{
	const uint32_t slot = blockIdx.x;
	const uint32_t g = gramOfSlot[slot];
	const uint32_t lo = gramOff[g], hi = gramOff[g + 1];
	constexpr uint32_t W = INL ? 8 : 4;                                    // 32-bit words per group
	uint32_t* mine = groups + (size_t)slot * nGroups * W;
	for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
		const uint32_t b = postings[i] >> shift;
		const uint32_t v = b >> 5, gr = v / 3u, k = v - 3u * gr;            // virtual word v = 3 * group + k
		atomicOr(&mine[(size_t)gr * W + 1 + k], 1u << (b & 31));
	}
	for (uint32_t gr = threadIdx.x; gr < nGroups; gr += blockDim.x) {
		const unsigned long long firstId = ((unsigned long long)gr * SIEVE_GROUP_BLOCKS) << shift;
		uint32_t a = lo, z = hi;
		while (a < z) {
			const uint32_t m = a + ((z - a) >> 1);
			if ((unsigned long long)postings[m] < firstId) a = m + 1; else z = m;
		}
		mine[(size_t)gr * W] = a;
	}
	if (!INL) return;
	__syncthreads();                                                       // the bits of this list are complete
	const uint32_t bps = sieve_slot_bits(shift), nSlot = sieve_slots_per_group(shift);
	for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
		const uint32_t id = postings[i], b = id >> shift;
		if (i > lo && (postings[i - 1] >> shift) == b) continue;           // not the block's first posting
		const uint32_t multi = (i + 1 < hi && (postings[i + 1] >> shift) == b) ? 1u : 0u;
		const uint32_t v = b >> 5, gr = v / 3u, k = v - 3u * gr, bit = b & 31u;
		const uint32_t* w = mine + (size_t)gr * W;
		const uint32_t r = (k > 0 ? __popc(w[1]) : 0) + (k > 1 ? __popc(w[2]) : 0) + __popc(w[1 + k] & ((1u << bit) - 1u));
		if (r >= nSlot) continue;                                          // (k_sieve_search sends such blocks to the postings)
		const unsigned long long val = (unsigned long long)((multi << shift) | (id & ((1u << shift) - 1u)));
		const uint32_t at = r * bps, wd = at >> 5, sh = at & 31u;
		const unsigned long long sv = val << sh;
		atomicOr(&mine[(size_t)gr * W + 4 + wd], (uint32_t)sv);
		if ((uint32_t)(sv >> 32)) atomicOr(&mine[(size_t)gr * W + 4 + wd + 1], (uint32_t)(sv >> 32));
	}
}

}   // namespace

void build_sieve(DeviceIndex& ix, cudaStream_t s)
{
	ix.sieveShift = 0; ix.sieveNGroups = 0; ix.sieveSlots = 0;
	ix.sieveSlot.release(); ix.sieveGroups.release(); ix.sieveInline.release();
	static const bool off = getenv("SSB_NO_SIEVE") != nullptr;
	if (off || !ix.indexed || ix.G == 0 || ix.P == 0 || ix.T == 0) return;

	DevBuf<unsigned long long> acc(2, s);
	acc.zero(s);
	k_sieve_stats<<<grid_for(ix.G), TPB, 0, s>>>(ix.gramOff.get(), ix.G, acc.get());
	SSB_LAUNCH_CHECK();
	unsigned long long h[2] = {0, 0};
	SSB_CUDA(cudaMemcpyAsync(h, acc.get(), sizeof h, cudaMemcpyDeviceToHost, s));
	SSB_CUDA(cudaStreamSynchronize(s));
	if (h[0] == 0) return;
	const double lw = (double)h[1] / (double)h[0];          // posting-weighted mean list length
	// block size: the typical list sets about one bit in 32 (then "c of K lists hit the block" is rare for c >= 2 while a
	// bitmap costs about as many bytes as the list it summarises)
	int shift = (int)std::lround(std::log2(std::max(1.0, (double)ix.T / (32.0 * lw))));
	if (const char* e = getenv("SSB_SIEVE_SHIFT")) shift = atoi(e);
	shift = std::max(0, std::min(shift, 20));
	const uint64_t nBlocks = (((uint64_t)ix.T - 1) >> shift) + 1;
	const uint64_t nGroups = (nBlocks + SIEVE_GROUP_BLOCKS - 1) / SIEVE_GROUP_BLOCKS;
	// a gram gets a bitmap when its block bits (16 bytes per group) are at most twice as large as its posting list — or always,
	// on indexes so small that bitmaps for every gram stay within that budget overall
	uint32_t minLen = (uint32_t)std::max<uint64_t>(1, 2 * nGroups);
	if ((uint64_t)ix.G * nGroups * 4 <= 2ull * ix.P + (1u << 18)) minLen = 1;

	DevBuf<uint8_t> flag(ix.G, s);
	DevBuf<uint32_t> pos(ix.G, s), dN(1, s);
	k_sieve_flags<<<grid_for(ix.G), TPB, 0, s>>>(ix.gramOff.get(), ix.G, minLen, flag.get());
	SSB_LAUNCH_CHECK();
	exclusive_scan_u8_u32(flag.get(), pos.get(), ix.G, dN.get(), s);
	uint32_t nSlots = 0;
	SSB_CUDA(cudaMemcpyAsync(&nSlots, dN.get(), 4, cudaMemcpyDeviceToHost, s));
	SSB_CUDA(cudaStreamSynchronize(s));
	if (nSlots == 0) return;
	DevBuf<uint32_t> gramOfSlot(nSlots, s);
	ix.sieveSlot.alloc(ix.G, s);
	k_sieve_slots<<<grid_for(ix.G), TPB, 0, s>>>(flag.get(), pos.get(), ix.G, ix.sieveSlot.get(), gramOfSlot.get());
	SSB_LAUNCH_CHECK();
	ix.sieveGroups.alloc((size_t)nSlots * nGroups * 4, s);
	ix.sieveGroups.zero(s);
	k_sieve_fill<false><<<nSlots, TPB, 0, s>>>(gramOfSlot.get(), ix.gramOff.get(), ix.postings.get(), shift, (uint32_t)nGroups, ix.sieveGroups.get());
	SSB_LAUNCH_CHECK();
	// the 32-byte layout on top (twice the bytes again) where the first phase of a query covers a good part of the id range:
	// measured ahead at 10M keys (0.69 vs 0.72 ms per batch), behind at 100M (7.0 vs 5.7 ms: the stream dominates there)
	bool inl = ix.T <= SIEVE_INLINE_MAX_TERMS;
	if (const char* e = getenv("SSB_SIEVE_INLINE")) inl = atoi(e) != 0;
	if (inl) {
		ix.sieveInline.alloc((size_t)nSlots * nGroups * 8, s);
		ix.sieveInline.zero(s);
		k_sieve_fill<true><<<nSlots, TPB, 0, s>>>(gramOfSlot.get(), ix.gramOff.get(), ix.postings.get(), shift, (uint32_t)nGroups, ix.sieveInline.get());
		SSB_LAUNCH_CHECK();
	}
	ix.sieveShift = shift; ix.sieveNGroups = (uint32_t)nGroups; ix.sieveSlots = nSlots;      // (scratch buffers are freed in stream order)
}

}   // namespace ssb
