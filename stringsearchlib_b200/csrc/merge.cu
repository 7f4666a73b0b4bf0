// merge.cu — multi-GPU top-k merge (no reference equivalent).
// Every key-range shard answers the whole query batch and exports, per query, `limit` fixed-size records
// (score, key length, global first row, key hash). After one all-gather (NCCL, done by the caller) each rank
// merges: equal key strings from different shards collapse to their best score (the reference's per-key max,
// nGramSearch.hpp:326), then the global order (score desc, key length asc, first row asc) is cut at `limit`.
#include "search.h"
#include "textops.cuh"
#include <algorithm>

namespace ssb {
namespace {

__global__ void k_export_records(const uint32_t* __restrict__ outCount, const uint32_t* __restrict__ outKid, const float* __restrict__ outScore,
                                 uint32_t stride, uint32_t nq, uint32_t limit, int unit, const uint32_t* __restrict__ keyOff,
                                 const uint8_t* __restrict__ keyChars, const uint64_t* __restrict__ keyRow, MergeRecord* __restrict__ out)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (uint64_t)nq * limit) return;
	const uint32_t q = (uint32_t)(i / limit), r = (uint32_t)(i % limit);
	MergeRecord rec;
	if (r < outCount[q] && r < stride) {
		const uint32_t kid = outKid[(size_t)q * stride + r];
		const uint32_t b = keyOff[kid], len = keyOff[kid + 1] - b - 1;
		uint64_t h = 0x6b657973ull;
		if (unit == 1) for (uint32_t j = 0; j < len; ++j) h = hash_step(h, keyChars[b + j]);
		else for (uint32_t j = 0; j < len; ++j) h = hash_step(h, reinterpret_cast<const uint32_t*>(keyChars)[b + j]);
		rec.scoreBits = __float_as_uint(outScore[(size_t)q * stride + r]);
		rec.keyLen = len;
		rec.row = keyRow[kid];
		rec.hash = hash_done(h, len);
	} else {
		rec.scoreBits = 0; rec.keyLen = 0; rec.row = ~0ull; rec.hash = ~0ull;
	}
	out[i] = rec;
}

template <bool BY_HASH>
__device__ __forceinline__ bool rec_less(uint64_t a1, uint64_t b1, uint64_t h1, uint64_t a2, uint64_t b2, uint64_t h2)
{
	if (BY_HASH && h1 != h2) return h1 < h2;
	if (a1 != a2) return a1 < a2;
	return b1 < b2;
}
template <bool BY_HASH>
__device__ void bitonic3(uint64_t* A, uint64_t* B, uint64_t* H, uint32_t p)
{
	for (uint32_t size = 2; size <= p; size <<= 1)
		for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
			for (uint32_t i = threadIdx.x; i < (p >> 1); i += blockDim.x) {
				uint32_t lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
				bool up = (lo & size) == 0;
				bool lt = rec_less<BY_HASH>(A[hi], B[hi], H[hi], A[lo], B[lo], H[lo]);   // hi < lo
				if (lt == up) {
					uint64_t t;
					t = A[lo]; A[lo] = A[hi]; A[hi] = t;
					t = B[lo]; B[lo] = B[hi]; B[hi] = t;
					t = H[lo]; H[lo] = H[hi]; H[hi] = t;
				}
			}
			__syncthreads();
		}
}

// one CTA per query
__global__ void k_merge_records(const MergeRecord* __restrict__ gathered, uint32_t nShards, uint32_t nq, uint32_t limit, uint32_t p,
                                MergeRecord* __restrict__ merged, uint32_t* __restrict__ counts)
{
	extern __shared__ __align__(16) unsigned char mg_smem[];
	uint64_t* A = reinterpret_cast<uint64_t*>(mg_smem);
	uint64_t* B = A + p;
	uint64_t* H = B + p;
	__shared__ uint32_t sCount;
	const uint32_t q = blockIdx.x;
	const uint32_t n = nShards * limit;
	for (uint32_t i = threadIdx.x; i < p; i += blockDim.x) {
		uint64_t a = ~0ull, b = ~0ull, h = ~0ull;
		if (i < n) {
			const uint32_t sh = i / limit, r = i % limit;
			const MergeRecord rec = gathered[((size_t)sh * nq + q) * limit + r];
			if (rec.row != ~0ull) {
				a = ((uint64_t)(~float_to_ordered(__uint_as_float(rec.scoreBits))) << 32) | rec.keyLen;
				b = rec.row; h = rec.hash;
			}
		}
		A[i] = a; B[i] = b; H[i] = h;
	}
	if (threadIdx.x == 0) sCount = 0;
	__syncthreads();
	bitonic3<true>(A, B, H, p);
	// equal (hash, length) = same key string seen by several shards: keep the best (first) one
	for (uint32_t base = 0; base < p; base += blockDim.x) {
		uint32_t i = base + threadIdx.x;
		bool dead = false;
		if (i > 0 && i < p) dead = H[i] == H[i - 1] && (uint32_t)A[i] == (uint32_t)A[i - 1] && A[i] != ~0ull;
		__syncthreads();
		if (dead) { A[i] = ~0ull; B[i] = ~0ull; }
		__syncthreads();
	}
	bitonic3<false>(A, B, H, p);
	for (uint32_t i = threadIdx.x; i < limit; i += blockDim.x) {
		MergeRecord rec;
		if (i < p && A[i] != ~0ull) {
			rec.scoreBits = __float_as_uint(ordered_to_float(~(uint32_t)(A[i] >> 32)));
			rec.keyLen = (uint32_t)A[i];
			rec.row = B[i]; rec.hash = H[i];
			atomicAdd(&sCount, 1u);
		} else { rec.scoreBits = 0; rec.keyLen = 0; rec.row = ~0ull; rec.hash = ~0ull; }
		merged[(size_t)q * limit + i] = rec;
	}
	__syncthreads();
	if (threadIdx.x == 0) counts[q] = sCount;
}

}   // namespace

void export_records(const DeviceIndex& ix, SearchWorkspace& ws, uint32_t nq, uint32_t limit, MergeRecord* d_out, cudaStream_t s)
{
	if (limit == 0 || limit > FAST_KLIM) throw std::runtime_error("export_records: limit must be in [1, FAST_KLIM] on the sharded path");
	if (nq > ws.capQ) throw std::runtime_error("export_records: no such batch");
	const uint64_t n = (uint64_t)nq * limit;
	if (!ix.indexed || ix.K == 0) {
		// an empty shard contributes only empty slots
		SSB_CUDA(cudaMemsetAsync(d_out, 0xFF, n * sizeof(MergeRecord), s));
		return;
	}
	k_export_records<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ws.outCount.get(), ws.outKid.get(), ws.outScore.get(), limit, nq, limit,
	                                                             ix.unit, ix.keyOff.get(), ix.keyChars.get(), ix.keyRow.get(), d_out);
	SSB_LAUNCH_CHECK();
}

void merge_records(const MergeRecord* d_gathered, uint32_t nShards, uint32_t nq, uint32_t limit, MergeRecord* d_out, uint32_t* d_outCount,
                   cudaStream_t s)
{
	const uint32_t n = nShards * limit;
	uint32_t p = 2;
	while (p < n) p <<= 1;
	const size_t smem = (size_t)p * 24;
	if (smem > 200 * 1024) throw std::runtime_error("merge_records: nShards * limit too large for one CTA (max 8192 records)");
	static size_t attr = 0;
	if (smem > attr) {
		SSB_CUDA(cudaFuncSetAttribute(k_merge_records, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		attr = smem;
	}
	const unsigned threads = std::min<uint32_t>(1024, std::max<uint32_t>(32, p / 2));
	k_merge_records<<<nq, threads, smem, s>>>(d_gathered, nShards, nq, limit, p, d_out, d_outCount);
	SSB_LAUNCH_CHECK();
}

}   // namespace ssb
