// search.cu — batched n-gram search on the GPU (replaces StringIndex::_search / searchLong / searchShort /
// calcScore / partial_sort, nGramSearch.hpp:182-404).
//
//   k_query_prep   one warp per query: escapeBlank/trim/toUpper (hpp:372-376), gram windows -> codes (hpp:29-36),
//                  warp match-any de-duplication with multiplicities, binary search of the sorted gram table,
//                  integer thresholds cmin/cprom evaluated with the reference's exact float expressions.
//   k_fast_search  persistent CTAs, one query at a time. The query's posting lists are streamed through shared
//                  memory in term-id-aligned windows; every thread owns a term-id sub-range and does a K-way merge
//                  of the (sorted) windows in registers, which yields (term, match count) with NO atomics and no
//                  per-term counters. Scores use the reference's float ops ((float)c / (float)|Q|, w * s), candidates
//                  are filtered against a running k-th best (tau) and kept in a shared-memory buffer that is reduced
//                  with an in-CTA radix select; the final <= limit entries are bitonic-sorted. Rank key =
//                  (score desc, kid asc) where kid already encodes (key length asc, key id asc) (nGramSearch.h:262-269).
//   dense path     general fallback (short terms / edit-distance path hpp:182-270, wildcard hpp:356-369, long
//                  queries, limit > FAST_KLIM, buffer overflow): per-term counters + per-key max-merge in global
//                  memory, then a device radix sort. O(T + K) per query, exact for every input.
#include "search.h"
#include "primitives.cuh"
#include "textops.cuh"
#include <algorithm>
#include <cmath>
#include <mutex>

namespace ssb {

SearchWorkspace::~SearchWorkspace()
{
	if (g_exiting.load()) return;
	if (ev0) cudaEventDestroy(ev0);
	if (ev1) cudaEventDestroy(ev1);
	if (ev2) cudaEventDestroy(ev2);
	if (ev3) cudaEventDestroy(ev3);
}

namespace {

struct IndexView {
	int g, unit, symBits;
	uint32_t K, T, G, A, nShort;
	int identity, uniformW, multiEdgeKeys;
	float wmax;
	const uint8_t* validLut;
	const uint32_t* symTable;
	const uint64_t* gramCodes;
	const uint32_t* gramOff;
	const uint32_t* postings;
	const uint32_t* termOff;
	const uint8_t* termChars;
	const uint32_t* termEdgeOff;
	const uint32_t* edgeKid;
	const float* edgeW;
	const uint32_t* shortTerms;
	const uint32_t* keyOff;
	const uint8_t* keyChars;
	const uint32_t* keyEdgeOff;
	const uint32_t* keyEdgeTerm;
	const float* keyEdgeW;
	const float* keyWmax;
	const uint64_t* keyRow;
	// sieve (block bitmaps, sieve.cu); sieveNGroups == 0: none
	int sieveShift;
	uint32_t sieveNGroups;
	const uint32_t* sieveSlot;
	const uint4* sieveGroups;                               // 16-byte groups
	const uint4* sieveInline;                               // 32-byte groups (nullptr: not built for this index)
};

IndexView make_view(const DeviceIndex& ix)
{
	IndexView v;
	v.g = ix.g; v.unit = ix.unit; v.symBits = ix.symBits;
	v.K = ix.K; v.T = ix.T; v.G = ix.G; v.A = ix.A; v.nShort = ix.nShort;
	v.identity = ix.identity; v.uniformW = ix.uniformW; v.multiEdgeKeys = ix.multiEdgeKeys;
	v.wmax = ix.wmax;
	v.validLut = ix.validLut.get(); v.symTable = ix.symTable.get();
	v.gramCodes = ix.gramCodes.get(); v.gramOff = ix.gramOff.get(); v.postings = ix.postings.get();
	v.termOff = ix.termOff.get(); v.termChars = ix.termChars.get();
	v.termEdgeOff = ix.termEdgeOff.get(); v.edgeKid = ix.edgeKid.get(); v.edgeW = ix.edgeW.get();
	v.shortTerms = ix.shortTerms.get();
	v.keyOff = ix.keyOff.get(); v.keyChars = ix.keyChars.get();
	v.keyEdgeOff = ix.keyEdgeOff.get(); v.keyEdgeTerm = ix.keyEdgeTerm.get(); v.keyEdgeW = ix.keyEdgeW.get();
	v.keyWmax = ix.keyWmax.get(); v.keyRow = ix.keyRow.get();
	v.sieveShift = ix.sieveShift; v.sieveNGroups = ix.sieveSlots ? ix.sieveNGroups : 0u;
	v.sieveSlot = ix.sieveSlot.get(); v.sieveGroups = reinterpret_cast<const uint4*>(ix.sieveGroups.get());
	v.sieveInline = reinterpret_cast<const uint4*>(ix.sieveInline.get());
	return v;
}

// the reference's score of a long-path term: (float)count / (float)grams.size()   (hpp:300)
__device__ __forceinline__ float term_score(uint32_t c, uint32_t n) { return __fdiv_rn((float)c, (float)n); }
// std::max(w * s, entryScore-default 0.0f) with -0.0f canonicalised (hpp:326)
__device__ __forceinline__ float edge_value(float w, float s)
{
	float v = __fmul_rn(w, s);
	return v > 0.0f ? v : 0.0f;
}

// trim(escapeBlank(key)) == query ?  The key is NOT upper-cased (hpp:330-335).
template <class U>
__device__ bool key_equals_query(const IndexView& ix, uint32_t kid, const U* __restrict__ q, uint32_t qlen)
{
	const U* key = reinterpret_cast<const U*>(ix.keyChars) + ix.keyOff[kid];
	uint32_t b = 0, e = ix.keyOff[kid + 1] - ix.keyOff[kid] - 1;
	while (b < e && unit_blank(key[b], ix.validLut)) ++b;
	while (e > b && unit_blank(key[e - 1], ix.validLut)) --e;
	if (e - b != qlen) return false;
	for (uint32_t j = 0; j < qlen; ++j)
		if (unit_escape(key[b + j], ix.validLut) != q[j]) return false;
	return true;
}

// match count of the query against one long term, straight from the term's characters:
// sum over query windows (duplicates kept) of [window gram occurs in the term]   (hpp:289-298 restated per term)
template <class U>
__device__ uint32_t count_for_term(const IndexView& ix, uint32_t t, const uint64_t* __restrict__ qcode, uint32_t nGrams)
{
	const U* tc = reinterpret_cast<const U*>(ix.termChars);
	uint32_t b = ix.termOff[t], len = ix.termOff[t + 1] - b;
	if (len < 2u * ix.g) return 0;
	const uint64_t mask = (ix.g * ix.symBits >= 64) ? ~0ull : ((1ull << (ix.g * ix.symBits)) - 1);
	uint64_t code = 0;
	uint32_t hits = 0;
	for (uint32_t j = 0; j < len; ++j) {
		code = ((code << ix.symBits) | sym_of(tc[b + j], ix.symTable, ix.A)) & mask;
		if (j + 1 >= (uint32_t)ix.g)
			for (uint32_t i = 0; i < nGrams; ++i)
				if (qcode[i] == code) hits |= 1u << i;
	}
	return __popc(hits);
}

// stringMatch (hpp:182-222) restated column-wise: semi-global edit distance, free start and end in the source
template <class U>
__device__ uint32_t string_match(const U* __restrict__ q, uint32_t nq, const U* __restrict__ s, uint32_t ns)
{
	if (nq == 1) {
		for (uint32_t j = 0; j < ns; ++j) if (s[j] == q[0]) return 1;
		return 0;
	}
	uint32_t col[3 * MAX_GSIZE + 1];
	for (uint32_t a = 0; a <= nq; ++a) col[a] = a;
	uint32_t best = col[nq];
	for (uint32_t j = 0; j < ns; ++j) {
		uint32_t diag = col[0];          // D[a][j]
		const U sc = s[j];
		for (uint32_t a = 0; a < nq; ++a) {
			uint32_t up = col[a + 1];    // D[a+1][j]
			uint32_t v = min(min(col[a] + 1, up + 1), diag + (q[a] != sc ? 1u : 0u));
			diag = up;
			col[a + 1] = v;
		}
		best = min(best, col[nq]);
	}
	return nq - best;
}
// ---------------------------------------------------------------------------------------------- query prep
template <class U>
__global__ void __launch_bounds__(256) k_query_prep(IndexView ix, const U* __restrict__ qchars, const uint32_t* __restrict__ qoff,
                                                    uint32_t nq, float threshold, uint32_t limit, U* __restrict__ qnorm,
                                                    QDesc* __restrict__ desc, uint32_t* __restrict__ fastList,
                                                    uint32_t* __restrict__ denseList, uint32_t* __restrict__ counters,
                                                    uint32_t* __restrict__ outCount)
{
	const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (q >= nq) return;
	QDesc& d = desc[q];
	uint32_t b = qoff[q], e = qoff[q + 1];
	const uint32_t g = ix.g;
	// wildcard: raw query "" or "*" (hpp:356)
	const bool wildcard = (e == b) || (e - b == 1 && qchars[b] == (U)'*');
	// trim after escapeBlank (hpp:372-373)
	uint32_t tb = b, te = e;
	if (!wildcard) {
		while (tb < te && unit_blank(qchars[tb], ix.validLut)) ++tb;
		while (te > tb && unit_blank(qchars[te - 1], ix.validLut)) --te;
	}
	const uint32_t qlen = wildcard ? 0 : te - tb;
	for (uint32_t j = lane; j < qlen; j += 32) qnorm[b + j] = unit_norm(qchars[tb + j], ix.validLut);   // toUpper (hpp:376)
	__syncwarp();
	const uint32_t nGrams = qlen >= g ? qlen - g + 1 : 0;
	uint32_t cls;
	const bool shortActive = qlen < 3 * g && (ix.nShort > 0 || qlen <= g);           // hpp:381, 240-253
	if (wildcard) cls = QC_WILDCARD;
	else if (qlen == 0) cls = QC_EMPTY;                                               // hpp:374-375
	// the fused kernel handles the edit-distance path only in its variant for indexes WITH short terms; on an index without
	// them the path runs only for |q| <= gSize (scan of every long term), which is left to the dense path
	else if ((nGrams >= 1 || shortActive) && (!shortActive || ix.nShort > 0) && nGrams <= MAX_QGRAMS && limit <= FAST_KLIM) cls = QC_FAST;
	else cls = QC_DENSE;

	uint32_t nLists = 0, total = 0;
	bool sieveOk = false;
	if (cls == QC_FAST || (cls == QC_DENSE && nGrams >= 1 && nGrams <= MAX_QGRAMS)) {
		// lane i owns window i
		const bool act = (uint32_t)lane < nGrams;
		uint64_t code = ~0ull;
		if (act) {
			code = 0;
			bool ok = true;
			for (uint32_t j = 0; j < g; ++j) {
				uint32_t sy = sym_of(qnorm[b + lane + j], ix.symTable, ix.A);
				ok &= sy != 0;
				code = (code << ix.symBits) | sy;
			}
			if (!ok) code = ~0ull;
		}
		d.qcode[lane] = code;
		// duplicates: one list per distinct gram, multiplicity = number of windows (hpp:289-298 adds once per window)
		unsigned peers = __match_any_sync(0xffffffffu, code);
		const bool leader = act && code != ~0ull && (peers & ((1u << lane) - 1)) == 0;
		uint32_t lo = 0, len = 0, bm = 0xFFFFFFFFu;
		if (leader) {
			uint32_t a = 0, z = ix.G;
			while (a < z) { uint32_t m = (a + z) >> 1; if (ix.gramCodes[m] < code) a = m + 1; else z = m; }
			if (a < ix.G && ix.gramCodes[a] == code) {
				lo = ix.gramOff[a]; len = ix.gramOff[a + 1] - lo;
				if (ix.sieveNGroups) bm = ix.sieveSlot[a];
			}
		}
		unsigned found = __ballot_sync(0xffffffffu, leader && len > 0);
		if (leader && len > 0) {
			uint32_t slot = __popc(found & ((1u << lane) - 1));
			d.listOff[slot] = lo; d.listLen[slot] = len; d.listMult[slot] = (uint8_t)__popc(peers); d.listSlot[slot] = bm;
		}
		nLists = __popc(found);
		sieveOk = ix.sieveNGroups != 0 && nLists >= 1 && nLists <= SV_MAXL && !__any_sync(0xffffffffu, leader && len > 0 && bm == 0xFFFFFFFFu);
		total = len;
		for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
	}
	if (lane == 0) {
		uint32_t cmin = nGrams + 1, cprom = nGrams + 1;
		for (uint32_t c = nGrams; c >= 1; --c) {
			float s = term_score(c, nGrams);
			if (!(s < threshold)) cmin = c;                                           // hpp:315
			if ((double)s > 0.999) cprom = c;                                         // hpp:328
		}
		d.cls = cls; d.qlen = qlen; d.nGrams = nGrams; d.nLists = nLists; d.cmin = cmin; d.cprom = cprom; d.total = total; d.qbeg = b;
		d.flags = (shortActive ? 1u : 0u) | ((shortActive && qlen <= g) ? 2u : 0u) | ((sieveOk && !shortActive && cls == QC_FAST) ? 4u : 0u);
		outCount[q] = 0;
		if (cls == QC_FAST) {
			fastList[atomicAdd(&counters[0], 1u)] = q;
			// posting volume of the batch's fast queries, the figure the choice of kernel is made from (run_device). A query
			// that needs two or more matches from the start (a threshold) has no dense first phase in k_sieve_search and
			// counts five-fold: measured, the sieve is 2x ahead of the merge there from ~18 000 postings per query on
			atomicAdd(reinterpret_cast<unsigned long long*>(counters + 30), (unsigned long long)total * (cmin >= 2 ? 5u : 1u));
		}
		else if (cls == QC_DENSE || cls == QC_WILDCARD) denseList[atomicAdd(&counters[1], 1u)] = q;
	}
}

}   // namespace

#include "fast_path.cuh"   // k_fast_search (inside namespace ssb)
#include "sieve_path.cuh"  // k_sieve_search

namespace {

// Longest-processing-time-first order of the fast queries: with only a handful of queries per persistent CTA the tail of
// a batch is set by whichever long query starts last; starting the heavy ones first shortens it. An exact order is not
// needed: one CTA does a 256-bin counting sort by posting volume (descending); larger batches are left as they are.
constexpr uint32_t ORDER_MAX = 16384;
__global__ void __launch_bounds__(1024) k_order_fast(const QDesc* __restrict__ desc, uint32_t* __restrict__ fastList,
                                                     const uint32_t* __restrict__ counters)
{
	extern __shared__ __align__(16) unsigned char order_smem[];
	uint32_t* qs = reinterpret_cast<uint32_t*>(order_smem);          // [ORDER_MAX] query ids
	uint32_t* bins = qs + ORDER_MAX;                                 // [ORDER_MAX] bin of each entry
	__shared__ uint32_t hist[256], maxTotal;
	const uint32_t n = counters[0];
	if (n < 2 || n > ORDER_MAX) return;
	if (threadIdx.x < 256) hist[threadIdx.x] = 0;
	if (threadIdx.x == 0) maxTotal = 0;
	__syncthreads();
	uint32_t localMax = 0;
	for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
		const uint32_t q = fastList[i];
		qs[i] = q;
		// estimated work: the merge costs ~(4 + K) instructions per posting with K lists (min tree + one compare / advance per list)
		bins[i] = (uint32_t)min((unsigned long long)desc[q].total * (4u + desc[q].nLists) >> 4, 0xFFFFFFFFull);
		localMax = max(localMax, bins[i]);
	}
	for (int o = 16; o; o >>= 1) localMax = max(localMax, __shfl_xor_sync(0xffffffffu, localMax, o));
	if ((threadIdx.x & 31) == 0) atomicMax(&maxTotal, localMax);
	__syncthreads();
	const uint32_t div = maxTotal / 255 + 1;
	for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
		const uint32_t b = 255 - bins[i] / div;                      // heavy queries -> low bins -> first
		bins[i] = b;
		atomicAdd(&hist[b], 1u);
	}
	__syncthreads();
	if (threadIdx.x < 32) {                                          // exclusive scan of 256 bins by one warp
		uint32_t loc[8], sum = 0;
#pragma unroll
		for (int j = 0; j < 8; ++j) { loc[j] = hist[threadIdx.x * 8 + j]; sum += loc[j]; }
		uint32_t inc = sum;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (threadIdx.x >= o) inc += t; }
		uint32_t run = inc - sum;
#pragma unroll
		for (int j = 0; j < 8; ++j) { hist[threadIdx.x * 8 + j] = run; run += loc[j]; }
	}
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) fastList[atomicAdd(&hist[bins[i]], 1u)] = qs[i];
}

// ---------------------------------------------------------------------------------------------- dense path
// one CTA per gram window of the query: counts[t] += 1 for every term of the window's posting list (hpp:289-298)
template <class U>
__global__ void __launch_bounds__(256) k_dense_count(IndexView ix, const U* __restrict__ qn, uint32_t qlen, uint32_t* __restrict__ cnt)
{
	__shared__ uint32_t sLo, sLen;
	const uint32_t w = blockIdx.x;
	if (threadIdx.x == 0) {
		uint64_t code = 0;
		bool ok = true;
		for (int j = 0; j < ix.g; ++j) {
			uint32_t sy = sym_of(qn[w + j], ix.symTable, ix.A);
			ok &= sy != 0;
			code = (code << ix.symBits) | sy;
		}
		uint32_t lo = 0, len = 0;
		if (ok) {
			uint32_t a = 0, z = ix.G;
			while (a < z) { uint32_t m = (a + z) >> 1; if (ix.gramCodes[m] < code) a = m + 1; else z = m; }
			if (a < ix.G && ix.gramCodes[a] == code) { lo = ix.gramOff[a]; len = ix.gramOff[a + 1] - lo; }
		}
		sLo = lo; sLen = len;
	}
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < sLen; i += blockDim.x) atomicAdd(&cnt[ix.postings[sLo + i]], 1u);
	(void)qlen;
}

// per-key max-merge: entry[kid] = max(entry[kid], value) on order-preserving bits; 0 = key absent
template <class U>
__device__ __forceinline__ void dense_apply_term(const IndexView& ix, uint32_t t, float s, float threshold, const U* __restrict__ qn,
                                                 uint32_t qlen, uint32_t* __restrict__ entry)
{
	if (s < threshold) return;                                                        // hpp:315
	const bool perfect = (double)s > 0.999;                                           // hpp:328
	for (uint32_t e = ix.termEdgeOff[t]; e < ix.termEdgeOff[t + 1]; ++e) {
		const uint32_t kid = ix.edgeKid[e];
		float v = edge_value(ix.edgeW[e], s);
		if (perfect && key_equals_query<U>(ix, kid, qn, qlen)) v = 100.0f;
		atomicMax(&entry[kid], float_to_ordered(v));
	}
}
template <class U>
__global__ void __launch_bounds__(256) k_dense_long(IndexView ix, const uint32_t* __restrict__ cnt, uint32_t nGrams, float threshold,
                                                    const U* __restrict__ qn, uint32_t qlen, uint32_t* __restrict__ entry)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= ix.T) return;
	uint32_t c = cnt[t];
	if (c) dense_apply_term<U>(ix, t, term_score(c, nGrams), threshold, qn, qlen, entry);
}
// short path (hpp:232-254): every short term, and every long term too when |q| <= gSize
template <class U>
__global__ void __launch_bounds__(256) k_dense_short(IndexView ix, uint32_t nScan, int allTerms, float threshold, const U* __restrict__ qn,
                                                     uint32_t qlen, uint32_t* __restrict__ entry)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nScan) return;
	uint32_t t;
	if (allTerms) {
		t = i;
		if (ix.termOff[t + 1] - ix.termOff[t] < 2u * ix.g) return;      // short terms are handled by the other launch
	} else t = ix.shortTerms[i];
	const U* tc = reinterpret_cast<const U*>(ix.termChars) + ix.termOff[t];
	uint32_t m = string_match<U>(qn, qlen, tc, ix.termOff[t + 1] - ix.termOff[t]);
	dense_apply_term<U>(ix, t, __fdiv_rn((float)m, (float)qlen), threshold, qn, qlen, entry);   // hpp:244
}
__global__ void k_dense_wildcard(IndexView ix, uint32_t* __restrict__ entry)
{
	uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= ix.K) return;
	if (ix.keyEdgeOff[k + 1] > ix.keyEdgeOff[k]) entry[k] = float_to_ordered(ix.keyWmax[k]);   // hpp:358-368
}
__global__ void k_dense_flags(const uint32_t* __restrict__ entry, uint32_t K, uint8_t* __restrict__ flag)
{
	uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < K) flag[k] = entry[k] != 0;
}
__global__ void k_dense_collect(const uint32_t* __restrict__ entry, const uint8_t* __restrict__ flag, const uint32_t* __restrict__ pos,
                                uint32_t K, uint64_t* __restrict__ keys)
{
	uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < K && flag[k]) keys[pos[k]] = ((uint64_t)(~entry[k]) << 32) | k;
}
__global__ void k_dense_emit(const uint64_t* __restrict__ sorted, uint32_t n, uint32_t* __restrict__ kid, float* __restrict__ score)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { kid[i] = (uint32_t)sorted[i]; score[i] = rank_key_score(sorted[i]); }
}

// ---------------------------------------------------------------------------------------------- CSR finalisation
__global__ void k_clamp_counts(const uint32_t* __restrict__ outCount, uint32_t nq, uint32_t stride, uint32_t* __restrict__ clamped,
                               uint32_t* __restrict__ anyBig)
{
	uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= nq) return;
	uint32_t c = outCount[q];
	if (c > stride) { c = 0; *anyBig = 1; }       // 0xFFFFFFFF marks a result list delivered out of band
	clamped[q] = c;
}
// one warp per query: fixed-stride rows -> CSR, plus the key's global row and the host address of its string
__global__ void __launch_bounds__(256) k_finalize_csr(const uint32_t* __restrict__ count, const uint32_t* __restrict__ off,
                                                      const uint32_t* __restrict__ outKid, const float* __restrict__ outScore,
                                                      uint32_t nq, uint32_t stride, const uint32_t* __restrict__ keyOff,
                                                      const uint64_t* __restrict__ keyRow, uint64_t strBase, uint32_t unit,
                                                      uint32_t* __restrict__ csrKid, float* __restrict__ csrScore,
                                                      uint64_t* __restrict__ csrRow, uint64_t* __restrict__ csrStr)
{
	const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (q >= nq) return;
	const uint32_t n = count[q], o = off[q];
	for (uint32_t i = lane; i < n; i += 32) {
		const uint32_t kid = outKid[(size_t)q * stride + i];
		csrKid[o + i] = kid;
		csrScore[o + i] = outScore[(size_t)q * stride + i];
		csrRow[o + i] = keyRow[kid];
		csrStr[o + i] = strBase + (uint64_t)keyOff[kid] * unit;
	}
}

// tiny batches (<= 32 queries): clamp + scan + gather in ONE CTA that stores straight into the caller's pinned result arrays
// and the pinned counters (zero-copy over PCIe) — a single search() call then costs two small uploads, a memset, three kernels
// and one synchronisation instead of ~20 API calls
__global__ void __launch_bounds__(256) k_finalize_tiny(const uint32_t* __restrict__ outCount, const uint32_t* __restrict__ outKid,
                                                       const float* __restrict__ outScore, uint32_t nq, uint32_t stride,
                                                       const uint32_t* __restrict__ keyOff, const uint64_t* __restrict__ keyRow,
                                                       uint64_t strBase, uint32_t unit, uint32_t* __restrict__ counters,
                                                       uint32_t* __restrict__ hOff, uint32_t* __restrict__ hKid, float* __restrict__ hScore,
                                                       uint64_t* __restrict__ hRow, uint64_t* __restrict__ hStr, uint32_t* __restrict__ hCounters)
{
	__shared__ uint32_t off[33];
	const uint32_t tid = threadIdx.x;
	if (tid < 32) {
		uint32_t c = tid < nq ? outCount[tid] : 0u;
		const bool big = c > stride;                      // delivered out of band (dense path with limit > stride)
		if (big) c = 0;
		if (__any_sync(0xffffffffu, big) && tid == 0) counters[4] = 1;
		uint32_t inc = c;
		for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= (uint32_t)o) inc += t; }
		off[tid + 1] = inc;
		if (tid == 0) off[0] = 0;
	}
	__syncthreads();
	for (uint32_t q = 0; q < nq; ++q) {
		const uint32_t o = off[q], n = off[q + 1] - o;
		for (uint32_t i = tid; i < n; i += blockDim.x) {
			const uint32_t kid = outKid[(size_t)q * stride + i];
			hKid[o + i] = kid;
			hScore[o + i] = outScore[(size_t)q * stride + i];
			hRow[o + i] = keyRow[kid];
			hStr[o + i] = strBase + (uint64_t)keyOff[kid] * unit;
		}
	}
	if (tid <= nq) hOff[tid] = off[tid];
	if (tid < 32) hCounters[tid] = counters[tid];
}

inline unsigned grid_for(size_t n, int tpb = 256) { return (unsigned)std::max<size_t>(1, (n + tpb - 1) / tpb); }

template <class U>
void run_dense_query(const DeviceIndex& ix, const IndexView& v, SearchWorkspace& ws, const QDesc& d, uint32_t q, float threshold,
                     uint32_t limit, uint32_t stride, BatchOutput* out, cudaStream_t s)
{
	const U* qn = reinterpret_cast<const U*>(ws.qnorm.get()) + d.qbeg;
	ws.entry.reserve(std::max(ix.K, 1u), s);
	SSB_CUDA(cudaMemsetAsync(ws.entry.get(), 0, (size_t)ix.K * sizeof(uint32_t), s));
	if (d.cls == QC_WILDCARD) {
		k_dense_wildcard<<<grid_for(ix.K), 256, 0, s>>>(v, ws.entry.get());
		SSB_LAUNCH_CHECK();
	} else {
		const uint32_t g = ix.g;
		if (d.qlen < 3 * g) {                                                           // hpp:381
			if (ix.nShort) {
				k_dense_short<U><<<grid_for(ix.nShort), 256, 0, s>>>(v, ix.nShort, 0, threshold, qn, d.qlen, ws.entry.get());
				SSB_LAUNCH_CHECK();
			}
			if (d.qlen <= g) {                                                          // hpp:247
				k_dense_short<U><<<grid_for(ix.T), 256, 0, s>>>(v, ix.T, 1, threshold, qn, d.qlen, ws.entry.get());
				SSB_LAUNCH_CHECK();
			}
		}
		if (d.nGrams >= 1 && ix.P) {                                                    // hpp:281
			ws.cnt.reserve(ix.T, s);
			SSB_CUDA(cudaMemsetAsync(ws.cnt.get(), 0, (size_t)ix.T * sizeof(uint32_t), s));
			k_dense_count<U><<<d.nGrams, 256, 0, s>>>(v, qn, d.qlen, ws.cnt.get());
			SSB_LAUNCH_CHECK();
			k_dense_long<U><<<grid_for(ix.T), 256, 0, s>>>(v, ws.cnt.get(), d.nGrams, threshold, qn, d.qlen, ws.entry.get());
			SSB_LAUNCH_CHECK();
		}
	}
	// collect present keys, sort by rank key, keep the first `limit`
	DevBuf<uint8_t> flag(ix.K, s);
	DevBuf<uint32_t> pos(ix.K, s), dN(1, s);
	k_dense_flags<<<grid_for(ix.K), 256, 0, s>>>(ws.entry.get(), ix.K, flag.get());
	SSB_LAUNCH_CHECK();
	exclusive_scan_u8_u32(flag.get(), pos.get(), ix.K, dN.get(), s);
	uint32_t n = 0;
	SSB_CUDA(cudaMemcpyAsync(&n, dN.get(), 4, cudaMemcpyDeviceToHost, s));
	SSB_CUDA(cudaStreamSynchronize(s));
	uint32_t take = std::min(n, limit);
	if (n) {
		DevBuf<uint64_t> k1(n, s), k2(n, s);
		k_dense_collect<<<grid_for(ix.K), 256, 0, s>>>(ws.entry.get(), flag.get(), pos.get(), ix.K, k1.get());
		SSB_LAUNCH_CHECK();
		int kidBits = 1;
		while ((1ull << kidBits) < ix.K) ++kidBits;
		bool a1 = radix_sort_u64(k1.get(), k2.get(), nullptr, nullptr, n, 0, kidBits, s);
		uint64_t *ka = a1 ? k2.get() : k1.get(), *kb = a1 ? k1.get() : k2.get();
		bool a2 = radix_sort_u64(ka, kb, nullptr, nullptr, n, 32, 64, s);
		const uint64_t* sorted = a2 ? kb : ka;
		if (take <= stride) {
			k_dense_emit<<<grid_for(take), 256, 0, s>>>(sorted, take, ws.outKid.get() + (size_t)q * stride,
			                                           ws.outScore.get() + (size_t)q * stride);
			SSB_LAUNCH_CHECK();
			SSB_CUDA(cudaMemcpyAsync(ws.outCount.get() + q, &take, 4, cudaMemcpyHostToDevice, s));
			SSB_CUDA(cudaStreamSynchronize(s));   // `take` lives on this stack frame
		} else if (out) {
			DevBuf<uint32_t> dk(take, s);
			DevBuf<float> dsc(take, s);
			k_dense_emit<<<grid_for(take), 256, 0, s>>>(sorted, take, dk.get(), dsc.get());
			SSB_LAUNCH_CHECK();
			std::vector<uint32_t> hkid(take);
			std::vector<float> hsc(take);
			SSB_CUDA(cudaMemcpyAsync(hkid.data(), dk.get(), (size_t)take * 4, cudaMemcpyDeviceToHost, s));
			SSB_CUDA(cudaMemcpyAsync(hsc.data(), dsc.get(), (size_t)take * 4, cudaMemcpyDeviceToHost, s));
			uint32_t marker = 0xFFFFFFFFu;
			SSB_CUDA(cudaMemcpyAsync(ws.outCount.get() + q, &marker, 4, cudaMemcpyHostToDevice, s));
			SSB_CUDA(cudaStreamSynchronize(s));
			std::vector<std::pair<uint32_t, float>> lst(take);
			for (uint32_t i = 0; i < take; ++i) lst[i] = {hkid[i], hsc[i]};
			out->big.emplace_back(q, std::move(lst));
		} else {
			// device-only / resident callers (key-range shards, kernel timing) have no out-of-band channel: they get the
			// first `stride` results of the list (limit clamped to FAST_KLIM, include/ssb_engine.h) instead of none
			k_dense_emit<<<grid_for(stride), 256, 0, s>>>(sorted, stride, ws.outKid.get() + (size_t)q * stride, ws.outScore.get() + (size_t)q * stride);
			SSB_LAUNCH_CHECK();
			SSB_CUDA(cudaMemcpyAsync(ws.outCount.get() + q, &stride, 4, cudaMemcpyHostToDevice, s));
			SSB_CUDA(cudaStreamSynchronize(s));   // `stride` lives on this stack frame
		}
	} else {
		uint32_t zero = 0;
		SSB_CUDA(cudaMemcpyAsync(ws.outCount.get() + q, &zero, 4, cudaMemcpyHostToDevice, s));
		SSB_CUDA(cudaStreamSynchronize(s));
	}
}

void ensure_workspace(SearchWorkspace& ws, uint32_t nq, size_t charBytes, uint32_t stride, cudaStream_t s)
{
	if (!ws.ev0) {
		SSB_CUDA(cudaEventCreate(&ws.ev0)); SSB_CUDA(cudaEventCreate(&ws.ev1));
		SSB_CUDA(cudaEventCreate(&ws.ev2)); SSB_CUDA(cudaEventCreate(&ws.ev3));
		ws.counters.alloc(64, s);
		ws.hostCounters.reserve(64);
	}
	if (nq > ws.capQ || stride > ws.stride) {
		uint32_t cq = std::max(nq, ws.capQ), st = std::max(stride, ws.stride);
		ws.qoff.alloc(cq + 1, s); ws.desc.alloc(cq, s); ws.fastList.alloc(cq, s); ws.fastList2.alloc(cq, s); ws.denseList.alloc(cq, s);
		ws.outCount.alloc(cq, s); ws.outKid.alloc((size_t)cq * st, s); ws.outScore.alloc((size_t)cq * st, s);
		ws.csrCount.alloc(cq + 1, s); ws.csrOff.alloc(cq + 1, s); ws.csrKid.alloc((size_t)cq * st, s); ws.csrScore.alloc((size_t)cq * st, s);
		ws.csrRow.alloc((size_t)cq * st, s); ws.csrStr.alloc((size_t)cq * st, s);
		ws.capQ = cq; ws.stride = st;
	}
	if (charBytes > ws.capChars) {
		ws.capChars = charBytes + charBytes / 4 + 64;
		ws.qchars.alloc(ws.capChars, s); ws.qnorm.alloc(ws.capChars, s);
	}
}

template <class U>
void run_device(const DeviceIndex& ix, SearchWorkspace& ws, uint32_t nq, float threshold, uint32_t limit, uint32_t stride,
                BatchOutput* out, cudaStream_t s, SearchStats* st, HostResult* hr)
{
	const IndexView v = make_view(ix);
	// candidate buffer: 2 * limit entries rounded up to a power of two (>= 1024); it decides the CTA's shared memory
	// footprint and with it the number of resident CTAs per SM (6 x 128 threads at limit <= 512)
	uint32_t bufCap = 1024;
	while (bufCap < 2 * std::min<uint32_t>(limit, FAST_KLIM)) bufCap <<= 1;
	// indexes with descriptions / weights defer every survivor of the merge (edge and weight look-ups): they get a deferred
	// list of their own behind the candidate buffer instead of the small one that aliases the select histogram
	const uint32_t defExt = (ix.identity && ix.uniformW) ? 0u : (uint32_t)FS_DEF_EXT;
	const size_t smemBytes = ((sizeof(FastShared) + 15) & ~size_t(15)) + (size_t)bufCap * 8 + (size_t)defExt * 8;
	// per-device launch configuration, set up once and safely under concurrent searches on different indexes (only the index's
	// own mutex is held here): shared-memory opt-in of every kernel variant first, occupancy per configuration on demand
	struct LaunchCfg {
		std::once_flag once;
		std::mutex mu;
		int numSMs = 0;
		int ctasFor[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
		int sieveCtas[6] = {0, 0, 0, 0, 0, 0};
	};
	static LaunchCfg cfgOf[64];
	int dev = 0;
	SSB_CUDA(cudaGetDevice(&dev));
	LaunchCfg& lc = cfgOf[dev & 63];
	const bool withShort = ix.nShort > 0;
	const int bufIdx = bufCap == 1024 ? 0 : (bufCap == 2048 ? 1 : 2);
	const int cfg = bufIdx + (withShort ? 3 : 0) + (defExt ? 6 : 0);
	const size_t svBytes = ((sizeof(SieveShared) + 15) & ~size_t(15)) + (size_t)bufCap * 8 + (size_t)defExt * 8;
	std::call_once(lc.once, [&] {
		int n = 0;
		SSB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
		const int maxSmem = (int)(((sizeof(FastShared) + 15) & ~size_t(15)) + (size_t)FS_BUF_MAX * 8 + (size_t)FS_DEF_EXT * 8);
		SSB_CUDA(cudaFuncSetAttribute(k_fast_search<uint8_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
		SSB_CUDA(cudaFuncSetAttribute(k_fast_search<uint32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
		SSB_CUDA(cudaFuncSetAttribute(k_fast_search<uint8_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
		SSB_CUDA(cudaFuncSetAttribute(k_fast_search<uint32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
		const int maxSv = (int)(((sizeof(SieveShared) + 15) & ~size_t(15)) + (size_t)FS_BUF_MAX * 8 + (size_t)FS_DEF_EXT * 8);
		SSB_CUDA(cudaFuncSetAttribute(k_sieve_search<uint8_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSv));
		SSB_CUDA(cudaFuncSetAttribute(k_sieve_search<uint32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSv));
		SSB_CUDA(cudaFuncSetAttribute(k_sieve_search<uint8_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSv));
		SSB_CUDA(cudaFuncSetAttribute(k_sieve_search<uint32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSv));
		SSB_CUDA(cudaFuncSetAttribute(k_order_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ORDER_MAX * 8)));
		lc.numSMs = n;
	});
	const int numSMs = lc.numSMs;
	int ctasPerSM, sieveCtasPerSM;
	{
		std::lock_guard<std::mutex> lk(lc.mu);
		if (!lc.ctasFor[cfg]) {
			int c = 0;
			if (withShort) SSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, k_fast_search<U, true>, FS_THREADS, smemBytes));
			else SSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, k_fast_search<U, false>, FS_THREADS, smemBytes));
			lc.ctasFor[cfg] = std::max(c, 1);
			if (getenv("SSB_DEBUG"))
				fprintf(stderr, "[ssb] k_fast_search: sizeof(FastShared)=%zu, bufCap=%u, dynamic smem=%zu B, %d CTAs/SM x %d threads (short-term variant: %d)\n",
				        sizeof(FastShared), bufCap, smemBytes, lc.ctasFor[cfg], FS_THREADS, (int)withShort);
		}
		const int scfg = bufIdx + (defExt ? 3 : 0);
		if (!lc.sieveCtas[scfg]) {
			int c = 0;
			SSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, k_sieve_search<U, true>, SV_THREADS, svBytes));      // (both layouts: same bounds)
			lc.sieveCtas[scfg] = std::max(c, 1);
			if (getenv("SSB_DEBUG"))
				fprintf(stderr, "[ssb] k_sieve_search: sizeof(SieveShared)=%zu, dynamic smem=%zu B, %d CTAs/SM x %d threads\n", sizeof(SieveShared),
				        svBytes, lc.sieveCtas[scfg], SV_THREADS);
		}
		ctasPerSM = lc.ctasFor[cfg];
		sieveCtasPerSM = lc.sieveCtas[scfg];
	}
	unsigned long long* statCounters = reinterpret_cast<unsigned long long*>(ws.counters.get() + 8);
	// tiny batch with pinned result arrays: see k_finalize_tiny (no events: nobody profiles a single query through them)
	const bool tiny = hr && ix.K && nq <= 32 && hr->capQ >= nq && hr->capResults >= (size_t)nq * stride;
	if (!tiny) SSB_CUDA(cudaEventRecord(ws.ev0, s));
	ws.counters.zero(s);
	const U* qc = reinterpret_cast<const U*>(ws.qchars.get());
	U* qn = reinterpret_cast<U*>(ws.qnorm.get());
	// k_sieve_search pays a fixed price per query (set-up, pass plans, the general resolutions) and then grows slowly with the
	// lists; the register merge of k_fast_search costs nothing up front and grows with every posting. Measured on B200 (batch
	// 4096, ~9 lists per query, 32-byte sieve groups): a tie at ~19 000 postings per query (a 4M-key index), k_sieve_search
	// 3 % ahead at 23 000, 25 % at 46 000, 25 % behind at 11 500. The two kernels run one after the other, so the choice is made per batch, on the device, from the
	// mean posting volume of its fast queries (k_sieve_search hands the whole batch over when it is below the mark; queries
	// under a threshold count five-fold, see k_query_prep).
	// ssb_set_search_path(1) sends every eligible query to k_sieve_search (the parity tests cover it on small indexes that way).
	static const uint32_t sieveMinDefault = getenv("SSB_SIEVE_MIN_POSTINGS") ? (uint32_t)atoll(getenv("SSB_SIEVE_MIN_POSTINGS")) : 20000u;
	const uint32_t sieveMinTotal = g_search_path.load(std::memory_order_relaxed) == 1 ? 0u : sieveMinDefault;
	k_query_prep<U><<<grid_for((size_t)nq * 32), 256, 0, s>>>(v, qc, ws.qoff.get(), nq, threshold, limit, qn, ws.desc.get(),
	                                                          ws.fastList.get(), ws.denseList.get(), ws.counters.get(),
	                                                          ws.outCount.get());
	SSB_LAUNCH_CHECK();
	if (nq >= 256 && nq <= ORDER_MAX) {      // a handful of queries gain nothing from LPT order
		k_order_fast<<<1, 1024, ORDER_MAX * 8, s>>>(ws.desc.get(), ws.fastList.get(), ws.counters.get());
		SSB_LAUNCH_CHECK();
	}
	if (!tiny) SSB_CUDA(cudaEventRecord(ws.ev1, s));
	// persistent grids: one wave of resident CTAs, queries handed out by an atomic cursor. k_sieve_search goes through the fast
	// queries first and hands the ones it does not take (a list without bitmap, > 16 lists, edit-distance path) or cannot finish
	// to k_fast_search through fastList2. ssb_set_search_path(0) / SSB_MERGE=1: k_fast_search answers everything (the parity tests
	// run every batch both ways).
	static const int envPath = getenv("SSB_MERGE") ? 0 : 1;
	int pathOpt = g_search_path.load(std::memory_order_relaxed);
	if (pathOpt < 0) pathOpt = envPath;
	const bool useSieve = pathOpt == 1 && v.sieveNGroups != 0;
	// direct delivery (see HostOut): possible when the caller gave pinned CSR arrays, the stride is the limit itself and only
	// k_fast_search writes results; it holds if every query turns out to fill its `limit` slots (checked after the kernel)
	const bool direct = !tiny && hr && ix.K && hr->capQ >= nq && hr->capResults >= (size_t)nq * stride && stride == limit &&
	                    (size_t)nq * stride < (1ull << 32);
	HostOut ho{};
	if (direct) { ho.kids = hr->kids; ho.scores = hr->scores; ho.rows = hr->rows; ho.strs = reinterpret_cast<uint64_t*>(hr->strs); ho.strBase = hr->strBase; }
	if (useSieve) {
		const unsigned sgrid = (unsigned)std::min<size_t>((size_t)numSMs * sieveCtasPerSM, std::max<uint32_t>(nq, 1));
		// which layout of the bitmaps the batch streams (index.h): the 32-byte groups resolve a candidate block without touching
		// the postings but stream twice the bytes — they pay where candidate blocks are dense, i.e. while a query without (much
		// of) a threshold still looks for its k two-match terms: 10M keys, batch 4096, top-100 at threshold 0 / 0.3 / 0.5 (top-500):
		// 0.69 / 0.58 / 0.47 ms against 0.72 / 0.61 / 0.36 ms with the 16-byte groups
		static const float inlineMaxThreshold = getenv("SSB_SIEVE_INLINE_MAX_THRESHOLD") ? (float)atof(getenv("SSB_SIEVE_INLINE_MAX_THRESHOLD")) : 0.3f;
		const bool inl = v.sieveInline != nullptr && !(threshold > inlineMaxThreshold);
		if (inl)
			k_sieve_search<U, true><<<sgrid, SV_THREADS, svBytes, s>>>(v, ws.desc.get(), qn, ws.fastList.get(), ws.counters.get(), ws.fastList2.get(),
			                                                          limit, stride, ws.outCount.get(), ws.outKid.get(), ws.outScore.get(),
			                                                          statCounters, bufCap, threshold, defExt, sieveMinTotal, ho);
		else
			k_sieve_search<U, false><<<sgrid, SV_THREADS, svBytes, s>>>(v, ws.desc.get(), qn, ws.fastList.get(), ws.counters.get(), ws.fastList2.get(),
			                                                           limit, stride, ws.outCount.get(), ws.outKid.get(), ws.outScore.get(),
			                                                           statCounters, bufCap, threshold, defExt, sieveMinTotal, ho);
		SSB_LAUNCH_CHECK();
	}
	const uint32_t* fl = useSieve ? ws.fastList2.get() : ws.fastList.get();
	const uint32_t nIdx = useSieve ? 5u : 0u, curIdx = useSieve ? 12u : 2u;
	const unsigned grid = (unsigned)std::min<size_t>((size_t)numSMs * ctasPerSM, std::max<uint32_t>(nq, 1));
	if (withShort)
		k_fast_search<U, true><<<grid, FS_THREADS, smemBytes, s>>>(v, ws.desc.get(), qn, fl, ws.counters.get(),
		                                                           ws.denseList.get(), limit, stride, ws.outCount.get(), ws.outKid.get(),
		                                                           ws.outScore.get(), statCounters, bufCap, threshold, nIdx, curIdx, defExt, ho);
	else
		k_fast_search<U, false><<<grid, FS_THREADS, smemBytes, s>>>(v, ws.desc.get(), qn, fl, ws.counters.get(),
		                                                            ws.denseList.get(), limit, stride, ws.outCount.get(), ws.outKid.get(),
		                                                            ws.outScore.get(), statCounters, bufCap, threshold, nIdx, curIdx, defExt, ho);
	SSB_LAUNCH_CHECK();
	if (!tiny) SSB_CUDA(cudaEventRecord(ws.ev2, s));

	// CSR finalisation (compaction + gather of row ids / string addresses) is enqueued speculatively, before the host
	// knows whether any query needs the dense path: the common case costs ONE host synchronisation for small batches
	// (results copied at capacity together with the counters) and two for large ones (exact-size copy of the arrays).
	const bool csr = hr && ix.K && hr->capQ >= nq && hr->capResults >= (size_t)nq * stride;
	const bool small = (size_t)nq * stride <= 8192;
	auto copy_arrays = [&](size_t total) {
		SSB_CUDA(cudaMemcpyAsync(hr->kids, ws.csrKid.get(), total * 4, cudaMemcpyDeviceToHost, s));
		SSB_CUDA(cudaMemcpyAsync(hr->scores, ws.csrScore.get(), total * 4, cudaMemcpyDeviceToHost, s));
		SSB_CUDA(cudaMemcpyAsync(hr->rows, ws.csrRow.get(), total * 8, cudaMemcpyDeviceToHost, s));
		SSB_CUDA(cudaMemcpyAsync(hr->strs, ws.csrStr.get(), total * 8, cudaMemcpyDeviceToHost, s));
	};
	auto enqueue_finalize = [&] {
		k_clamp_counts<<<grid_for(nq), 256, 0, s>>>(ws.outCount.get(), nq, stride, ws.csrCount.get(), ws.counters.get() + 4);
		SSB_LAUNCH_CHECK();
		exclusive_scan_u32_u32(ws.csrCount.get(), ws.csrOff.get(), nq, ws.csrOff.get() + nq, s);
		k_finalize_csr<<<grid_for((size_t)nq * 32), 256, 0, s>>>(ws.csrCount.get(), ws.csrOff.get(), ws.outKid.get(), ws.outScore.get(), nq,
		                                                       stride, ix.keyOff.get(), ix.keyRow.get(), hr->strBase, (uint32_t)ix.unit,
		                                                       ws.csrKid.get(), ws.csrScore.get(), ws.csrRow.get(), ws.csrStr.get());
		SSB_LAUNCH_CHECK();
		SSB_CUDA(cudaMemcpyAsync(hr->offsets, ws.csrOff.get(), (size_t)(nq + 1) * 4, cudaMemcpyDeviceToHost, s));
		if (small) copy_arrays((size_t)nq * stride);
	};
	uint32_t* hc = ws.hostCounters.get();
	if (tiny) {
		k_finalize_tiny<<<1, 256, 0, s>>>(ws.outCount.get(), ws.outKid.get(), ws.outScore.get(), nq, stride, ix.keyOff.get(), ix.keyRow.get(),
		                                  hr->strBase, (uint32_t)ix.unit, ws.counters.get(), hr->offsets, hr->kids, hr->scores, hr->rows,
		                                  hr->strs, hc);
		SSB_LAUNCH_CHECK();
		SSB_CUDA(cudaStreamSynchronize(s));
	} else {
		if (csr && !direct) enqueue_finalize();
		SSB_CUDA(cudaMemcpyAsync(hc, ws.counters.get(), 64 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
		SSB_CUDA(cudaEventRecord(ws.ev3, s));
		SSB_CUDA(cudaEventSynchronize(ws.ev3));
	}
	bool delivered = false;
	if (direct) {
		if (hc[0] == nq && hc[1] == 0 && hc[22] == 0) {
			// every query is a fast query with exactly `limit` results: the stores of the kernel ARE the CSR arrays
			for (uint32_t q = 0; q <= nq; ++q) hr->offsets[q] = q * stride;
			delivered = true;
		} else if (hc[1] == 0) {
			enqueue_finalize();                       // some list is short: compact on the device as usual
			SSB_CUDA(cudaStreamSynchronize(s));
		}
	}
	const uint32_t nFast = hc[0], nDense = hc[1], nOverflow = hc[3];
	unsigned long long postings;
	memcpy(&postings, hc + 8, 8);
	if (nDense) {
		std::vector<uint32_t> dl(nDense);
		SSB_CUDA(cudaMemcpyAsync(dl.data(), ws.denseList.get(), (size_t)nDense * 4, cudaMemcpyDeviceToHost, s));
		SSB_CUDA(cudaStreamSynchronize(s));
		std::sort(dl.begin(), dl.end());
		for (uint32_t q : dl) {
			QDesc d;
			SSB_CUDA(cudaMemcpyAsync(&d, ws.desc.get() + q, sizeof(QDesc), cudaMemcpyDeviceToHost, s));
			SSB_CUDA(cudaStreamSynchronize(s));
			run_dense_query<U>(ix, v, ws, d, q, threshold, limit, stride, out, s);
		}
		if (csr && out && out->big.empty()) {          // the dense queries changed the fixed-stride rows: finalise again
			enqueue_finalize();
			SSB_CUDA(cudaStreamSynchronize(s));
		}
	}
	if (csr && out && out->big.empty()) {
		if (!small && !delivered) {
			const size_t total = hr->offsets[nq];
			if (total) { copy_arrays(total); SSB_CUDA(cudaStreamSynchronize(s)); }
		}
		out->csr = true;
	}
	if (st) {
		st->nFast = nFast; st->nDense = nDense; st->nOverflow = nOverflow;
		st->nSieve = hc[6]; st->nSieveHandover = hc[5]; st->nSieveGiveUp = hc[28]; st->sieveProbes = hc[25];
		static const bool dbg = getenv("SSB_DEBUG") != nullptr;
		if (dbg && nFast)
			fprintf(stderr, "[ssb] k_fast_search per query: %.1f passes, %.2f redone, %.1f selects, %.1f deferred evaluations\n", hc[16] / (double)nFast,
			        hc[17] / (double)nFast, hc[18] / (double)nFast, hc[19] / (double)nFast);
		if (dbg && hc[6])
			fprintf(stderr, "[ssb] k_sieve_search per query: %.1f passes (%.2f redone), %.0f candidate blocks, %.0f probes; gave up on %u\n",
			        hc[24] / (double)hc[6], hc[27] / (double)hc[6], hc[26] / (double)hc[6], hc[25] / (double)hc[6], hc[28]);
		if (dbg && hc[6] && hc[34])
			fprintf(stderr, "[ssb] k_sieve_search kilocycles per query: set-up %.1f, plan %.1f, sieve+resolve %.1f, slow+deferred %.1f, select %.1f, final %.1f\n",
			        hc[32] * 0.064 / hc[6], hc[33] * 0.064 / hc[6], hc[34] * 0.064 / hc[6], hc[35] * 0.064 / hc[6], hc[36] * 0.064 / hc[6], hc[37] * 0.064 / hc[6]),
			fprintf(stderr, "[ssb]   slowest query %u kilocycles; queries over 0.5 / 1 / 2 megacycles: %u / %u / %u\n", hc[43], hc[46], hc[44], hc[45]),
			fprintf(stderr, "[ssb]   CTAs by the time they ran out of work (steps of 65.5 us): %u %u %u %u %u %u %u %u %u %u %u %u\n", hc[48], hc[49], hc[50], hc[51],
			        hc[52], hc[53], hc[54], hc[55], hc[56], hc[57], hc[58], hc[59]),
			fprintf(stderr, "[ssb]   thread 0's general resolutions: list words %.1f, groups + postings %.1f, counting %.1f kilocycles; %.1f lists hit\n", hc[60] * 0.064 / hc[6], hc[61] * 0.064 / hc[6], hc[62] * 0.064 / hc[6], hc[63] / (double)hc[6]),
			fprintf(stderr, "[ssb]   of slow+deferred: thread 0's own general resolutions %.1f, waiting for the CTA's %.1f\n", hc[38] * 0.064 / hc[6], hc[39] * 0.064 / hc[6]),
			fprintf(stderr, "[ssb]   of sieve+resolve, warp 0: loads + counters %.1f, queue pushes %.1f, resolve rounds %.1f\n", hc[40] * 0.064 / hc[6],
			        hc[41] * 0.064 / hc[6], hc[42] * 0.064 / hc[6]);
		st->postingsTouched = postings;
		st->msDevice = st->msFastKernel = 0;
		if (!tiny) {
			SSB_CUDA(cudaEventElapsedTime(&st->msDevice, ws.ev0, ws.ev3));
			SSB_CUDA(cudaEventElapsedTime(&st->msFastKernel, ws.ev1, ws.ev2));
		}
	}
}

}   // namespace

void search_batch(const DeviceIndex& ix, SearchWorkspace& ws, const void* qchars, const uint32_t* qoff, uint32_t nq, float threshold,
                  uint32_t limit, BatchOutput& out, cudaStream_t s, SearchStats* stats, HostResult* hr, bool deviceOnly)
{
	out = BatchOutput();
	out.nq = nq;
	if (limit == 0) limit = 0x7FFFFFFFu;                                               // hpp:420-421,454-455
	limit = std::min<uint32_t>(limit, std::max<uint32_t>(ix.K, 1u));                   // a result list never holds more than K keys: keeps
	                                                                                   // "unlimited" queries on small indexes on the fast path
	const uint32_t stride = std::min<uint32_t>(limit, FAST_KLIM);
	out.stride = stride;
	out.count.assign(nq, 0);
	if (!ix.indexed || nq == 0) return;                                                // hpp:417,451
	const size_t charBytes = (size_t)qoff[nq] * ix.unit;
	ensure_workspace(ws, nq, charBytes, stride, s);
	if (charBytes) SSB_CUDA(cudaMemcpyAsync(ws.qchars.get(), qchars, charBytes, cudaMemcpyHostToDevice, s));
	SSB_CUDA(cudaMemcpyAsync(ws.qoff.get(), qoff, (size_t)(nq + 1) * 4, cudaMemcpyHostToDevice, s));
	HostResult* use = deviceOnly ? nullptr : hr;
	if (ix.unit == 1) run_device<uint8_t>(ix, ws, nq, threshold, limit, stride, &out, s, stats, use);
	else run_device<uint32_t>(ix, ws, nq, threshold, limit, stride, &out, s, stats, use);
	if (deviceOnly || out.csr) return;          // sharded path: results stay in HBM; CSR path: already delivered into hr
	out.kid.resize((size_t)nq * stride);
	out.score.resize((size_t)nq * stride);
	SSB_CUDA(cudaMemcpyAsync(out.count.data(), ws.outCount.get(), (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
	SSB_CUDA(cudaMemcpyAsync(out.kid.data(), ws.outKid.get(), (size_t)nq * stride * 4, cudaMemcpyDeviceToHost, s));
	SSB_CUDA(cudaMemcpyAsync(out.score.data(), ws.outScore.get(), (size_t)nq * stride * 4, cudaMemcpyDeviceToHost, s));
	SSB_CUDA(cudaStreamSynchronize(s));
	for (auto& b : out.big) out.count[b.first] = (uint32_t)b.second.size();
}

void search_batch_resident(const DeviceIndex& ix, SearchWorkspace& ws, uint32_t nq, float threshold, uint32_t limit, cudaStream_t s,
                           SearchStats* stats)
{
	if (limit == 0) limit = 0x7FFFFFFFu;
	limit = std::min<uint32_t>(limit, std::max<uint32_t>(ix.K, 1u));
	if (!ix.indexed || nq == 0 || nq > ws.capQ) throw std::runtime_error("search_batch_resident: upload a batch with search_batch first");
	const uint32_t stride = std::min<uint32_t>(limit, FAST_KLIM);
	if (stride > ws.stride) throw std::runtime_error("search_batch_resident: limit larger than the uploaded batch's");
	if (ix.unit == 1) run_device<uint8_t>(ix, ws, nq, threshold, limit, stride, nullptr, s, stats, nullptr);
	else run_device<uint32_t>(ix, ws, nq, threshold, limit, stride, nullptr, s, stats, nullptr);
}

}   // namespace ssb
