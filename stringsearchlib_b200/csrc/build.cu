// build.cu — index build on the GPU (replaces StringIndex::StringIndex / init / buildGrams,
// nGramSearch.hpp:120-172, 54-108, 13-21, 41-46).
//
// Pipeline (every step a hand-written sm_100a kernel or the radix sort / scan of primitives.cu):
//   K-norm    k_word_spans      trim / escapeBlank / toUpper spans + 64-bit hashes per word (nGramSearch.h:30-98)
//   K-dedupe  hash sort + verify  string interning of keys and terms (unordered_set, hpp:58-74)
//   K-edges   (term,kid) sort, last-writer-wins weights (hpp:146-147,164-165), CSR both ways
//   K-grams   k_emit_grams      gram windows of long terms -> symbol-packed codes (nGramSearch.h:147-150)
//   K-sort    radix_sort_u64    by gram code (stable: term ids stay ascending)
//   K-rle     k_rle_*           unique (code, term) + run-length -> CSR postings (hpp:19)
#include "index.h"
#include "primitives.cuh"
#include "textops.cuh"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <thread>
#include <vector>

namespace ssb {
namespace {

constexpr int TPB = 256;
inline unsigned grid_for(size_t n, int tpb = TPB) { return (unsigned)std::max<size_t>(1, (n + tpb - 1) / tpb); }

inline int bits_for(uint64_t count)   // bits needed to represent values in [0, count)
{
	int b = 1;
	while (b < 64 && (1ull << b) < count) ++b;
	return b;
}

struct Props {
	unsigned int notIdentity, notUniform, multiEdge, collision;
	unsigned int wmaxOrdered;   // float_to_ordered(max weight)
	unsigned int maxTermLen, maxKeyLen, pad;
};

// ---------------------------------------------------------------------------------------------- K-norm
template <class U>
__global__ void __launch_bounds__(TPB) k_word_spans(const U* __restrict__ chars, const uint64_t* __restrict__ off,
                                                    const uint8_t* __restrict__ isNull, const float* __restrict__ weight,
                                                    uint32_t n, uint32_t rowSize, const uint8_t* __restrict__ lut, uint64_t seed,
                                                    uint64_t* __restrict__ termBeg, uint32_t* __restrict__ termLen,
                                                    uint64_t* __restrict__ termHash, uint8_t* __restrict__ edgeValid,
                                                    uint64_t* __restrict__ rowKeyBeg, uint32_t* __restrict__ rowKeyLen,
                                                    uint64_t* __restrict__ rowKeyHash, uint8_t* __restrict__ rowValid)
{
	uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= n) return;
	const uint32_t r = w / rowSize, kw = r * rowSize;
	const bool isKey = (w == kw);
	// the row is valid iff its key cell is non-null and non-empty after trim (hpp:129-135)
	bool rowOk = false;
	uint64_t kb = 0, ke = 0;
	if (!isNull[kw]) {
		kb = off[kw]; ke = off[kw + 1];
		while (kb < ke && unit_space((uint32_t)chars[kb])) ++kb;
		rowOk = kb < ke;
		if (isKey && rowOk) while (ke > kb && unit_space((uint32_t)chars[ke - 1])) --ke;
	}
	if (isKey) {
		uint64_t h = seed ^ 0x6b657973ull;
		for (uint64_t i = kb; i < ke && rowOk; ++i) h = hash_step(h, (uint32_t)chars[i]);
		rowKeyBeg[r] = kb;
		rowKeyLen[r] = rowOk ? (uint32_t)(ke - kb) : 0;
		rowKeyHash[r] = hash_done(h, (uint32_t)(ke - kb));
		rowValid[r] = rowOk;
	}
	// search term of this cell: toUpper(trim(escapeBlank(raw))) (hpp:136-139,153-156)
	uint64_t tb = 0, te = 0;
	bool ok = rowOk && !isNull[w];
	if (ok) {
		tb = off[w]; te = off[w + 1];
		while (tb < te && unit_blank(chars[tb], lut)) ++tb;
		while (te > tb && unit_blank(chars[te - 1], lut)) --te;
		if (!isKey && tb == te) ok = false;                 // empty description cell (hpp:157); an empty KEY term is kept (hpp:146)
		float wt = weight ? weight[w] : 1.0f;
		if (wt == 0.0f) ok = false;                         // hpp:144,162
	}
	uint64_t h = seed;
	for (uint64_t i = tb; i < te && ok; ++i) h = hash_step(h, (uint32_t)unit_norm(chars[i], lut));
	termBeg[w] = tb;
	termLen[w] = ok ? (uint32_t)(te - tb) : 0;
	termHash[w] = hash_done(h, (uint32_t)(te - tb));
	edgeValid[w] = ok;
}

// ---------------------------------------------------------------------------------------------- generic helpers
__global__ void k_compact_ids(const uint8_t* __restrict__ flag, const uint32_t* __restrict__ pos, const uint64_t* __restrict__ hash,
                              uint32_t n, uint64_t* __restrict__ outKey, uint32_t* __restrict__ outId)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && flag[i]) { outKey[pos[i]] = hash[i]; outId[pos[i]] = i; }
}
__global__ void k_flag_new(const uint64_t* __restrict__ keys, uint32_t n, uint8_t* __restrict__ flag)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) flag[i] = (i == 0 || keys[i] != keys[i - 1]);
}
__global__ void k_set_leaders(const uint8_t* __restrict__ flag, const uint32_t* __restrict__ grpExcl, const uint32_t* __restrict__ ids,
                              uint32_t n, uint32_t* __restrict__ leaderOfGrp)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && flag[i]) leaderOfGrp[grpExcl[i]] = ids[i];
}

// keys: raw trimmed strings must be byte-equal inside a hash group (interning by exact equality, hpp:58-65)
template <class U>
__global__ void k_verify_keys(const U* __restrict__ chars, const uint8_t* __restrict__ flag, const uint32_t* __restrict__ grpExcl,
                              const uint32_t* __restrict__ ids, const uint32_t* __restrict__ leaderOfGrp, uint32_t n,
                              const uint64_t* __restrict__ rowKeyBeg, const uint32_t* __restrict__ rowKeyLen,
                              uint32_t* __restrict__ rowLeader, uint8_t* __restrict__ isLeaderRow, Props* props)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t r = ids[i];
	uint32_t lead = leaderOfGrp[grpExcl[i] + flag[i] - 1];
	rowLeader[r] = lead;
	if (lead == r) { isLeaderRow[r] = 1; return; }
	uint32_t len = rowKeyLen[r];
	bool same = len == rowKeyLen[lead];
	const U* a = chars + rowKeyBeg[r];
	const U* b = chars + rowKeyBeg[lead];
	for (uint32_t j = 0; j < len && same; ++j) same = a[j] == b[j];
	if (!same) atomicOr(&props->collision, 1u);
}
template <class U>
__global__ void k_verify_terms(const U* __restrict__ chars, const uint8_t* __restrict__ lut, const uint8_t* __restrict__ flag,
                               const uint32_t* __restrict__ grpExcl, const uint32_t* __restrict__ ids,
                               const uint32_t* __restrict__ leaderOfGrp, uint32_t n, const uint64_t* __restrict__ termBeg,
                               const uint32_t* __restrict__ termLen, uint32_t* __restrict__ wordLeader,
                               uint8_t* __restrict__ isLeaderWord, Props* props)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t w = ids[i];
	uint32_t lead = leaderOfGrp[grpExcl[i] + flag[i] - 1];
	wordLeader[w] = lead;
	if (lead == w) { isLeaderWord[w] = 1; return; }
	uint32_t len = termLen[w];
	bool same = len == termLen[lead];
	const U* a = chars + termBeg[w];
	const U* b = chars + termBeg[lead];
	for (uint32_t j = 0; j < len && same; ++j) same = unit_norm(a[j], lut) == unit_norm(b[j], lut);
	if (!same) atomicOr(&props->collision, 1u);
}

// ---------------------------------------------------------------------------------------------- keys
__global__ void k_ext_info(const uint8_t* __restrict__ isLeaderRow, const uint32_t* __restrict__ extOfRow,
                           const uint32_t* __restrict__ rowKeyLen, uint32_t R, uint64_t* __restrict__ sortKey,
                           uint32_t* __restrict__ sortVal, uint32_t* __restrict__ extRow, Props* props)
{
	uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= R || !isLeaderRow[r]) return;
	uint32_t e = extOfRow[r];
	sortKey[e] = rowKeyLen[r];
	sortVal[e] = e;
	extRow[e] = r;
	atomicMax(&props->maxKeyLen, rowKeyLen[r]);
}
// sorted position k (by length, stable) is the kid of ext id sortedExt[k]
__global__ void k_kid_assign(const uint32_t* __restrict__ sortedExt, const uint32_t* __restrict__ extRow, uint32_t K,
                             uint32_t* __restrict__ kidOfExt, uint32_t* __restrict__ kidRowLocal)
{
	uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= K) return;
	uint32_t e = sortedExt[k];
	kidOfExt[e] = k;
	kidRowLocal[k] = extRow[e];
}
__global__ void k_row_kid(const uint8_t* __restrict__ rowValid, const uint32_t* __restrict__ rowLeader, const uint32_t* __restrict__ extOfRow,
                          const uint32_t* __restrict__ kidOfExt, uint32_t R, uint32_t* __restrict__ rowKid)
{
	uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r < R) rowKid[r] = rowValid[r] ? kidOfExt[extOfRow[rowLeader[r]]] : 0xFFFFFFFFu;
}

// ---------------------------------------------------------------------------------------------- terms
__global__ void k_min_kid(const uint8_t* __restrict__ edgeValid, const uint32_t* __restrict__ wordLeader, const uint32_t* __restrict__ tmpOfWord,
                          const uint32_t* __restrict__ rowKid, uint32_t n, uint32_t rowSize, uint32_t* __restrict__ minKid,
                          uint32_t* __restrict__ tmpLeaderWord)
{
	uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= n || !edgeValid[w]) return;
	uint32_t lead = wordLeader[w];
	uint32_t tmp = tmpOfWord[lead];
	atomicMin(&minKid[tmp], rowKid[w / rowSize]);
	if (lead == w) tmpLeaderWord[tmp] = w;
}
__global__ void k_term_sort_keys(const uint32_t* __restrict__ minKid, uint32_t T, uint64_t* __restrict__ key, uint32_t* __restrict__ val)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t < T) { key[t] = minKid[t]; val[t] = t; }
}
__global__ void k_term_assign(const uint32_t* __restrict__ sortedTmp, const uint32_t* __restrict__ tmpLeaderWord,
                              const uint32_t* __restrict__ wordTermLen, uint32_t T, uint32_t twoG, uint32_t* __restrict__ termOfTmp,
                              uint32_t* __restrict__ termWord, uint32_t* __restrict__ termLen, uint8_t* __restrict__ isShort, Props* props)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= T) return;
	uint32_t tmp = sortedTmp[t];
	termOfTmp[tmp] = t;
	uint32_t w = tmpLeaderWord[tmp];
	termWord[t] = w;
	uint32_t len = wordTermLen[w];
	termLen[t] = len;
	isShort[t] = len < twoG;                                 // hpp:82 (literal 6 == 2*gSize)
	atomicMax(&props->maxTermLen, len);
}

// ---------------------------------------------------------------------------------------------- edges
__global__ void k_edge_keys(const uint8_t* __restrict__ edgeValid, const uint32_t* __restrict__ pos, const uint32_t* __restrict__ wordLeader,
                            const uint32_t* __restrict__ tmpOfWord, const uint32_t* __restrict__ termOfTmp,
                            const uint32_t* __restrict__ rowKid, uint32_t n, uint32_t rowSize, uint64_t* __restrict__ key,
                            uint32_t* __restrict__ val)
{
	uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= n || !edgeValid[w]) return;
	uint32_t t = termOfTmp[tmpOfWord[wordLeader[w]]];
	key[pos[w]] = ((uint64_t)t << 32) | rowKid[w / rowSize];
	val[pos[w]] = w;
}
__global__ void k_flag_last(const uint64_t* __restrict__ keys, uint32_t n, uint8_t* __restrict__ flag)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) flag[i] = (i + 1 == n || keys[i] != keys[i + 1]);   // the LAST occurrence's weight wins (hpp:147,165)
}
__global__ void k_compact_edges(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ words, const uint8_t* __restrict__ flag,
                                const uint32_t* __restrict__ pos, const float* __restrict__ weight, uint32_t n,
                                uint32_t* __restrict__ edgeTerm, uint32_t* __restrict__ edgeKid, float* __restrict__ edgeW, Props* props)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || !flag[i]) return;
	uint32_t e = pos[i];
	uint32_t t = (uint32_t)(keys[i] >> 32), kid = (uint32_t)keys[i];
	float w = weight ? weight[words[i]] : 1.0f;
	edgeTerm[e] = t; edgeKid[e] = kid; edgeW[e] = w;
	if (t != e || kid != e) props->notIdentity = 1;
	if (w != 1.0f) props->notUniform = 1;
	atomicMax(&props->wmaxOrdered, float_to_ordered(w));
}
// CSR offsets from a sorted id column; ids may skip values (keys without edges)
__global__ void k_csr_offsets(const uint32_t* __restrict__ sortedId, uint32_t n, uint32_t nIds, uint32_t* __restrict__ off)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > n) return;
	uint32_t lo = (i == 0) ? 0 : sortedId[i - 1] + 1;
	uint32_t hi = (i == n) ? nIds + 1 : sortedId[i] + 1;      // off[lo..hi) = i
	for (uint32_t k = lo; k < hi; ++k) off[k] = i;
}
__global__ void k_rev_keys(const uint32_t* __restrict__ edgeTerm, const uint32_t* __restrict__ edgeKid, uint32_t E,
                           uint64_t* __restrict__ key, uint32_t* __restrict__ val)
{
	uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e < E) { key[e] = ((uint64_t)edgeKid[e] << 32) | edgeTerm[e]; val[e] = e; }
}
__global__ void k_rev_fill(const uint64_t* __restrict__ key, const uint32_t* __restrict__ val, const float* __restrict__ edgeW, uint32_t E,
                           uint32_t* __restrict__ revKid, uint32_t* __restrict__ keyEdgeTerm, float* __restrict__ keyEdgeW)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= E) return;
	revKid[i] = (uint32_t)(key[i] >> 32);
	keyEdgeTerm[i] = (uint32_t)key[i];
	keyEdgeW[i] = edgeW[val[i]];
}
__global__ void k_key_wmax(const uint32_t* __restrict__ keyEdgeOff, const float* __restrict__ keyEdgeW, uint32_t K,
                           float* __restrict__ keyWmax, Props* props)
{
	uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= K) return;
	uint32_t b = keyEdgeOff[k], e = keyEdgeOff[k + 1];
	float m = -INFINITY;
	for (uint32_t i = b; i < e; ++i) m = fmaxf(m, keyEdgeW[i]);
	keyWmax[k] = m;
	if (e - b > 1) props->multiEdge = 1;
}

// ---------------------------------------------------------------------------------------------- strings
__global__ void k_len_plus1(const uint32_t* __restrict__ kidRowLocal, const uint32_t* __restrict__ rowKeyLen, uint32_t K,
                            uint32_t* __restrict__ out)
{
	uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < K) out[k] = rowKeyLen[kidRowLocal[k]] + 1;
}
template <class U>
__global__ void k_copy_terms(const U* __restrict__ chars, const uint8_t* __restrict__ lut, const uint32_t* __restrict__ termWord,
                             const uint64_t* __restrict__ wordTermBeg, const uint32_t* __restrict__ termOff, uint32_t T,
                             U* __restrict__ out)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= T) return;
	const U* src = chars + wordTermBeg[termWord[t]];
	uint32_t b = termOff[t], len = termOff[t + 1] - b;
	for (uint32_t j = 0; j < len; ++j) out[b + j] = unit_norm(src[j], lut);
}
template <class U>
__global__ void k_copy_keys(const U* __restrict__ chars, const uint32_t* __restrict__ kidRowLocal, const uint64_t* __restrict__ rowKeyBeg,
                            const uint32_t* __restrict__ keyOff, uint32_t K, uint64_t rowBase, U* __restrict__ out,
                            uint64_t* __restrict__ keyRow, uint64_t* __restrict__ keyHash)
{
	uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= K) return;
	uint32_t r = kidRowLocal[k];
	const U* src = chars + rowKeyBeg[r];
	uint32_t b = keyOff[k], len = keyOff[k + 1] - b - 1;
	uint64_t h = 0x6b657973ull;                         // fixed seed: equal key strings hash equally in every shard
	for (uint32_t j = 0; j < len; ++j) { out[b + j] = src[j]; h = hash_step(h, (uint32_t)src[j]); }
	out[b + len] = 0;
	keyRow[k] = rowBase + r;
	keyHash[k] = hash_done(h, len);
}
__global__ void k_compact_short(const uint8_t* __restrict__ isShort, const uint32_t* __restrict__ pos, uint32_t T, uint32_t* __restrict__ out)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t < T && isShort[t]) out[pos[t]] = t;
}

// ---------------------------------------------------------------------------------------------- grams
// narrow alphabet: presence of every byte value inside long terms
__global__ void k_presence_narrow(const uint8_t* __restrict__ termChars, const uint32_t* __restrict__ termOff, uint32_t T, uint32_t twoG,
                                  uint32_t* __restrict__ present /*256*/)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= T) return;
	uint32_t b = termOff[t], e = termOff[t + 1];
	if (e - b < twoG) return;
	for (uint32_t j = b; j < e; ++j) present[termChars[j]] = 1;
}
__global__ void k_build_lut_narrow(const uint32_t* __restrict__ present, uint32_t* __restrict__ lut, uint32_t* __restrict__ A)
{
	if (threadIdx.x == 0) {
		uint32_t a = 0;
		for (int u = 0; u < 256; ++u) lut[u] = present[u] ? ++a : 0;
		*A = a;
	}
}
// wide alphabet: all units of long terms -> sort -> unique
__global__ void k_units_wide(const uint32_t* __restrict__ termChars, const uint32_t* __restrict__ termOff, uint32_t T, uint32_t twoG,
                             uint64_t* __restrict__ out, uint32_t nUnits)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= T) return;
	uint32_t b = termOff[t], e = termOff[t + 1];
	bool isLong = e - b >= twoG;
	for (uint32_t j = b; j < e; ++j) out[j] = isLong ? (uint64_t)termChars[j] : 0xFFFFFFFFFFull;   // short-term units sort last
}
__global__ void k_compact_alpha(const uint64_t* __restrict__ sorted, const uint8_t* __restrict__ flag, const uint32_t* __restrict__ pos,
                                uint32_t n, uint32_t* __restrict__ alphabet)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && flag[i] && sorted[i] != 0xFFFFFFFFFFull) alphabet[pos[i]] = (uint32_t)sorted[i];
}

__global__ void k_window_counts(const uint32_t* __restrict__ termOff, uint32_t T, uint32_t g, uint32_t* __restrict__ cnt)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= T) return;
	uint32_t len = termOff[t + 1] - termOff[t];
	cnt[t] = len >= 2 * g ? len - g + 1 : 0;                  // windows i = 0..len-g of long terms (hpp:16,82)
}
// one warp per 32 consecutive terms would be wasteful for short strings; one thread per term writes a contiguous run
template <class U, bool PACKED>
__global__ void __launch_bounds__(TPB) k_emit_grams(const U* __restrict__ termChars, const uint32_t* __restrict__ termOff,
                                                   const uint32_t* __restrict__ winOff, uint32_t T, uint32_t g, uint32_t symBits,
                                                   const uint32_t* __restrict__ symTable, uint32_t A, uint64_t* __restrict__ keys,
                                                   uint32_t* __restrict__ vals)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= T) return;
	uint32_t b = termOff[t], len = termOff[t + 1] - b;
	if (len < 2 * g) return;
	uint32_t o = winOff[t];
	const uint64_t mask = (g * symBits >= 64) ? ~0ull : ((1ull << (g * symBits)) - 1);
	uint64_t code = 0;
	for (uint32_t j = 0; j < len; ++j) {
		code = ((code << symBits) | sym_of(termChars[b + j], symTable, A)) & mask;
		if (j + 1 >= g) {
			if (PACKED) keys[o] = (code << 32) | t;
			else { keys[o] = code; vals[o] = t; }
			++o;
		}
	}
}
template <bool PACKED>
__global__ void k_rle_flags(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t n, uint8_t* __restrict__ newPair,
                            uint8_t* __restrict__ newGram)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	bool np, ng;
	if (i == 0) np = ng = true;
	else if (PACKED) { np = keys[i] != keys[i - 1]; ng = (keys[i] >> 32) != (keys[i - 1] >> 32); }
	else { ng = keys[i] != keys[i - 1]; np = ng || vals[i] != vals[i - 1]; }
	newPair[i] = np;
	newGram[i] = ng;
}
template <bool PACKED>
__global__ void k_rle_write(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint8_t* __restrict__ newPair,
                            const uint8_t* __restrict__ newGram, const uint32_t* __restrict__ pairPos, const uint32_t* __restrict__ gramPos,
                            uint32_t n, uint32_t* __restrict__ postings, uint64_t* __restrict__ gramCodes, uint32_t* __restrict__ gramOff)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (newPair[i]) postings[pairPos[i]] = PACKED ? (uint32_t)keys[i] : vals[i];
	if (newGram[i]) {
		gramCodes[gramPos[i]] = PACKED ? (keys[i] >> 32) : keys[i];
		gramOff[gramPos[i]] = pairPos[i];
	}
}
__global__ void k_set_u32(uint32_t* p, uint32_t v) { *p = v; }

struct Timer {
	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	double lap()
	{
		auto t1 = std::chrono::steady_clock::now();
		double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
		t0 = t1;
		return ms;
	}
};

template <class T> T d2h_scalar(const T* d, cudaStream_t s)
{
	T v;
	SSB_CUDA(cudaMemcpyAsync(&v, d, sizeof(T), cudaMemcpyDeviceToHost, s));
	SSB_CUDA(cudaStreamSynchronize(s));
	return v;
}

// sorts (keys, vals) and returns pointers to the sorted buffers
struct Sorted { uint64_t* keys; uint32_t* vals; };
Sorted sort_pairs(DevBuf<uint64_t>& k, DevBuf<uint64_t>& kAlt, DevBuf<uint32_t>& v, DevBuf<uint32_t>& vAlt, size_t n, int lo, int hi,
                  cudaStream_t s)
{
	bool alt = radix_sort_u64(k.get(), kAlt.get(), v.get(), vAlt.get(), n, lo, hi, s);
	return alt ? Sorted{kAlt.get(), vAlt.get()} : Sorted{k.get(), v.get()};
}

// The host copy of the key table travels while the device goes on with edges, grams and bitmaps: a second stream copies
// keyChars into the pinned staging buffer as soon as k_copy_keys has written it, a host thread then spreads it over the
// pageable table with parallel page faults (a direct D2H into fresh pageable memory runs at ~2 GB/s).
struct KeyTableCopy {
	cudaStream_t s2 = nullptr;
	cudaEvent_t ev = nullptr;
	std::thread worker;
	cudaError_t err = cudaSuccess;
	~KeyTableCopy()
	{
		if (worker.joinable()) worker.join();
		if (g_exiting.load()) return;
		if (ev) cudaEventDestroy(ev);
		if (s2) cudaStreamDestroy(s2);
	}
	void start(cudaStream_t s, const uint8_t* d_src, uint8_t* dst, size_t bytes, void* pinned, size_t pinnedBytes)
	{
		SSB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
		SSB_CUDA(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
		SSB_CUDA(cudaEventRecord(ev, s));
		SSB_CUDA(cudaStreamWaitEvent(s2, ev, 0));
		const bool bounce = pinned && pinnedBytes >= bytes && bytes > (1u << 20);
		SSB_CUDA(cudaMemcpyAsync(bounce ? pinned : dst, d_src, bytes, cudaMemcpyDeviceToHost, s2));
		int dev = 0;
		SSB_CUDA(cudaGetDevice(&dev));
		worker = std::thread([this, dev, bounce, dst, pinned, bytes] {
			cudaSetDevice(dev);
			err = cudaStreamSynchronize(s2);
			if (err != cudaSuccess || !bounce) return;
			const unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
			std::vector<std::thread> th;
			const size_t chunk = ((bytes + nt - 1) / nt + 4095) & ~size_t(4095);
			for (unsigned t = 0; t < nt; ++t) {
				const size_t b = (size_t)t * chunk;
				if (b >= bytes) break;
				const size_t e = std::min(bytes, b + chunk);
				th.emplace_back([=] { memcpy(dst + b, static_cast<const uint8_t*>(pinned) + b, e - b); });
			}
			for (auto& t : th) t.join();
		});
	}
	void finish()
	{
		if (worker.joinable()) worker.join();
		cuda_check(err, "key table D2H", __FILE__, __LINE__);
	}
};

template <class U>
void build_impl(DeviceIndex& ix, HostKeys& hk, const PackedWords& pw, uint16_t rowSize, uint16_t gSize, uint64_t rowBase,
                const uint8_t validLut[256], cudaStream_t s, BuildStats* st)
{
	Timer tAll, tPhase;
	ix = DeviceIndex();
	ix.g = gSize; ix.unit = sizeof(U); ix.rowBase = rowBase;
	ix.validLut.alloc(256, s);
	SSB_CUDA(cudaMemcpyAsync(ix.validLut.get(), validLut, 256, cudaMemcpyHostToDevice, s));
	hk.reset();
	KeyTableCopy keyCopy;                                       // joins its thread / frees its stream on every way out
	// hpp:122 — fewer than two words (or no array): nothing is indexed and every query returns 0
	if (pw.n < 2 || !pw.chars || rowSize == 0 || gSize == 0 || gSize > MAX_GSIZE) { SSB_CUDA(cudaStreamSynchronize(s)); return; }
	if (pw.n >= 0xFFFFFFF0ull || pw.totalUnits >= 0xFFFFFFF0ull) throw std::runtime_error("index: more than 2^32 words/units per shard");
	const uint32_t n = (uint32_t)pw.n;
	const uint32_t R = (n + rowSize - 1) / rowSize;

	// ---- H2D
	DevBuf<U> charsBuf;
	DevBuf<uint64_t> offBuf;
	DevBuf<uint8_t> isNullBuf;
	DevBuf<float> weight;
	const U* dChars = static_cast<const U*>(pw.d_chars);
	const uint64_t* dOff = pw.d_off;
	const uint8_t* dIsNull = pw.d_isNull;
	if (!dChars || !dOff || !dIsNull) {
		charsBuf.alloc(std::max<uint64_t>(pw.totalUnits, 1), s); offBuf.alloc(n + 1, s); isNullBuf.alloc(n, s);
		SSB_CUDA(cudaMemcpyAsync(charsBuf.get(), pw.chars, pw.totalUnits * sizeof(U), cudaMemcpyHostToDevice, s));
		SSB_CUDA(cudaMemcpyAsync(offBuf.get(), pw.off, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
		SSB_CUDA(cudaMemcpyAsync(isNullBuf.get(), pw.isNull, n, cudaMemcpyHostToDevice, s));
		dChars = charsBuf.get(); dOff = offBuf.get(); dIsNull = isNullBuf.get();
	}
	if (pw.weight) {
		weight.alloc(n, s);
		SSB_CUDA(cudaMemcpyAsync(weight.get(), pw.weight, n * sizeof(float), cudaMemcpyHostToDevice, s));
	}
	DevBuf<Props> props(1, s);
	if (st) { SSB_CUDA(cudaStreamSynchronize(s)); st->msH2D = tPhase.lap(); }

	DevBuf<uint64_t> termBeg(n, s), termHash(n, s), rowKeyBeg(R, s), rowKeyHash(R, s);
	DevBuf<uint32_t> wordTermLen(n, s), rowKeyLen(R, s);
	DevBuf<uint8_t> edgeValid(n, s), rowValid(R, s);
	DevBuf<uint32_t> rowLeader(R, s), extOfRow(R, s), rowKid(R, s), wordLeader(n, s), tmpOfWord(n, s), edgePos(n, s);
	DevBuf<uint8_t> isLeaderRow(R, s), isLeaderWord(n, s);
	DevBuf<uint32_t> dCount(4, s);
	uint32_t K = 0, T = 0, nEdgeWords = 0;
	DevBuf<uint32_t> kidRowLocal, kidOfExt, termOfTmp, termWord, termLen;
	DevBuf<uint8_t> isShort;

	uint64_t seed = 0x243F6A8885A308D3ull;
	for (int attempt = 0;; ++attempt) {
		if (attempt > 4) throw std::runtime_error("index: persistent 64-bit hash collisions");
		props.zero(s);
		k_word_spans<U><<<grid_for(n), TPB, 0, s>>>(dChars, dOff, dIsNull, weight.get(), n, rowSize, ix.validLut.get(),
		                                             seed, termBeg.get(), wordTermLen.get(), termHash.get(), edgeValid.get(),
		                                             rowKeyBeg.get(), rowKeyLen.get(), rowKeyHash.get(), rowValid.get());
		SSB_LAUNCH_CHECK();
		// ---- keys: intern trimmed raw key strings (hpp:131-135, 58-65)
		DevBuf<uint32_t> rowPos(R, s);
		exclusive_scan_u8_u32(rowValid.get(), rowPos.get(), R, dCount.get(), s);
		const uint32_t nv = d2h_scalar(dCount.get(), s);
		isLeaderRow.zero(s);
		if (nv) {
			DevBuf<uint64_t> k1(nv, s), k2(nv, s);
			DevBuf<uint32_t> v1(nv, s), v2(nv, s), grp(nv, s), leaders(nv, s);
			DevBuf<uint8_t> flag(nv, s);
			k_compact_ids<<<grid_for(R), TPB, 0, s>>>(rowValid.get(), rowPos.get(), rowKeyHash.get(), R, k1.get(), v1.get());
			SSB_LAUNCH_CHECK();
			Sorted so = sort_pairs(k1, k2, v1, v2, nv, 0, 64, s);
			k_flag_new<<<grid_for(nv), TPB, 0, s>>>(so.keys, nv, flag.get());
			SSB_LAUNCH_CHECK();
			exclusive_scan_u8_u32(flag.get(), grp.get(), nv, nullptr, s);
			k_set_leaders<<<grid_for(nv), TPB, 0, s>>>(flag.get(), grp.get(), so.vals, nv, leaders.get());
			SSB_LAUNCH_CHECK();
			k_verify_keys<U><<<grid_for(nv), TPB, 0, s>>>(dChars, flag.get(), grp.get(), so.vals, leaders.get(), nv, rowKeyBeg.get(),
			                                              rowKeyLen.get(), rowLeader.get(), isLeaderRow.get(), props.get());
			SSB_LAUNCH_CHECK();
		}
		exclusive_scan_u8_u32(isLeaderRow.get(), extOfRow.get(), R, dCount.get(), s);
		K = d2h_scalar(dCount.get(), s);
		kidRowLocal.alloc(std::max(K, 1u), s);
		kidOfExt.alloc(std::max(K, 1u), s);
		if (K) {
			DevBuf<uint64_t> k1(K, s), k2(K, s);
			DevBuf<uint32_t> v1(K, s), v2(K, s), extRow(K, s);
			k_ext_info<<<grid_for(R), TPB, 0, s>>>(isLeaderRow.get(), extOfRow.get(), rowKeyLen.get(), R, k1.get(), v1.get(), extRow.get(),
			                                       props.get());
			SSB_LAUNCH_CHECK();
			Props hp = d2h_scalar(props.get(), s);
			Sorted so = sort_pairs(k1, k2, v1, v2, K, 0, bits_for((uint64_t)hp.maxKeyLen + 1), s);
			k_kid_assign<<<grid_for(K), TPB, 0, s>>>(so.vals, extRow.get(), K, kidOfExt.get(), kidRowLocal.get());
			SSB_LAUNCH_CHECK();
		}
		k_row_kid<<<grid_for(R), TPB, 0, s>>>(rowValid.get(), rowLeader.get(), extOfRow.get(), kidOfExt.get(), R, rowKid.get());
		SSB_LAUNCH_CHECK();

		// ---- terms: intern normalised terms of every edge-creating word
		exclusive_scan_u8_u32(edgeValid.get(), edgePos.get(), n, dCount.get(), s);
		nEdgeWords = d2h_scalar(dCount.get(), s);
		isLeaderWord.zero(s);
		if (nEdgeWords) {
			DevBuf<uint64_t> k1(nEdgeWords, s), k2(nEdgeWords, s);
			DevBuf<uint32_t> v1(nEdgeWords, s), v2(nEdgeWords, s), grp(nEdgeWords, s), leaders(nEdgeWords, s);
			DevBuf<uint8_t> flag(nEdgeWords, s);
			k_compact_ids<<<grid_for(n), TPB, 0, s>>>(edgeValid.get(), edgePos.get(), termHash.get(), n, k1.get(), v1.get());
			SSB_LAUNCH_CHECK();
			Sorted so = sort_pairs(k1, k2, v1, v2, nEdgeWords, 0, 64, s);
			k_flag_new<<<grid_for(nEdgeWords), TPB, 0, s>>>(so.keys, nEdgeWords, flag.get());
			SSB_LAUNCH_CHECK();
			exclusive_scan_u8_u32(flag.get(), grp.get(), nEdgeWords, nullptr, s);
			k_set_leaders<<<grid_for(nEdgeWords), TPB, 0, s>>>(flag.get(), grp.get(), so.vals, nEdgeWords, leaders.get());
			SSB_LAUNCH_CHECK();
			k_verify_terms<U><<<grid_for(nEdgeWords), TPB, 0, s>>>(dChars, ix.validLut.get(), flag.get(), grp.get(), so.vals,
			                                                       leaders.get(), nEdgeWords, termBeg.get(), wordTermLen.get(),
			                                                       wordLeader.get(), isLeaderWord.get(), props.get());
			SSB_LAUNCH_CHECK();
		}
		exclusive_scan_u8_u32(isLeaderWord.get(), tmpOfWord.get(), n, dCount.get(), s);
		T = d2h_scalar(dCount.get(), s);
		Props hp = d2h_scalar(props.get(), s);
		if (!hp.collision) break;
		seed = mix64(seed + attempt + 1);                       // distinct strings shared a 64-bit hash: re-hash with a new seed
		if (st) st->hashRetries++;
	}
	if (st) st->msIntern = tPhase.lap();

	ix.K = K; ix.T = T;
	if (T == 0 || K == 0) {                                     // valid call, nothing indexable: indexed, but every table is empty
		ix.indexed = true;
		ix.gramCodes.alloc(1, s); ix.gramOff.alloc(1, s); ix.gramOff.zero(s);
		SSB_CUDA(cudaStreamSynchronize(s));
		if (st) st->msTotal = tAll.lap();
		return;
	}

	// ---- term ids: rank by (min kid, first word)
	termOfTmp.alloc(T, s); termWord.alloc(T, s); termLen.alloc(T, s); isShort.alloc(T, s);
	{
		DevBuf<uint32_t> minKid(T, s), tmpLeaderWord(T, s);
		SSB_CUDA(cudaMemsetAsync(minKid.get(), 0xFF, T * sizeof(uint32_t), s));
		k_min_kid<<<grid_for(n), TPB, 0, s>>>(edgeValid.get(), wordLeader.get(), tmpOfWord.get(), rowKid.get(), n, rowSize, minKid.get(),
		                                      tmpLeaderWord.get());
		SSB_LAUNCH_CHECK();
		DevBuf<uint64_t> k1(T, s), k2(T, s);
		DevBuf<uint32_t> v1(T, s), v2(T, s);
		k_term_sort_keys<<<grid_for(T), TPB, 0, s>>>(minKid.get(), T, k1.get(), v1.get());
		SSB_LAUNCH_CHECK();
		Sorted so = sort_pairs(k1, k2, v1, v2, T, 0, bits_for(K), s);
		k_term_assign<<<grid_for(T), TPB, 0, s>>>(so.vals, tmpLeaderWord.get(), wordTermLen.get(), T, 2u * gSize, termOfTmp.get(),
		                                          termWord.get(), termLen.get(), isShort.get(), props.get());
		SSB_LAUNCH_CHECK();
	}

	// ---- edges: unique (term, kid), last weight wins; CSR by term and by key
	uint32_t E = 0;
	{
		DevBuf<uint64_t> k1(nEdgeWords, s), k2(nEdgeWords, s);
		DevBuf<uint32_t> v1(nEdgeWords, s), v2(nEdgeWords, s), pos(nEdgeWords, s);
		DevBuf<uint8_t> flag(nEdgeWords, s);
		k_edge_keys<<<grid_for(n), TPB, 0, s>>>(edgeValid.get(), edgePos.get(), wordLeader.get(), tmpOfWord.get(), termOfTmp.get(),
		                                        rowKid.get(), n, rowSize, k1.get(), v1.get());
		SSB_LAUNCH_CHECK();
		bool a1 = radix_sort_u64(k1.get(), k2.get(), v1.get(), v2.get(), nEdgeWords, 0, bits_for(K), s);
		uint64_t *ka = a1 ? k2.get() : k1.get(), *kb = a1 ? k1.get() : k2.get();
		uint32_t *va = a1 ? v2.get() : v1.get(), *vb = a1 ? v1.get() : v2.get();
		bool a2 = radix_sort_u64(ka, kb, va, vb, nEdgeWords, 32, 32 + bits_for(T), s);
		uint64_t* sk = a2 ? kb : ka;
		uint32_t* sv = a2 ? vb : va;
		k_flag_last<<<grid_for(nEdgeWords), TPB, 0, s>>>(sk, nEdgeWords, flag.get());
		SSB_LAUNCH_CHECK();
		exclusive_scan_u8_u32(flag.get(), pos.get(), nEdgeWords, dCount.get(), s);
		E = d2h_scalar(dCount.get(), s);
		ix.E = E;
		DevBuf<uint32_t> edgeTerm(E, s);
		ix.edgeKid.alloc(E, s); ix.edgeW.alloc(E, s);
		k_compact_edges<<<grid_for(nEdgeWords), TPB, 0, s>>>(sk, sv, flag.get(), pos.get(), weight.get(), nEdgeWords, edgeTerm.get(),
		                                                     ix.edgeKid.get(), ix.edgeW.get(), props.get());
		SSB_LAUNCH_CHECK();
		ix.termEdgeOff.alloc(T + 1, s);
		k_csr_offsets<<<grid_for(E + 1), TPB, 0, s>>>(edgeTerm.get(), E, T, ix.termEdgeOff.get());
		SSB_LAUNCH_CHECK();
		// reverse CSR (key -> terms)
		DevBuf<uint64_t> r1(E, s), r2(E, s);
		DevBuf<uint32_t> rv1(E, s), rv2(E, s), revKid(E, s);
		k_rev_keys<<<grid_for(E), TPB, 0, s>>>(edgeTerm.get(), ix.edgeKid.get(), E, r1.get(), rv1.get());
		SSB_LAUNCH_CHECK();
		bool b1 = radix_sort_u64(r1.get(), r2.get(), rv1.get(), rv2.get(), E, 0, bits_for(T), s);
		uint64_t *rka = b1 ? r2.get() : r1.get(), *rkb = b1 ? r1.get() : r2.get();
		uint32_t *rva = b1 ? rv2.get() : rv1.get(), *rvb = b1 ? rv1.get() : rv2.get();
		bool b2 = radix_sort_u64(rka, rkb, rva, rvb, E, 32, 32 + bits_for(K), s);
		ix.keyEdgeOff.alloc(K + 1, s); ix.keyEdgeTerm.alloc(E, s); ix.keyEdgeW.alloc(E, s); ix.keyWmax.alloc(K, s);
		k_rev_fill<<<grid_for(E), TPB, 0, s>>>(b2 ? rkb : rka, b2 ? rvb : rva, ix.edgeW.get(), E, revKid.get(), ix.keyEdgeTerm.get(),
		                                       ix.keyEdgeW.get());
		SSB_LAUNCH_CHECK();
		k_csr_offsets<<<grid_for(E + 1), TPB, 0, s>>>(revKid.get(), E, K, ix.keyEdgeOff.get());
		SSB_LAUNCH_CHECK();
		k_key_wmax<<<grid_for(K), TPB, 0, s>>>(ix.keyEdgeOff.get(), ix.keyEdgeW.get(), K, ix.keyWmax.get(), props.get());
		SSB_LAUNCH_CHECK();
	}
	if (st) { SSB_CUDA(cudaStreamSynchronize(s)); st->msEdges = tPhase.lap(); }

	// ---- strings: normalised terms (by term id), raw keys (by kid, 0-terminated)
	ix.termOff.alloc(T + 1, s);
	exclusive_scan_u32_u32(termLen.get(), ix.termOff.get(), T, ix.termOff.get() + T, s);
	DevBuf<uint32_t> keyLen1(K, s);
	ix.keyOff.alloc(K + 1, s);
	k_len_plus1<<<grid_for(K), TPB, 0, s>>>(kidRowLocal.get(), rowKeyLen.get(), K, keyLen1.get());
	SSB_LAUNCH_CHECK();
	exclusive_scan_u32_u32(keyLen1.get(), ix.keyOff.get(), K, ix.keyOff.get() + K, s);
	const uint32_t termUnits = d2h_scalar(ix.termOff.get() + T, s);
	const uint32_t keyUnits = d2h_scalar(ix.keyOff.get() + K, s);
	ix.termUnits = termUnits; ix.keyUnits = keyUnits;
	ix.termChars.alloc(std::max<size_t>((size_t)termUnits * sizeof(U), 4), s);
	ix.keyChars.alloc((size_t)keyUnits * sizeof(U), s);
	ix.keyRow.alloc(K, s);
	ix.keyHash.alloc(K, s);
	k_copy_terms<U><<<grid_for(T), TPB, 0, s>>>(dChars, ix.validLut.get(), termWord.get(), termBeg.get(), ix.termOff.get(), T,
	                                            reinterpret_cast<U*>(ix.termChars.get()));
	SSB_LAUNCH_CHECK();
	k_copy_keys<U><<<grid_for(K), TPB, 0, s>>>(dChars, kidRowLocal.get(), rowKeyBeg.get(), ix.keyOff.get(), K, rowBase,
	                                           reinterpret_cast<U*>(ix.keyChars.get()), ix.keyRow.get(), ix.keyHash.get());
	SSB_LAUNCH_CHECK();
	// host key table: on its way from here on (the staging buffer is free: the upload of the words was ordered before k_copy_keys)
	hk.K = K;
	hk.charBytes = (size_t)keyUnits * sizeof(U);
	hk.chars = static_cast<uint8_t*>(malloc(hk.charBytes ? hk.charBytes : 1));
	if (!hk.chars) throw std::runtime_error("index: out of host memory for the key table");
	if (hk.charBytes) keyCopy.start(s, ix.keyChars.get(), hk.chars, hk.charBytes, pw.pinnedScratch, pw.pinnedScratchBytes);
	// short terms list (hpp:82-85)
	{
		DevBuf<uint32_t> pos(T, s);
		exclusive_scan_u8_u32(isShort.get(), pos.get(), T, dCount.get(), s);
		ix.nShort = d2h_scalar(dCount.get(), s);
		ix.shortTerms.alloc(std::max(ix.nShort, 1u), s);
		k_compact_short<<<grid_for(T), TPB, 0, s>>>(isShort.get(), pos.get(), T, ix.shortTerms.get());
		SSB_LAUNCH_CHECK();
	}

	// ---- grams
	const uint32_t twoG = 2u * gSize;
	if (sizeof(U) == 1) {
		DevBuf<uint32_t> present(256, s);
		present.zero(s);
		ix.symTable.alloc(256, s);
		k_presence_narrow<<<grid_for(T), TPB, 0, s>>>(ix.termChars.get(), ix.termOff.get(), T, twoG, present.get());
		SSB_LAUNCH_CHECK();
		k_build_lut_narrow<<<1, 32, 0, s>>>(present.get(), ix.symTable.get(), dCount.get());
		SSB_LAUNCH_CHECK();
		ix.A = d2h_scalar(dCount.get(), s);
	} else {
		DevBuf<uint64_t> u1(std::max(termUnits, 1u), s), u2(std::max(termUnits, 1u), s);
		DevBuf<uint8_t> flag(std::max(termUnits, 1u), s);
		DevBuf<uint32_t> pos(std::max(termUnits, 1u), s);
		k_units_wide<<<grid_for(T), TPB, 0, s>>>(reinterpret_cast<const uint32_t*>(ix.termChars.get()), ix.termOff.get(), T, twoG,
		                                         u1.get(), termUnits);
		SSB_LAUNCH_CHECK();
		bool alt = radix_sort_u64(u1.get(), u2.get(), nullptr, nullptr, termUnits, 0, 40, s);
		uint64_t* su = alt ? u2.get() : u1.get();
		k_flag_new<<<grid_for(termUnits), TPB, 0, s>>>(su, termUnits, flag.get());
		SSB_LAUNCH_CHECK();
		exclusive_scan_u8_u32(flag.get(), pos.get(), termUnits, dCount.get(), s);
		uint32_t distinct = termUnits ? d2h_scalar(dCount.get(), s) : 0;
		// the sentinel (units of short terms) sorts last; drop it from the alphabet if present
		uint64_t lastVal = 0;
		if (termUnits) lastVal = d2h_scalar(su + termUnits - 1, s);
		ix.A = distinct - ((termUnits && lastVal == 0xFFFFFFFFFFull) ? 1 : 0);
		ix.symTable.alloc(std::max(ix.A, 1u), s);
		if (termUnits) {
			k_compact_alpha<<<grid_for(termUnits), TPB, 0, s>>>(su, flag.get(), pos.get(), termUnits, ix.symTable.get());
			SSB_LAUNCH_CHECK();
		}
	}
	ix.symBits = bits_for((uint64_t)ix.A + 1);
	if (gSize * ix.symBits > 64) throw std::runtime_error("index: gSize * log2(alphabet) exceeds 64 bits");
	ix.packed = gSize * ix.symBits <= 32;
	{
		DevBuf<uint32_t> wcnt(T, s), woff(T + 1, s);
		k_window_counts<<<grid_for(T), TPB, 0, s>>>(ix.termOff.get(), T, gSize, wcnt.get());
		SSB_LAUNCH_CHECK();
		DevBuf<uint64_t> total64(1, s);
		{
			// total window count may approach 2^32: accumulate the grand total in 64 bits
			DevBuf<uint64_t> woff64(T, s);
			exclusive_scan_u32_u64(wcnt.get(), woff64.get(), T, total64.get(), s);
		}
		const uint64_t W = d2h_scalar(total64.get(), s);
		if (W >= 0xFFFFFFF0ull) throw std::runtime_error("index: more than 2^32 gram windows per shard");
		exclusive_scan_u32_u32(wcnt.get(), woff.get(), T, woff.get() + T, s);
		if (st) st->windows = W;
		const uint32_t nW = (uint32_t)W;
		if (nW) {
			DevBuf<uint64_t> k1(nW, s), k2(nW, s);
			DevBuf<uint32_t> v1, v2;
			if (!ix.packed) { v1.alloc(nW, s); v2.alloc(nW, s); }
			const U* tc = reinterpret_cast<const U*>(ix.termChars.get());
			if (ix.packed)
				k_emit_grams<U, true><<<grid_for(T), TPB, 0, s>>>(tc, ix.termOff.get(), woff.get(), T, gSize, ix.symBits, ix.symTable.get(),
				                                                  ix.A, k1.get(), nullptr);
			else
				k_emit_grams<U, false><<<grid_for(T), TPB, 0, s>>>(tc, ix.termOff.get(), woff.get(), T, gSize, ix.symBits, ix.symTable.get(),
				                                                   ix.A, k1.get(), v1.get());
			SSB_LAUNCH_CHECK();
			const int codeBits = gSize * ix.symBits;
			bool alt = ix.packed ? radix_sort_u64(k1.get(), k2.get(), nullptr, nullptr, nW, 32, 32 + codeBits, s)
			                     : radix_sort_u64(k1.get(), k2.get(), v1.get(), v2.get(), nW, 0, codeBits, s);
			uint64_t* sk = alt ? k2.get() : k1.get();
			uint32_t* sv = ix.packed ? nullptr : (alt ? v2.get() : v1.get());
			DevBuf<uint8_t> newPair(nW, s), newGram(nW, s);
			DevBuf<uint32_t> pairPos(nW, s), gramPos(nW, s);
			if (ix.packed) k_rle_flags<true><<<grid_for(nW), TPB, 0, s>>>(sk, sv, nW, newPair.get(), newGram.get());
			else k_rle_flags<false><<<grid_for(nW), TPB, 0, s>>>(sk, sv, nW, newPair.get(), newGram.get());
			SSB_LAUNCH_CHECK();
			exclusive_scan_u8_u32(newPair.get(), pairPos.get(), nW, dCount.get(), s);
			exclusive_scan_u8_u32(newGram.get(), gramPos.get(), nW, dCount.get() + 1, s);
			ix.P = d2h_scalar(dCount.get(), s);
			ix.G = d2h_scalar(dCount.get() + 1, s);
			ix.postings.alloc((size_t)ix.P + 4, s);   // +4: TMA window copies round up to 16 bytes
			ix.gramCodes.alloc(ix.G, s); ix.gramOff.alloc(ix.G + 1, s);
			if (ix.packed)
				k_rle_write<true><<<grid_for(nW), TPB, 0, s>>>(sk, sv, newPair.get(), newGram.get(), pairPos.get(), gramPos.get(), nW,
				                                               ix.postings.get(), ix.gramCodes.get(), ix.gramOff.get());
			else
				k_rle_write<false><<<grid_for(nW), TPB, 0, s>>>(sk, sv, newPair.get(), newGram.get(), pairPos.get(), gramPos.get(), nW,
				                                                ix.postings.get(), ix.gramCodes.get(), ix.gramOff.get());
			SSB_LAUNCH_CHECK();
			k_set_u32<<<1, 1, 0, s>>>(ix.gramOff.get() + ix.G, ix.P);
			SSB_LAUNCH_CHECK();
		} else {
			ix.gramCodes.alloc(1, s); ix.gramOff.alloc(1, s); ix.gramOff.zero(s); ix.postings.alloc(4, s);
		}
	}
	if (st) { SSB_CUDA(cudaStreamSynchronize(s)); st->msGrams = tPhase.lap(); }

	// ---- sieve: block bitmaps over the posting lists (derived; sieve.cu)
	ix.indexed = true;                                          // hpp:45
	build_sieve(ix, s);
	if (st) { SSB_CUDA(cudaStreamSynchronize(s)); st->msSieve = tPhase.lap(); }

	// ---- properties + host key table
	Props hp = d2h_scalar(props.get(), s);
	ix.identity = !hp.notIdentity && T == K && E == T;
	ix.uniformW = !hp.notUniform;
	ix.multiEdgeKeys = hp.multiEdge != 0;
	ix.wmax = ordered_to_float(hp.wmaxOrdered);
	ix.maxTermLen = hp.maxTermLen; ix.maxKeyLen = hp.maxKeyLen;
	keyCopy.finish();                                           // (usually there already: it started before the grams)
	if (st) { st->msD2H = tPhase.lap(); st->msTotal = tAll.lap(); }
}

}   // namespace

void build_index(DeviceIndex& ix, HostKeys& hk, const PackedWords& w, uint16_t rowSize, uint16_t gSize, int unitBytes, uint64_t rowBase,
                 const uint8_t validLut[256], cudaStream_t s, BuildStats* stats)
{
	if (unitBytes == 1) build_impl<uint8_t>(ix, hk, w, rowSize, gSize, rowBase, validLut, s, stats);
	else if (unitBytes == 4) build_impl<uint32_t>(ix, hk, w, rowSize, gSize, rowBase, validLut, s, stats);
	else throw std::runtime_error("index: unit size must be 1 or 4 bytes");
}

}   // namespace ssb
