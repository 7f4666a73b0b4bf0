// sieve_path.cuh — k_sieve_search: top-k over posting lists WITHOUT reading most postings (included by search.cu after
// fast_path.cuh, whose scoring / select / sort helpers it shares).
//
// The reference increments a counter for every posting of every query gram (nGramSearch.hpp:289-298). Here the K lists of a
// query are first looked at through their block bitmaps (sieve.cu): one bit per block of 2^shift consecutive term ids. With
// tau the running k-th best (or the threshold, hpp:313-316) a term needs c = cRej + 1 matches to matter, and it can only
// have them inside a block in which at least c of the K bitmaps have their bit set. Per pass over a range of bitmap words:
//   SIEVE    every thread streams two 32-block words of each list (one coalesced 8-byte load per list, all K in flight)
//            through the bit-sliced threshold counter  ge[c] |= ge[c-1] & x  (c LOP3 per word and list): ge[c-1] = blocks
//            hit by >= c lists. The numbers of the blocks that pass are queued (one warp scan + one shared atomic per warp).
//   RESOLVE  one thread per queued block: which lists hit it (their words are still in L1/L2), then for up to four lists at a
//            time the rank of the block inside its 256-block group (popcount of the lower bits of the group's 8 words) + the
//            group's directory entry = position of the block's first posting, which is read from the bit-exact CSR postings
//            (a second posting of the same list in the block, rare, follows it). The terms found are counted in registers —
//            the reference's match count, with multiplicities — filtered by the integer bound derived from tau and handed to
//            the same emit / deferred-evaluation code as k_fast_search.
// Passes run in ascending term-id order; tau moves from class to class through bounds that cost nothing (k candidates with
// >= c matches are known, all below the end of the pass), so that c rises and the sieve gets sparser as the query proceeds.
// On identity indexes without threshold the terms with ONE match are left for last: the first phase looks for terms with
// two or more matches only (no seeding pass over every posting of the smallest ids), and only if fewer than k exist does a
// second phase fill the list with the smallest single-match ids.
// Bytes read per query: K bitmaps (~ the bytes of the K lists at the chosen block size) + ~100 B per (block, list) resolved,
// against 4 bytes + ~150 instructions per posting for the merge. Exactness does not depend on the bitmaps being selective:
// a block that passes is resolved from the postings themselves.
// Queries the kernel does not take (a list without bitmap, more than 16 lists, the edit-distance path) or cannot finish
// (candidate buffer too small for one 64-block step, more than 8 distinct terms of interest in one block) are handed to
// k_fast_search through handList.
#pragma once

namespace {

constexpr int SV_THREADS = 128;
constexpr int SV_CCAP = 2048;                           // candidate blocks per pass
constexpr int SV_PU = 4;                                // lists of one block resolved at a time (independent load chains)
constexpr int SV_DISTINCT = 8;                          // distinct terms tracked per block

struct SieveShared {
	const uint32_t* sievePtr[SV_MAXL];                  // bitmaps as the sieve counts them: list j repeated mult[j] times
	const uint32_t* bitsPtr[SV_MAXL];                   // bitmap of list j
	const uint32_t* cumPtr[SV_MAXL];                    // its group directory
	uint32_t listEnd[SV_MAXL], mult[SV_MAXL];
	uint32_t cand[SV_CCAP];                             // numbers of the blocks that passed the sieve
	unsigned long long* buf;                            // candidate keys: bufCap entries behind this struct in dynamic smem
	uint32_t bufCap;
	uint32_t defCount, defCap;
	unsigned long long* defList;
	union {
		unsigned long long deferred[FS_DEF_CAP];
		uint32_t hist[256];
	};
	uint32_t geHist[MAX_QGRAMS + 1];                    // candidates appended with exactly c matches (class mode)
	uint32_t geSnap[MAX_QGRAMS + 1];                    // ... at the start of the pass in flight (restored when it is redone)
	float rateTab[9];                                   // expected candidate blocks per bitmap word when c lists must hit a block
	uint64_t qcode[MAX_QGRAMS];
	uint32_t hiKey[MAX_QGRAMS + 1], exactKey[MAX_QGRAMS + 1];
	float sOf[MAX_QGRAMS + 1];
	unsigned long long tau;
	uint32_t promotedKey;
	uint32_t bufCount, overflow, query, selPrefixLo, selK;
	uint32_t nCand, nProbe;
	uint32_t passW0, passW1, cRej, done;
	uint32_t qlen, nGrams, cmin, cprom, nLists, sumMult;
	uint32_t shortActive, scanLong, scanLongOnly;       // read by the shared scoring helpers; always 0 here
	float threshold;
	int topShift, kidBits;
	uint32_t classMode;
	IndexView ix;
};

__device__ __forceinline__ uint2 ldg_u2(const uint32_t* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }

// ---- SIEVE: words [w0, w1) (both even) of every list -> queue of blocks hit by >= CREQ lists. A list whose gram occurs m
// times in the query appears m times in sievePtr (hpp:289-298 counts once per window), so the counter needs no multiplicities.
template <int CREQ>
__device__ __forceinline__ void sieve_words(SieveShared& sh, const uint32_t w0, const uint32_t w1, const uint32_t nS)
{
	const int tid = threadIdx.x, lane = tid & 31;
	for (uint32_t base = w0; base < w1; base += 2 * SV_THREADS) {
		const uint32_t w = base + 2 * tid;
		const bool act = w < w1;
		uint32_t ga[CREQ], gb[CREQ];
#pragma unroll
		for (int c = 0; c < CREQ; ++c) { ga[c] = 0; gb[c] = 0; }
		if (act) {
#pragma unroll 4
			for (uint32_t j = 0; j < nS; ++j) {
				const uint2 x = ldg_u2(sh.sievePtr[j] + w);
#pragma unroll
				for (int c = CREQ - 1; c >= 1; --c) { ga[c] |= ga[c - 1] & x.x; gb[c] |= gb[c - 1] & x.y; }
				ga[0] |= x.x; gb[0] |= x.y;
			}
		}
		uint32_t ma = ga[CREQ - 1], mb = gb[CREQ - 1];
		if (!__any_sync(0xffffffffu, (ma | mb) != 0)) continue;
		const uint32_t n = __popc(ma) + __popc(mb);
		uint32_t inc = n;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
		const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
		uint32_t cBase = 0;
		if (lane == 31) cBase = atomicAdd(&sh.nCand, tot);
		cBase = __shfl_sync(0xffffffffu, cBase, 31);
		if (cBase + tot > SV_CCAP) continue;                    // pass too long: counted, redone shorter
		uint32_t ci = cBase + inc - n;
		while (ma) { sh.cand[ci++] = (w << 5) | (__ffs(ma) - 1); ma &= ma - 1; }
		while (mb) { sh.cand[ci++] = ((w + 1) << 5) | (__ffs(mb) - 1); mb &= mb - 1; }
	}
}

__device__ __forceinline__ void sieve_dispatch(SieveShared& sh, uint32_t creq, uint32_t w0, uint32_t w1, uint32_t nS)
{
	// thresholds without an instantiation of their own use the next lower one: a superset of the blocks, RESOLVE counts exactly
	if (creq <= 1) sieve_words<1>(sh, w0, w1, nS);
	else if (creq == 2) sieve_words<2>(sh, w0, w1, nS);
	else if (creq == 3) sieve_words<3>(sh, w0, w1, nS);
	else if (creq < 6) sieve_words<4>(sh, w0, w1, nS);
	else if (creq < 8) sieve_words<6>(sh, w0, w1, nS);
	else sieve_words<8>(sh, w0, w1, nS);
}

// a term with exact match count c > cRej (same tail as merge_pass); emits are rare -> one copy of the code
__device__ __noinline__ void sieve_emit(SieveShared& sh, uint32_t term, uint32_t c, unsigned long long tau)
{
	if (sh.classMode && c < sh.cprom) {
		const unsigned long long key = (((unsigned long long)sh.exactKey[c]) << 32) | term;
		if (key < tau) {
			append_key(sh, key);
			atomicAdd(&sh.geHist[c], 1u);
		}
	} else {
		const uint32_t di = atomicAdd(&sh.defCount, 1u);
		if (di < sh.defCap) sh.defList[di] = (((unsigned long long)c) << 32) | term; else sh.overflow = 1;
	}
}

// position of the first posting of list j in block (grp * 8 + wi) * 32 + bit, given that the list's bitmap has the bit set:
// directory entry of the 256-block group + number of occupied blocks below (exact unless one of those holds several postings)
__device__ __forceinline__ uint32_t sieve_locate(const SieveShared& sh, uint32_t j, uint32_t grp, uint32_t wi, uint32_t bit)
{
	const uint4* gp = reinterpret_cast<const uint4*>(sh.bitsPtr[j] + (grp << 3));
	uint32_t idx = __ldg(sh.cumPtr[j] + grp);
	const uint4 a = __ldg(gp), b = __ldg(gp + 1);
	const uint32_t ws[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
	for (int t = 0; t < 8; ++t) {
		const uint32_t m = (uint32_t)t < wi ? 0xFFFFFFFFu : ((uint32_t)t == wi ? ((1u << bit) - 1u) : 0u);
		idx += __popc(ws[t] & m);
	}
	return idx;
}

// general resolution of one block (more than SV_PU lists hit it, or a list has several postings in it): one list at a
// time into local arrays, counts by comparing all pairs. Rare.
__device__ __noinline__ void resolve_slow(SieveShared& sh, uint32_t blk, uint32_t lm, uint32_t cRej, uint32_t onlyCount, unsigned long long tau)
{
	constexpr int CAP = 48;
	const int shift = sh.ix.sieveShift;
	const uint32_t* __restrict__ postings = sh.ix.postings;
	const uint32_t w = blk >> 5, bit = blk & 31u, grp = w >> 3, wi = w & 7u;
	uint32_t ids[CAP], mus[CAP], m = 0;
	while (lm) {
		const uint32_t j = (uint32_t)__ffs(lm) - 1u;
		lm &= lm - 1;
		const uint32_t end = sh.listEnd[j], mu = sh.mult[j];
		uint32_t p = sieve_locate(sh, j, grp, wi, bit);
		bool any = false;
		for (; p < end; ++p) {
			const uint32_t id = __ldg(postings + p);
			const uint32_t b = id >> shift;
			if (b < blk) continue;                           // extra postings of lower blocks
			if (b > blk) break;
			if (m < CAP) { ids[m] = id; mus[m] = mu; ++m; } else sh.overflow = 2;
			any = true;
		}
		if (!any) sh.overflow = 2;                           // bitmap and postings disagree: never trust, hand the query over
	}
	for (uint32_t a = 0; a < m; ++a) {
		uint32_t c = 0;
		bool first = true;
		for (uint32_t b = 0; b < m; ++b)
			if (ids[b] == ids[a]) { if (b < a) first = false; c += mus[b]; }
		if (first && c > cRej && (!onlyCount || c == onlyCount)) sieve_emit(sh, ids[a], c, tau);
	}
}

// ---- RESOLVE: one thread per queued block. onlyCount != 0: emit only terms with exactly that many matches (second
// phase: the others are in the buffer already).
__device__ __forceinline__ void resolve_stage(SieveShared& sh, const uint32_t nCand, const uint32_t nL, const uint32_t onlyCount)
{
	const bool identMode = sh.classMode != 0;
	const unsigned long long tau = sh.tau;
	const uint32_t nG = sh.nGrams;
	const int shift = sh.ix.sieveShift;
	const uint32_t* __restrict__ postings = sh.ix.postings;
	uint32_t probes = 0;
	for (uint32_t i = threadIdx.x; i < nCand; i += SV_THREADS) {
		const uint32_t blk = sh.cand[i];
		const uint32_t w = blk >> 5, bit = blk & 31u, grp = w >> 3, wi = w & 7u;
		// which lists hit the block (their words have just been streamed by the sieve: L1 / L2 hits)
		uint32_t lm = 0;
#pragma unroll 4
		for (uint32_t j = 0; j < nL; ++j) lm |= ((__ldg(sh.bitsPtr[j] + w) >> bit) & 1u) << j;
		probes += __popc(lm);
		// integer pre-filter refined by the block's first id (identity mode: kid == term id >= first id of the block)
		uint32_t cRej = sh.cRej;
		if (identMode) {
			const unsigned long long lo = (unsigned long long)blk << shift;
			while (cRej < nG && ((((unsigned long long)sh.hiKey[cRej + 1]) << 32) | lo) >= tau) ++cRej;
		}
		bool slow = __popc(lm) > SV_PU;
		if (!slow) {
			// up to SV_PU lists: independent load chains, everything stays in registers
			uint32_t h[SV_PU], mu[SV_PU], idx[SV_PU], end[SV_PU], id0[SV_PU], id1[SV_PU];
			bool on[SV_PU];
			uint32_t rest = lm;
#pragma unroll
			for (int u = 0; u < SV_PU; ++u) {
				on[u] = rest != 0;
				const uint32_t j = on[u] ? (uint32_t)__ffs(rest) - 1u : 0u;
				rest &= rest - 1;
				idx[u] = on[u] ? sieve_locate(sh, j, grp, wi, bit) : 0u;
				end[u] = on[u] ? sh.listEnd[j] : 0u;
				mu[u] = sh.mult[j];
			}
#pragma unroll
			for (int u = 0; u < SV_PU; ++u) {
				id0[u] = idx[u] < end[u] ? __ldg(postings + idx[u]) : INF;
				id1[u] = idx[u] + 1 < end[u] ? __ldg(postings + idx[u] + 1) : INF;
			}
#pragma unroll
			for (int u = 0; u < SV_PU; ++u) {
				h[u] = INF;
				if (!on[u]) continue;
				// the common case: the first posting read is the block's only posting of the list; anything else -> general path
				if ((id0[u] >> shift) != blk || (id1[u] != INF && (id1[u] >> shift) == blk)) slow = true;
				h[u] = id0[u];
			}
			if (!slow) {
#pragma unroll
				for (int u = 0; u < SV_PU; ++u) {
					if (!on[u]) continue;
					uint32_t c = 0;
					bool first = true;
#pragma unroll
					for (int v = 0; v < SV_PU; ++v)
						if (on[v] && h[v] == h[u]) { c += mu[v]; if (v < u) first = false; }
					if (first && c > cRej && (!onlyCount || c == onlyCount)) sieve_emit(sh, h[u], c, tau);
				}
			}
		}
		if (slow) resolve_slow(sh, blk, lm, cRej, onlyCount, tau);
	}
	if (probes) atomicAdd(&sh.nProbe, probes);
}

// expected candidate blocks per bitmap word when a block needs >= c of K lists, every list setting a block with probability p
__device__ float sieve_rate(uint32_t K, uint32_t c, float p)
{
	if (p >= 0.999f) return 32.0f;
	float pmf = powf(1.0f - p, (float)K), e = c == 0 ? pmf : 0.0f;      // pmf(0)
	const float r = p / (1.0f - p);
	for (uint32_t m = 0; m < K; ++m) {
		pmf = pmf * (float)(K - m) / (float)(m + 1) * r;                 // pmf(m + 1)
		if (m + 1 >= c) e += pmf;
	}
	return 32.0f * e;
}

template <class U>
__global__ void __launch_bounds__(SV_THREADS, 6) k_sieve_search(IndexView ix, const QDesc* __restrict__ desc, const U* __restrict__ qnorm,
                                                             const uint32_t* __restrict__ fastList, uint32_t* __restrict__ counters,
                                                             uint32_t* __restrict__ handList, uint32_t limit, uint32_t outStride,
                                                             uint32_t* __restrict__ outCount, uint32_t* __restrict__ outKid,
                                                             float* __restrict__ outScore, unsigned long long* __restrict__ statCounters,
                                                             uint32_t bufCap, float threshold, uint32_t defExt, HostOut ho)
{
	extern __shared__ __align__(16) unsigned char sv_smem[];
	SieveShared& sh = *reinterpret_cast<SieveShared*>(sv_smem);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t nFast = counters[0];
	const uint32_t k = limit;
	const uint32_t W = ix.sieveWords, nGroups1 = (W >> 3) + 1;
	const int shift = ix.sieveShift;
	const uint32_t nBlocks = (uint32_t)((((unsigned long long)ix.T - 1) >> shift) + 1);
	const uint32_t Wused = (((nBlocks + 31) >> 5) + 1) & ~1u;          // even: the sieve reads word pairs (W % 8 == 0 covers it)
	if (tid == 0) {
		sh.ix = ix;
		sh.buf = reinterpret_cast<unsigned long long*>(sv_smem + ((sizeof(SieveShared) + 15) & ~size_t(15)));
		sh.bufCap = bufCap;
		sh.defList = defExt ? sh.buf + bufCap : sh.deferred;
		sh.defCap = defExt ? defExt : (uint32_t)FS_DEF_CAP;
		sh.shortActive = 0; sh.scanLong = 0; sh.scanLongOnly = 0; sh.threshold = threshold;
	}

	while (true) {
		__syncthreads();
		if (tid == 0) sh.query = atomicAdd(&counters[2], 1u);
		__syncthreads();
		const uint32_t wi = sh.query;
		if (wi >= nFast) break;
		const uint32_t q = fastList[wi];
		const QDesc& d = desc[q];
		if (!(d.flags & 4u)) {                          // not for this kernel: k_fast_search takes it
			if (tid == 0) handList[atomicAdd(&counters[5], 1u)] = q;
			continue;
		}
		const U* qn = qnorm + d.qbeg;
		const uint32_t nLists = d.nLists, nGrams = d.nGrams;
		// average bit density of a list's bitmap (pass-length estimates only)
		const float dens = fminf(1.0f, (float)d.total / ((float)nLists * (float)nBlocks));
		if (tid < SV_MAXL) {
			if ((uint32_t)tid < nLists) {
				const uint32_t slot = d.listSlot[tid];
				sh.bitsPtr[tid] = ix.sieveBits + (size_t)slot * W;
				sh.cumPtr[tid] = ix.sieveCum + (size_t)slot * nGroups1;
				sh.listEnd[tid] = d.listOff[tid] + d.listLen[tid];
				sh.mult[tid] = d.listMult[tid];
			} else { sh.bitsPtr[tid] = ix.sieveBits; sh.cumPtr[tid] = ix.sieveCum; sh.listEnd[tid] = 0; sh.mult[tid] = 0; }
		} else if (tid >= 32 && tid < 32 + 9) {
			const uint32_t c = tid - 32;
			sh.rateTab[c] = fmaxf(sieve_rate(nLists, max(c, 1u), dens), 1e-4f);
		}
		if (tid < MAX_QGRAMS) sh.qcode[tid] = (uint32_t)tid < nGrams ? d.qcode[tid] : ~0ull;
		const bool classMode = ix.identity && ix.uniformW;
		if (tid <= MAX_QGRAMS) {
			const uint32_t c = tid;
			const float s = (c <= nGrams && nGrams) ? term_score(c, nGrams) : 0.0f;
			sh.sOf[c] = s;
			sh.geHist[c] = 0;
			if (classMode) {
				sh.exactKey[c] = c <= nGrams ? nGrams - c + 1 : 255u;
				sh.hiKey[c] = c >= d.cprom ? 0u : sh.exactKey[c];
			} else {
				float ub = ix.uniformW ? edge_value(1.0f, s) : edge_value(ix.wmax, s);
				sh.exactKey[c] = ~float_to_ordered(edge_value(1.0f, s));
				if (c >= d.cprom) ub = fmaxf(ub, 100.0f);
				sh.hiKey[c] = ~float_to_ordered(ub);
			}
		}
		if (tid == 0) {
			sh.bufCount = 0; sh.overflow = 0; sh.tau = ~0ull; sh.done = 0; sh.nProbe = 0;
			sh.qlen = d.qlen; sh.nGrams = nGrams; sh.cmin = d.cmin; sh.cprom = d.cprom; sh.nLists = nLists;
			sh.classMode = classMode;
			sh.promotedKey = classMode ? 0u : ~float_to_ordered(100.0f);
			sh.topShift = classMode ? 32 : 56;
			int kb = 1;
			while (kb < 32 && (1ull << kb) < ix.K) ++kb;
			sh.kidBits = kb;
			uint32_t sum = 0;
			for (uint32_t i = 0; i < nLists; ++i)
				for (uint32_t r = 0; r < d.listMult[i]; ++r) {
					if (sum < SV_MAXL) sh.sievePtr[sum] = ix.sieveBits + (size_t)d.listSlot[i] * W;
					++sum;
				}
			sh.sumMult = sum;
		}
		// Two phases on identity indexes without a threshold: phase 0 looks for terms with >= 2 matches only (tau starts behind
		// the whole two-match class), phase 1 runs only if that found fewer than k and fills the list with the smallest
		// single-match ids (bounded by the seed: the k-th posting of a list). Everything else: one phase, nothing excluded.
		const bool twoPhase = classMode && d.cmin <= 1 && nGrams >= 2;
		__syncthreads();
		const uint32_t nSieve = sh.sumMult;
		bool failed = nSieve > SV_MAXL, tauReal = false;      // (repeated grams: more counted lists than the sieve takes)
		uint32_t selectAt = min(bufCap / 2, 2 * k + 64);
		uint32_t level = 0;
		for (uint32_t phase = twoPhase ? 0u : 1u; phase < 2 && !failed; ++phase) {
			__syncthreads();
			if (phase == 1) {
				if (twoPhase && level >= k) break;
				if (classMode && d.cmin <= 1) {
					// seed (as in k_fast_search): the first k postings of one list are k distinct keys with >= 1 match
					if (warp == 0) {
						uint32_t x = INF;
						if ((uint32_t)lane < nLists && d.listLen[lane] >= k) x = __ldg(ix.postings + d.listOff[lane] + (k - 1));
						for (int o = 16; o; o >>= 1) x = min(x, __shfl_xor_sync(0xffffffffu, x, o));
						if (lane == 0) sh.tau = x == INF ? ~0ull : ((((unsigned long long)sh.exactKey[1]) << 32) | x) + 1;
					}
				}
			} else if (tid == 0) sh.tau = ((unsigned long long)sh.exactKey[1]) << 32;      // "< tau": two or more matches
			const uint32_t onlyCount = (twoPhase && phase == 1) ? 1u : 0u;
			uint32_t wLo = 0, lastW1 = 0;
			uint32_t lastCreq = 0, lastLen = 0, lastCands = 0, forceLen = 0;
			while (wLo < Wused) {
				__syncthreads();                      // the previous pass is complete (its counters and plan have been read)
				// ---- warp 0: tau from the counts of the pass just finished, the count a term needs from here on, the next pass
				if (warp == 0) {
					if (classMode && lastW1) {
						// k candidates with >= c matches are known, all of them below the end of the last pass: (class of c, end
						// of that pass) bounds the k-th best. One warp scan; the select only runs to make room in the buffer.
						const uint32_t c = lane + 1;
						uint32_t h = c <= nGrams ? sh.geHist[c] : 0u;
#pragma unroll
						for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_down_sync(0xffffffffu, h, o); if (lane + o < 32) h += t; }
						const unsigned m = __ballot_sync(0xffffffffu, h >= k);
						if (lane == 0 && m) {
							const uint32_t cb = 32u - (uint32_t)__clz(m);
							const unsigned long long hiId = min(((unsigned long long)lastW1 << 5) << shift, 0xFFFFFFFFull);
							const unsigned long long t2 = (((unsigned long long)sh.exactKey[cb]) << 32) | hiId;
							if (t2 < sh.tau) sh.tau = t2;
						}
						__syncwarp();
					}
					sh.geSnap[lane + 1] = sh.geHist[lane + 1];
					uint32_t rej = sh.cmin ? sh.cmin - 1 : 0;
					const unsigned long long tau = sh.tau;
					unsigned long long lowKid = 0;
					if (classMode) lowKid = min(((unsigned long long)wLo << 5) << shift, 0xFFFFFFFFull);
					const uint32_t c = rej + 1 + lane;
					const bool out = c <= nGrams && ((((unsigned long long)sh.hiKey[min(c, (uint32_t)MAX_QGRAMS)]) << 32) | lowKid) >= tau;
					const unsigned mOut = __ballot_sync(0xffffffffu, out);
					rej += mOut == 0xffffffffu ? 32u : (uint32_t)__ffs(~mOut) - 1;
					if (lane == 0) {
						const uint32_t creq = rej + 1;
						sh.cRej = rej;
						sh.done = creq > sh.sumMult || (onlyCount && creq > onlyCount);
						uint32_t len;
						if (forceLen) len = forceLen;
						else {
							// candidate blocks the queue should receive: its capacity, and in phases where (nearly) every block
							// yields a candidate (c = 1) no more than the candidate buffer / deferred list can take
							uint32_t want = SV_CCAP;
							if (creq <= 1) want = min(want, classMode ? bufCap - level : min(sh.defCap, bufCap - level));
							const float rate = (creq == lastCreq && lastLen) ? fmaxf(1.25f * (float)lastCands / (float)lastLen, 1e-4f) : sh.rateTab[min(creq, 8u)];
							const float l = 0.7f * (float)want / rate;
							len = l >= (float)Wused ? Wused : max(2u, (uint32_t)l);
							if (len > 8 * SV_THREADS) len &= ~(uint32_t)(2 * SV_THREADS - 1);
						}
						len = (len + 1) & ~1u;
						sh.passW0 = wLo; sh.passW1 = min(Wused, wLo + len);
						sh.nCand = 0; sh.defCount = 0;
					}
				}
				__syncthreads();
				if (sh.done) break;
				const uint32_t w0 = sh.passW0, w1 = sh.passW1, creq = sh.cRej + 1;
				if (tid == 0) atomicAdd(&counters[24], 1u);
				// ---- sieve
				sieve_dispatch(sh, creq, w0, w1, nSieve);
				__syncthreads();
				const uint32_t nCand = sh.nCand;
				bool redo = nCand > SV_CCAP;
				uint32_t shorter = redo ? (uint32_t)((float)(w1 - w0) * 0.5f * (float)SV_CCAP / (float)nCand) : 0u;
				if (!redo) {
					// ---- resolve the queued blocks, evaluate what was deferred
					resolve_stage(sh, nCand, nLists, onlyCount);
					__syncthreads();
					if (sh.overflow == 0) {
						const uint32_t nd = sh.defCount;
						for (uint32_t i = tid; i < nd; i += SV_THREADS) {
							const unsigned long long e = sh.defList[i];
							emit_candidate<U, false>(sh, qn, (uint32_t)e, (uint32_t)(e >> 32));
						}
						if (nd) __syncthreads();
					}
					if (sh.overflow == 2) { failed = true; break; }
					if (sh.overflow) {
						// the candidate buffer or the deferred list overflowed: drop what this pass appended and redo it shorter
						__syncthreads();
						if (warp == 0) {
							sh.geHist[lane + 1] = sh.geSnap[lane + 1];
							if (lane == 0) { sh.bufCount = level; sh.overflow = 0; }
						}
						redo = true;
						shorter = (w1 - w0) >> 1;
					}
				}
				if (redo) {
					if (w1 - w0 <= 2) { failed = true; break; }
					forceLen = max(2u, shorter);
					if (tid == 0) atomicAdd(&counters[27], 1u);
					continue;
				}
				forceLen = 0;
				if (tid == 0) { atomicAdd(&counters[26], nCand); }
				lastCreq = creq; lastLen = w1 - w0; lastCands = nCand + 1;
				wLo = w1; lastW1 = w1;
				level = sh.bufCount;
				// ---- make room / establish tau where the counts cannot (weights: values are not classes)
				if (classMode) {
					if (level > bufCap / 2) { select_smallest<SV_THREADS>(sh, level, k); level = k; }
				} else if (level > k && (!tauReal || level > selectAt)) {
					select_smallest<SV_THREADS>(sh, level, k); level = k; tauReal = true;
					selectAt = bufCap / 2;
				}
			}
		}
		if (failed) {
			if (tid == 0) { handList[atomicAdd(&counters[5], 1u)] = q; atomicAdd(&counters[28], 1u); }
			continue;
		}
		// ---- final: top-k, sorted (same tail as k_fast_search)
		if (level > k) { select_smallest<SV_THREADS>(sh, level, k); level = k; }
		sort_buffer<SV_THREADS>(sh, level);
		for (uint32_t i = tid; i < level; i += SV_THREADS) {
			const unsigned long long v = sh.buf[i];
			const uint32_t kid = (uint32_t)v;
			const uint32_t hi = (uint32_t)(v >> 32);
			const float score = classMode ? (hi == 0 ? 100.0f : edge_value(1.0f, sh.sOf[nGrams - hi + 1])) : rank_key_score(v);
			const size_t o = (size_t)q * outStride + i;
			outKid[o] = kid;
			outScore[o] = score;
			if (ho.kids) {
				ho.kids[o] = kid;
				ho.scores[o] = score;
				ho.rows[o] = ix.keyRow[kid];
				ho.strs[o] = ho.strBase + (uint64_t)ix.keyOff[kid] * (uint32_t)ix.unit;
			}
		}
		if (tid == 0) {
			outCount[q] = level;
			atomicAdd(&statCounters[0], (unsigned long long)d.total);
			atomicAdd(&counters[6], 1u);
			atomicAdd(&counters[25], sh.nProbe);
			if (level != outStride) atomicAdd(&counters[22], 1u);
		}
	}
}

}   // namespace
