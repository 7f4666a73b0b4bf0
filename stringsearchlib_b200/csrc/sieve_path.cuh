// sieve_path.cuh — k_sieve_search: top-k over posting lists WITHOUT reading most postings (included by search.cu after
// fast_path.cuh, whose scoring / select / sort helpers it shares).
//
// The reference increments a counter for every posting of every query gram (nGramSearch.hpp:289-298). Here the K lists of a
// query are first looked at through their block bitmaps (sieve.cu): one bit per block of 2^shift consecutive term ids. With
// tau the running k-th best (or the threshold, hpp:313-316) a term needs c = cRej + 1 matches to matter, and it can only
// have them inside a block in which at least c of the K bitmaps have their bit set. Per pass over a range of bitmap words:
//   SIEVE    every thread loads one 32-block word of each list (coalesced, K independent loads in flight) and runs the
//            bit-sliced threshold counter  ge[c] |= ge[c-1] & x  (c LOP3 per list): ge[c-1] = blocks hit by >= c lists.
//            Blocks that pass are queued together with one PROBE per list that has its bit set (warp-aggregated slots).
//   PROBE    one thread per (block, list): rank of the block inside its 256-block group (popcount of the lower bits, two
//            16-byte loads of the bitmap the sieve has just streamed) + the group's directory entry = position of the
//            block's first posting; the posting is read from the bit-exact CSR postings. A second posting of the same list
//            in the same block (rare) goes to a small side list.
//   COMBINE  one thread per block: exact match count of every term found (the reference's count, with multiplicities),
//            the integer pre-filter against tau, then the same emit / deferred-evaluation code as k_fast_search.
// Passes run in ascending term-id order; tau moves from class to class through bounds that cost nothing (k candidates with
// >= c matches are known, all below the end of the pass), so that c rises and the sieve gets sparser as the query proceeds.
// Bytes read per query: K bitmaps (~ the bytes of the K lists at the chosen block size) + ~100 B per probe, against
// 4 bytes + ~150 instructions per posting for the merge. Exactness does not depend on the bitmaps being selective: a block
// that passes is resolved from the postings themselves.
// Queries the kernel does not take (a list without bitmap, more than 16 lists, the edit-distance path) or cannot finish
// (buffer overflow on indexes with many edges per term) are handed to k_fast_search through handList.
#pragma once

namespace {

constexpr int SV_THREADS = 128;
constexpr int SV_QCAP = 1024;                           // candidate blocks / probes per pass
constexpr int SV_XCAP = 64;                             // extra postings (second posting of one list in one block) per pass

struct SieveShared {
	const uint32_t* bitsPtr[SV_MAXL];                   // bitmap of list j
	const uint32_t* cumPtr[SV_MAXL];                    // its group directory
	uint32_t listEnd[SV_MAXL], mult[SV_MAXL];
	uint32_t candBlk[SV_QCAP];                          // block number of candidate i
	uint32_t candInfo[SV_QCAP];                         // first probe | number of probes << 16
	uint32_t probe[SV_QCAP];                            // candidate << 4 | list
	uint32_t hit[SV_QCAP];                              // term id found by the probe (INF: none)
	uint32_t xEnt[SV_XCAP], xId[SV_XCAP];               // extra postings: candidate << 4 | list, term id
	unsigned long long* buf;                            // candidate keys: bufCap entries behind this struct in dynamic smem
	uint32_t bufCap;
	uint32_t defCount, defCap;
	unsigned long long* defList;
	union {
		unsigned long long deferred[FS_DEF_CAP];
		uint32_t hist[256];
	};
	uint32_t geHist[MAX_QGRAMS + 1];                    // candidates appended with exactly c matches (class mode)
	uint64_t qcode[MAX_QGRAMS];
	uint32_t hiKey[MAX_QGRAMS + 1], exactKey[MAX_QGRAMS + 1];
	float sOf[MAX_QGRAMS + 1];
	unsigned long long tau;
	uint32_t promotedKey;
	uint32_t bufCount, overflow, query, selPrefixLo, selK;
	uint32_t nCand, nProbe, nExtra;
	uint32_t passW0, passW1, passCap, cRej, done;
	uint32_t qlen, nGrams, cmin, cprom, nLists, sumMult, allOnes;
	uint32_t shortActive, scanLong, scanLongOnly;       // read by the shared scoring helpers; always 0 here
	float threshold;
	int topShift, kidBits;
	uint32_t classMode;
	IndexView ix;
};

// ---- SIEVE
template <int CREQ, bool ONES>
__device__ __forceinline__ void sieve_words(SieveShared& sh, const uint32_t w0, const uint32_t w1, const uint32_t nL, const uint32_t cap)
{
	const int tid = threadIdx.x, lane = tid & 31;
	for (uint32_t base = w0; base < w1; base += SV_THREADS) {
		const uint32_t w = base + tid;
		const bool act = w < w1;
		uint32_t x[SV_MAXL];
#pragma unroll
		for (int j = 0; j < SV_MAXL; ++j) {
			x[j] = 0;
			if (act && (uint32_t)j < nL) x[j] = __ldg(sh.bitsPtr[j] + w);
		}
		uint32_t ge[CREQ];
#pragma unroll
		for (int c = 0; c < CREQ; ++c) ge[c] = 0;
#pragma unroll
		for (int j = 0; j < SV_MAXL; ++j) {
			const uint32_t xj = x[j];
			if (ONES) {
#pragma unroll
				for (int c = CREQ - 1; c >= 1; --c) ge[c] |= ge[c - 1] & xj;
				ge[0] |= xj;
			} else {
				const uint32_t m = sh.mult[j];                   // 0 for the slots beyond nL
				for (uint32_t r = 0; r < m; ++r) {
#pragma unroll
					for (int c = CREQ - 1; c >= 1; --c) ge[c] |= ge[c - 1] & xj;
					ge[0] |= xj;
				}
			}
		}
		uint32_t mask = ge[CREQ - 1];
		if (!__any_sync(0xffffffffu, mask != 0)) continue;
		// queue slots: candidates in the low half, probes in the high half of one packed warp scan
		uint32_t nc = __popc(mask), np = 0;
#pragma unroll
		for (int j = 0; j < SV_MAXL; ++j) np += __popc(x[j] & mask);
		uint32_t inc = nc | (np << 16);
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
		const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
		uint32_t cBase = 0, pBase = 0;
		if (lane == 31) { cBase = atomicAdd(&sh.nCand, tot & 0xffffu); pBase = atomicAdd(&sh.nProbe, tot >> 16); }
		cBase = __shfl_sync(0xffffffffu, cBase, 31);
		pBase = __shfl_sync(0xffffffffu, pBase, 31);
		if (cBase + (tot & 0xffffu) > cap || pBase + (tot >> 16) > cap) continue;       // pass too long: counted, redone shorter
		uint32_t ci = cBase + (inc & 0xffffu) - nc, pi = pBase + (inc >> 16) - np;
		while (mask) {
			const uint32_t b = __ffs(mask) - 1;
			mask &= mask - 1;
			sh.candBlk[ci] = (w << 5) | b;
			const uint32_t p0 = pi;
#pragma unroll
			for (int j = 0; j < SV_MAXL; ++j)
				if ((x[j] >> b) & 1u) sh.probe[pi++] = (ci << 4) | (uint32_t)j;
			sh.candInfo[ci] = p0 | ((pi - p0) << 16);
			++ci;
		}
	}
}

template <bool ONES>
__device__ __forceinline__ void sieve_dispatch(SieveShared& sh, uint32_t creq, uint32_t w0, uint32_t w1, uint32_t nL, uint32_t cap)
{
	switch (creq) {
	case 0: case 1: sieve_words<1, ONES>(sh, w0, w1, nL, cap); break;
	case 2: sieve_words<2, ONES>(sh, w0, w1, nL, cap); break;
	case 3: sieve_words<3, ONES>(sh, w0, w1, nL, cap); break;
	case 4: sieve_words<4, ONES>(sh, w0, w1, nL, cap); break;
	case 5: sieve_words<5, ONES>(sh, w0, w1, nL, cap); break;
	case 6: sieve_words<6, ONES>(sh, w0, w1, nL, cap); break;
	case 7: sieve_words<7, ONES>(sh, w0, w1, nL, cap); break;
	default: sieve_words<8, ONES>(sh, w0, w1, nL, cap); break;     // c > 8: a superset of the blocks; COMBINE counts exactly
	}
}

// ---- PROBE: one thread per (candidate block, list with its bit set)
__device__ __forceinline__ void probe_stage(SieveShared& sh, uint32_t nProbe)
{
	const int shift = sh.ix.sieveShift;
	const uint32_t* __restrict__ postings = sh.ix.postings;
	for (uint32_t p = threadIdx.x; p < nProbe; p += SV_THREADS) {
		const uint32_t e = sh.probe[p];
		const uint32_t j = e & 15u, ci = e >> 4;
		const uint32_t blk = sh.candBlk[ci];
		const uint32_t w = blk >> 5, grp = w >> 3, wi = w & 7u, bit = blk & 31u;
		const uint4* gp = reinterpret_cast<const uint4*>(sh.bitsPtr[j] + (grp << 3));
		uint32_t idx = __ldg(sh.cumPtr[j] + grp);
		const uint4 a = __ldg(gp), b4 = __ldg(gp + 1);
		const uint32_t ws[8] = {a.x, a.y, a.z, a.w, b4.x, b4.y, b4.z, b4.w};
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			const uint32_t m = (uint32_t)i < wi ? 0xFFFFFFFFu : ((uint32_t)i == wi ? ((1u << bit) - 1u) : 0u);
			idx += __popc(ws[i] & m);
		}
		// idx = group start + occupied blocks below: exact unless one of those blocks holds several postings -> step forward
		const uint32_t end = sh.listEnd[j];
		uint32_t id = INF;
		while (idx < end) {
			id = __ldg(postings + idx);
			if ((id >> shift) >= blk) break;
			++idx;
		}
		sh.hit[p] = (idx < end && (id >> shift) == blk) ? id : INF;
		for (++idx; idx < end; ++idx) {                  // further postings of this list in the same block
			const uint32_t id2 = __ldg(postings + idx);
			if ((id2 >> shift) != blk) break;
			const uint32_t xi = atomicAdd(&sh.nExtra, 1u);
			if (xi < SV_XCAP) { sh.xEnt[xi] = e; sh.xId[xi] = id2; }
		}
	}
}

// a term with exact match count c > cRej (same tail as merge_pass)
__device__ __forceinline__ void sieve_emit(SieveShared& sh, uint32_t term, uint32_t c, bool identMode, unsigned long long tau)
{
	if (identMode && c < sh.cprom) {
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

// ---- COMBINE: one thread per candidate block
__device__ __forceinline__ void combine_stage(SieveShared& sh, uint32_t nCand, uint32_t nExtra)
{
	const bool identMode = sh.classMode != 0;
	const unsigned long long tau = sh.tau;
	const uint32_t nG = sh.nGrams;
	const int shift = sh.ix.sieveShift;
	for (uint32_t ci = threadIdx.x; ci < nCand; ci += SV_THREADS) {
		const uint32_t info = sh.candInfo[ci];
		const uint32_t pb = info & 0xffffu, n = info >> 16;
		// integer pre-filter refined by the block's first id (identity mode: kid == term id >= first id of the block)
		uint32_t cRej = sh.cRej;
		if (identMode) {
			const unsigned long long lo = (unsigned long long)sh.candBlk[ci] << shift;
			while (cRej < nG && ((((unsigned long long)sh.hiKey[cRej + 1]) << 32) | lo) >= tau) ++cRej;
		}
		if (nExtra == 0) {
			for (uint32_t a = 0; a < n; ++a) {
				const uint32_t v = sh.hit[pb + a];
				if (v == INF) continue;
				uint32_t c = 0;
				bool first = true;
				for (uint32_t b = 0; b < n; ++b)
					if (sh.hit[pb + b] == v) { if (b < a) first = false; c += sh.mult[sh.probe[pb + b] & 15u]; }
				if (first && c > cRej) sieve_emit(sh, v, c, identMode, tau);
			}
		} else {
			// rare: some list has two or more postings in one block of this pass
			uint32_t ids[2 * SV_MAXL], mu[2 * SV_MAXL], m = 0;
			for (uint32_t a = 0; a < n; ++a)
				if (sh.hit[pb + a] != INF) { ids[m] = sh.hit[pb + a]; mu[m] = sh.mult[sh.probe[pb + a] & 15u]; ++m; }
			for (uint32_t x = 0; x < nExtra; ++x)
				if ((sh.xEnt[x] >> 4) == ci) {
					if (m < 2 * SV_MAXL) { ids[m] = sh.xId[x]; mu[m] = sh.mult[sh.xEnt[x] & 15u]; ++m; } else sh.overflow = 1;
				}
			for (uint32_t a = 0; a < m; ++a) {
				uint32_t c = 0;
				bool first = true;
				for (uint32_t b = 0; b < m; ++b)
					if (ids[b] == ids[a]) { if (b < a) first = false; c += mu[b]; }
				if (first && c > cRej) sieve_emit(sh, ids[a], c, identMode, tau);
			}
		}
	}
}

// expected probes per bitmap word when a block needs >= c of K lists, every list setting a block with probability p
__device__ float sieve_rate(uint32_t K, uint32_t c, float p)
{
	if (p >= 0.999f) return 32.0f * (float)K;
	float pmf = powf(1.0f - p, (float)K), e = 0.0f;         // pmf(0)
	const float r = p / (1.0f - p);
	for (uint32_t m = 0; m < K; ++m) {
		pmf = pmf * (float)(K - m) / (float)(m + 1) * r;     // pmf(m + 1)
		if (m + 1 >= c) e += (float)(m + 1) * pmf;
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
	const uint32_t Wused = (nBlocks + 31) >> 5;
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
		if (tid < SV_MAXL) {
			if ((uint32_t)tid < nLists) {
				const uint32_t slot = d.listSlot[tid];
				sh.bitsPtr[tid] = ix.sieveBits + (size_t)slot * W;
				sh.cumPtr[tid] = ix.sieveCum + (size_t)slot * nGroups1;
				sh.listEnd[tid] = d.listOff[tid] + d.listLen[tid];
				sh.mult[tid] = d.listMult[tid];
			} else { sh.bitsPtr[tid] = ix.sieveBits; sh.cumPtr[tid] = ix.sieveCum; sh.listEnd[tid] = 0; sh.mult[tid] = 0; }
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
			sh.bufCount = 0; sh.overflow = 0; sh.tau = ~0ull; sh.done = 0;
			sh.qlen = d.qlen; sh.nGrams = nGrams; sh.cmin = d.cmin; sh.cprom = d.cprom; sh.nLists = nLists;
			sh.classMode = classMode;
			sh.promotedKey = classMode ? 0u : ~float_to_ordered(100.0f);
			sh.topShift = classMode ? 32 : 56;
			int kb = 1;
			while (kb < 32 && (1ull << kb) < ix.K) ++kb;
			sh.kidBits = kb;
			uint32_t ones = 1, sum = 0;
			for (uint32_t i = 0; i < nLists; ++i) { ones &= d.listMult[i] == 1; sum += d.listMult[i]; }
			sh.allOnes = ones; sh.sumMult = sum;
		}
		__syncthreads();
		if (classMode && d.cmin <= 1 && k <= FAST_KLIM) {
			// seed (as in k_fast_search): the first k postings of one list are k distinct keys with >= 1 match
			if (warp == 0) {
				uint32_t x = INF;
				if ((uint32_t)lane < nLists && d.listLen[lane] >= k) x = __ldg(ix.postings + d.listOff[lane] + (k - 1));
				for (int o = 16; o; o >>= 1) x = min(x, __shfl_xor_sync(0xffffffffu, x, o));
				if (lane == 0 && x != INF) sh.tau = ((((unsigned long long)sh.exactKey[1]) << 32) | x) + 1;
			}
			__syncthreads();
		}
		// average bit density of a list's bitmap (pass-length estimate only)
		const float dens = fminf(1.0f, (float)d.total / ((float)nLists * (float)nBlocks));

		bool failed = false, tauReal = false;
		uint32_t selectAt = min(bufCap / 2, 2 * k + 64);
		uint32_t level = 0, wLo = 0;
		uint32_t lastCreq = 0, lastLen = 0, lastProbes = 0, forceLen = 0;
		while (wLo < Wused) {
			__syncthreads();                      // everybody has read the previous pass's counters and plan
			// ---- plan: the count a term needs from here on, the length of the pass
			if (warp == 0) {
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
					sh.done = creq > sh.sumMult;
					// every probe can append one candidate (class mode) or defer one term (indexes with edges / weights)
					uint32_t cap = min((uint32_t)SV_QCAP, bufCap - level);
					if (!classMode) cap = min(cap, sh.defCap);
					uint32_t len;
					if (forceLen) len = forceLen;
					else {
						float rate = (creq == lastCreq && lastLen) ? 1.25f * (float)lastProbes / (float)lastLen : sieve_rate(nLists, creq, dens);
						rate = fmaxf(rate, 1e-3f);
						const float want = 0.6f * (float)cap / rate;
						len = want >= (float)Wused ? Wused : max(1u, (uint32_t)want);
						if (len > SV_THREADS) len &= ~(uint32_t)(SV_THREADS - 1);
					}
					sh.passW0 = wLo; sh.passW1 = min(Wused, wLo + len); sh.passCap = cap;
					sh.nCand = 0; sh.nProbe = 0; sh.nExtra = 0; sh.defCount = 0;
				}
			}
			__syncthreads();
			if (sh.done) break;
			const uint32_t w0 = sh.passW0, w1 = sh.passW1, cap = sh.passCap, creq = sh.cRej + 1;
			if (tid == 0) atomicAdd(&counters[24], 1u);
			// ---- sieve
			if (sh.allOnes) sieve_dispatch<true>(sh, creq, w0, w1, nLists, cap);
			else sieve_dispatch<false>(sh, creq, w0, w1, nLists, cap);
			__syncthreads();
			const uint32_t nCand = sh.nCand, nProbe = sh.nProbe;
			if (nCand > cap || nProbe > cap) {
				// more candidates than the queues hold: redo the pass over a shorter range (nothing else was written)
				if (w1 - w0 <= 1) { failed = true; break; }
				const float f = 0.5f * (float)cap / (float)max(nCand, nProbe);
				forceLen = max(1u, (uint32_t)((float)(w1 - w0) * f));
				if (tid == 0) atomicAdd(&counters[27], 1u);
				continue;
			}
			forceLen = 0;
			// ---- probe, combine, deferred evaluation
			probe_stage(sh, nProbe);
			__syncthreads();
			const uint32_t nExtra = sh.nExtra;
			if (nExtra > SV_XCAP || nProbe + nExtra > cap) {
				if (w1 - w0 <= 1) { failed = true; break; }
				forceLen = max(1u, (w1 - w0) >> 1);
				if (tid == 0) atomicAdd(&counters[27], 1u);
				continue;
			}
			combine_stage(sh, nCand, nExtra);
			__syncthreads();
			{
				const uint32_t nd = min(sh.defCount, sh.defCap);
				for (uint32_t i = tid; i < nd; i += SV_THREADS) {
					const unsigned long long e = sh.defList[i];
					emit_candidate<U, false>(sh, qn, (uint32_t)e, (uint32_t)(e >> 32));
				}
				if (nd) __syncthreads();
			}
			if (sh.overflow) { failed = true; break; }
			if (tid == 0) { atomicAdd(&counters[25], nProbe); atomicAdd(&counters[26], nCand); }
			lastCreq = creq; lastLen = w1 - w0; lastProbes = nProbe + 1;
			wLo = w1;
			level = sh.bufCount;
			// ---- tau
			if (classMode) {
				// k candidates with >= c matches are known, all of them below the end of this pass: (class of c, end of the
				// pass) bounds the k-th best. Costs one warp scan; the select only runs to make room in the buffer.
				if (warp == 0) {
					const uint32_t c = lane + 1;
					uint32_t h = c <= nGrams ? sh.geHist[c] : 0u;
#pragma unroll
					for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_down_sync(0xffffffffu, h, o); if (lane + o < 32) h += t; }
					const unsigned m = __ballot_sync(0xffffffffu, h >= k);
					if (lane == 0 && m) {
						const uint32_t cb = 32u - (uint32_t)__clz(m);
						const unsigned long long hiId = min(((unsigned long long)w1 << 5) << shift, 0xFFFFFFFFull);
						const unsigned long long t2 = (((unsigned long long)sh.exactKey[cb]) << 32) | hiId;
						if (t2 < sh.tau) sh.tau = t2;
					}
				}
				__syncthreads();
				if (level > bufCap / 2) { select_smallest<SV_THREADS>(sh, level, k); level = k; }
			} else if (level > k && (!tauReal || level > selectAt)) {
				select_smallest<SV_THREADS>(sh, level, k); level = k; tauReal = true;
				selectAt = bufCap / 2;
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
			if (level != outStride) atomicAdd(&counters[22], 1u);
		}
	}
}

}   // namespace
