// sieve_path.cuh — k_sieve_search: top-k over posting lists WITHOUT reading most postings (included by search.cu after
// fast_path.cuh, whose scoring / select / sort helpers it shares).
//
// The reference increments a counter for every posting of every query gram (nGramSearch.hpp:289-298). Here the K lists of a
// query are first looked at through their block bitmaps (sieve.cu): one bit per block of 2^shift consecutive term ids. With
// tau the running k-th best (or the threshold, hpp:313-316) a term needs c = cRej + 1 matches to matter, and it can only
// have them inside a block in which at least c of the K bitmaps have their bit set. Per pass over a range of bitmap words:
//   SIEVE    every lane streams two 96-block groups of each list (one coalesced 16-byte load per group: first posting + block
//            bits) through the bit-sliced threshold counter  ge[c] |= ge[c-1] & x  (c LOP3 per word and list): ge[c-1] = blocks
//            hit by >= c lists. The numbers of the blocks that pass are queued (one warp scan per step, a queue per warp).
//   RESOLVE  one thread per queued block, a warp of blocks at a time: which lists hit it (their words are still in L1/L2), then
//            for up to four lists the block's rank among the occupied blocks of its group (popcount of the lower bits).
//            INL (32-byte groups, index.h): the rank selects the group's slot = the offset of the list's posting inside the
//            block, so block number + offset is the term id and no posting is read. !INL (16-byte groups): group's first
//            posting + rank = position of the block's first posting in the bit-exact CSR postings, three of which are read.
//            The terms found are counted in registers — the reference's match count, with multiplicities — filtered by the
//            integer bound derived from tau and handed to the same emit / deferred-evaluation code as k_fast_search. Blocks
//            in which a list has several postings, or that more than four lists hit, go to the general resolution
//            (resolve_slow), which walks the postings.
// Passes run in ascending term-id order; tau moves from class to class through bounds that cost nothing (k candidates with
// >= c matches are known, all below the end of the pass), so that c rises and the sieve gets sparser as the query proceeds.
// On identity indexes without threshold the terms with ONE match are left for last: the first phase looks for terms with
// two or more matches only (no seeding pass over every posting of the smallest ids), and only if fewer than k exist does a
// second phase fill the list with the smallest single-match ids.
// Bytes read per query: K bitmaps (16 or 32 bytes per 96 blocks: 1.3x / 2.5x the bytes of the K lists at the chosen block size;
// + one 32-byte sector per (block, list) resolved when !INL), against 4 bytes + ~150 instructions per posting for the merge.
// Exactness does not depend on the bitmaps being selective: a block that passes is resolved from the postings themselves or
// from slots that restate them exactly (sieve.cu), and a disagreement between bits and postings hands the query over.
// Queries the kernel does not take (a list without bitmap, more than 16 lists, the edit-distance path) or cannot finish
// (candidate buffer too small for one 64-block step, more than 8 distinct terms of interest in one block) are handed to
// k_fast_search through handList.
#pragma once

namespace {

// -DSV_TIMING=1: thread 0 of every CTA accumulates the cycles (>> 6) it spends per phase of a query into counters[32..39]
// (0 set-up, 1 plan, 2 sieve + resolve, 3 slow blocks + deferred, 4 select in the loop, 5 final select + sort + output)
#ifndef SV_TIMING
#define SV_TIMING 0
#endif
#if SV_TIMING
#define SV_TICK(slot) do { if (tid == 0) { const long long t_ = clock64(); atomicAdd(&counters[32 + (slot)], (uint32_t)((t_ - tick) >> 6)); tick = t_; } } while (0)
#else
#define SV_TICK(slot) do { } while (0)
#endif

#ifndef SV_MIN_CTAS
#define SV_MIN_CTAS 8
#endif
#ifndef SV_THREADS_
#define SV_THREADS_ 128
#endif
constexpr int SV_THREADS = SV_THREADS_;
#ifndef SV_DYNAMIC
#define SV_DYNAMIC 1
#endif
#ifndef SV_CCAP_
#define SV_CCAP_ 2048
#endif
constexpr int SV_CCAP = SV_CCAP_;                           // candidate blocks a pass aims at (its length is chosen accordingly)
constexpr int SV_WQCAP = 160;                           // per-warp queue of blocks that passed the sieve: < 32 left over + <= 128 pushed
constexpr int SV_PU = 4;                                // lists of one block resolved at a time (independent load chains)
// -DSV_DYN16=1: a dynamic pass over the 32-byte layout streams the 16-byte one behind its first rise (half the bytes, but the
// blocks that pass read their postings again). Measured slower: 0.71 vs 0.60 ms per batch at 10M keys — the scattered reads cost
// more than the stream saves even at three lists.
#ifndef SV_DYN16
#define SV_DYN16 0
#endif
#ifndef SV_DYN_LEVELS_
#define SV_DYN_LEVELS_ 2
#endif
constexpr int SV_DYN_LEVELS = SV_DYN_LEVELS_;           // a dynamic pass rises from two lists to three, to four (and to five) on its own
constexpr int SV_SLOWCAP = 512;                         // blocks per pass that need the general resolution

struct SieveShared {
	const uint4* sievePtr[SV_MAXL];                     // group arrays as the sieve counts them: list j repeated mult[j] times
	const uint4* grpPtr[SV_MAXL];                       // group array of list j: { first posting, 3 x 32 block bits } per 96 blocks
	const uint4* sievePtr16[SV_MAXL], *grpPtr16[SV_MAXL];      // the same in the 16-byte layout (k_sieve_search<U, false>: equal to the two above)
	uint32_t listEnd[SV_MAXL], mult[SV_MAXL];
	uint32_t wq[SV_THREADS / 32][SV_WQCAP];             // numbers of the blocks that passed the sieve, per warp
	uint32_t slowQ[SV_SLOWCAP];                         // of those, the ones left to resolve_slow
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
	uint32_t nCand, nProbe, nSlow;
	uint32_t dyn, kWant, maxPos;                        // dynamic pass (sieve_groups<2, true>): see there
	uint32_t nGe[SV_DYN_LEVELS], switchGrp[SV_DYN_LEVELS];      // candidates with >= 2 + i matches; from which group on 3 + i are needed
	uint32_t tim[9];                                    // SV_TIMING: warp 0's cycles >> 6 in the sieve loads, the queue pushes, the resolve rounds, ...
	uint32_t passG0, passG1, cRej, done;
	uint32_t qlen, nGrams, cmin, cprom, nLists, sumMult;
	uint32_t shortActive, scanLong, scanLongOnly;       // read by the shared scoring helpers; always 0 here
	float threshold;
	int topShift, kidBits;
	uint32_t classMode;
	IndexView ix;
};

// a term with exact match count c > cRej (same tail as merge_pass); emits are rare -> one copy of the code
__device__ __noinline__ void sieve_emit(SieveShared& sh, uint32_t term, uint32_t c, unsigned long long tau)
{
	if (sh.classMode && c < sh.cprom) {
		const unsigned long long key = (((unsigned long long)sh.exactKey[c]) << 32) | term;
		if (key < tau) {
			append_key(sh, key);
			atomicAdd(&sh.geHist[c], 1u);
			if (sh.dyn) {
				// the k-th candidate with cc or more matches of a dynamic pass: all k lie below the end of the furthest chunk any
				// warp has started, so from there on a term needs cc + 1 matches (chunks further down sieve for cc + 1 lists)
				for (uint32_t cc = 2; cc <= min(c, 1u + SV_DYN_LEVELS); ++cc)
					if (atomicAdd(&sh.nGe[cc - 2], 1u) + 1 == sh.kWant) {
						const uint32_t x = *reinterpret_cast<volatile uint32_t*>(&sh.maxPos);
						const unsigned long long xid = min(((unsigned long long)x * SIEVE_GROUP_BLOCKS) << sh.ix.sieveShift, 0xFFFFFFFFull);
						atomicMin(&sh.tau, (((unsigned long long)sh.exactKey[cc]) << 32) | xid);
						*reinterpret_cast<volatile uint32_t*>(&sh.switchGrp[cc - 2]) = x;
					}
			}
		}
	} else {
		const uint32_t di = atomicAdd(&sh.defCount, 1u);
		if (di < sh.defCap) sh.defList[di] = (((unsigned long long)c) << 32) | term; else sh.overflow = 1;
	}
}

// group x of a list, block = word k / bit of the group: is the block's bit set, and where the block's first posting is —
// the group's first posting + the occupied blocks below (exact unless one of those holds several postings)
__device__ __forceinline__ bool sieve_locate(const uint4 x, uint32_t k, uint32_t bit, uint32_t& idx)
{
	const uint32_t wk = k == 0 ? x.y : (k == 1 ? x.z : x.w);
	idx = x.x + (k > 0 ? __popc(x.y) : 0) + (k > 1 ? __popc(x.z) : 0) + __popc(wk & ((1u << bit) - 1u));
	return (wk >> bit) & 1u;
}

// general resolution of one block (more than SV_PU lists hit it, or a list has several postings in or just below it):
// local arrays, counts by comparing all pairs. Rare, and run for a compacted list of such blocks after the pass.
template <bool INL>
__device__ __noinline__ void resolve_slow(SieveShared& sh, uint32_t blk, uint32_t onlyCount, unsigned long long tau)
{
	constexpr int CAP = 48;
	constexpr uint32_t GU = INL ? 2 : 1;                     // 16-byte units per group
	const uint4* const* const gp = INL ? sh.grpPtr : sh.grpPtr16;
	const int shift = sh.ix.sieveShift;
	const uint32_t* __restrict__ postings = sh.ix.postings;
	const uint32_t v = blk >> 5, g = v / 3u, k = v - 3u * g, bit = blk & 31u;
	const uint32_t nL = sh.nLists;
	uint32_t cRej = sh.cRej;
	if (sh.classMode) {
		const unsigned long long lo = (unsigned long long)blk << shift;
		while (cRej < sh.nGrams && ((((unsigned long long)sh.hiKey[cRej + 1]) << 32) | lo) >= tau) ++cRej;
	}
	// the lists that hit the block (word k of every list's group: independent loads), then per list the postings inside the
	// block, four in flight at a time
	uint32_t lm = 0;
#if SV_TIMING
	long long ts = clock64();
#define SV_TICKS(slot) do { if (threadIdx.x == 0) { const long long t_ = clock64(); sh.tim[slot] += (uint32_t)((t_ - ts) >> 6); ts = t_; } } while (0)
#else
#define SV_TICKS(slot) do { } while (0)
#endif
#pragma unroll 4
	for (uint32_t j = 0; j < nL; ++j) lm |= ((__ldg(reinterpret_cast<const uint32_t*>(gp[j] + GU * g) + 1 + k) >> bit) & 1u) << j;
	unsigned long long ent[CAP];                             // (multiplicity, term id) of every posting inside the block
	uint32_t m = 0;
	SV_TICKS(5);
	if (SV_TIMING && threadIdx.x == 0) sh.tim[8] += __popc(lm);
	const uint32_t low = (1u << bit) - 1u;
	const uint32_t my = k == 0 ? low : 0xFFFFFFFFu, mz = k == 1 ? low : (k > 1 ? 0xFFFFFFFFu : 0u), mw = k == 2 ? low : 0u;
	const uint32_t bps = sieve_slot_bits(shift), nSlot = sieve_slots_per_group(shift);
	while (lm) {
		const uint32_t j = (uint32_t)__ffs(lm) - 1u;
		lm &= lm - 1;
		const uint4 x = __ldg(gp[j] + GU * g);
		const uint32_t end = sh.listEnd[j], mu = sh.mult[j];
		// a list with one posting in the block: its slot of the same sector says which term (no posting read)
		const uint32_t r = __popc(x.y & my) + __popc(x.z & mz) + __popc(x.w & mw);
		if (INL && r < nSlot) {
			const uint32_t* sp = reinterpret_cast<const uint32_t*>(gp[j] + GU * g + 1);
			const uint32_t wd = (r * bps) >> 5;
			const uint32_t s0 = __ldg(sp + wd), s1 = wd < 3 ? __ldg(sp + wd + 1) : 0u;
			const uint32_t v = __funnelshift_r(s0, s1, (r * bps) & 31u) & ((1u << bps) - 1u);
			if (!(v >> shift)) {
				if (m < CAP) ent[m++] = (((unsigned long long)mu) << 32) | ((blk << shift) | v); else sh.overflow = 2;
				continue;
			}
		}
		uint32_t p;
		sieve_locate(x, k, bit, p);
		bool any = false, more = true;
		while (more && p < end) {
			uint32_t id[4];
#pragma unroll
			for (int t = 0; t < 4; ++t) id[t] = p + t < end ? __ldg(postings + p + t) : INF;
#pragma unroll
			for (int t = 0; t < 4; ++t) {
				const uint32_t b = id[t] >> shift;
				if (id[t] == INF || b > blk) { more = false; break; }
				if (b < blk) continue;                       // extra postings of lower blocks
				if (m < CAP) ent[m++] = (((unsigned long long)mu) << 32) | id[t]; else sh.overflow = 2;
				any = true;
			}
			p += 4;
		}
		if (!any) sh.overflow = 2;                           // bitmap and postings disagree: never trust, hand the query over
	}
	SV_TICKS(6);
	// counts by comparing all pairs; the usual handful of entries is first copied to registers (independent loads of the
	// local array instead of two dependent ones per pair)
	if (m <= 8) {
		uint32_t e[8], w[8];
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			const unsigned long long x = (uint32_t)i < m ? ent[i] : (unsigned long long)INF;
			e[i] = (uint32_t)x; w[i] = (uint32_t)(x >> 32);
		}
#pragma unroll
		for (int a2 = 0; a2 < 8; ++a2) {
			uint32_t c = 0;
			bool first = true;
#pragma unroll
			for (int b2 = 0; b2 < 8; ++b2)
				if (e[b2] == e[a2]) { c += w[b2]; if (b2 < a2) first = false; }
			if ((uint32_t)a2 < m && first && c > cRej && (!onlyCount || c == onlyCount)) sieve_emit(sh, e[a2], c, tau);
		}
	} else {
		for (uint32_t a2 = 0; a2 < m; ++a2) {
			uint32_t c = 0;
			bool first = true;
			for (uint32_t b2 = 0; b2 < m; ++b2)
				if ((uint32_t)ent[b2] == (uint32_t)ent[a2]) { if (b2 < a2) first = false; c += (uint32_t)(ent[b2] >> 32); }
			if (first && c > cRej && (!onlyCount || c == onlyCount)) sieve_emit(sh, (uint32_t)ent[a2], c, tau);
		}
	}
	SV_TICKS(7);
}

// ---- RESOLVE one block that passed the sieve (one thread per block, a full warp of blocks at a time). Returns the number
// of lists that hit it. onlyCount != 0: emit only terms with exactly that many matches (second phase: the others are in the
// buffer already).
template <bool INL>
__device__ __noinline__ uint32_t resolve_block(SieveShared& sh, const uint32_t blk, const uint32_t onlyCount, const unsigned long long tau)
{
	constexpr uint32_t GU = INL ? 2 : 1;
	const uint4* const* const gp = INL ? sh.grpPtr : sh.grpPtr16;
	const int shift = sh.ix.sieveShift;
	const uint32_t* __restrict__ postings = sh.ix.postings;
	const uint32_t nL = sh.nLists;
	const uint32_t v = blk >> 5, g = v / 3u, k = v - 3u * g, bit = blk & 31u;
	// Three waits per block, each for independent loads: (1) word k of every list's group (streamed by this warp moments ago:
	// L1 / L2 hits) -> which lists hit the block; (2) the whole group of the first SV_PU of them -> where the block's first
	// posting is: the group's first posting + the occupied blocks below (exact unless one of those holds several); (3) postings.
	uint32_t lm = 0;
	{
		uint32_t wv[SV_MAXL];
#pragma unroll
		for (int j = 0; j < SV_MAXL; ++j)
			wv[j] = (uint32_t)j < nL ? __ldg(reinterpret_cast<const uint32_t*>(gp[j] + GU * g) + 1 + k) : 0u;
#pragma unroll
		for (int j = 0; j < SV_MAXL; ++j) lm |= ((wv[j] >> bit) & 1u) << j;
	}
	const uint32_t ns = __popc(lm);
	const uint32_t low = (1u << bit) - 1u;
	const uint32_t my = k == 0 ? low : 0xFFFFFFFFu, mz = k == 1 ? low : (k > 1 ? 0xFFFFFFFFu : 0u), mw = k == 2 ? low : 0u;
	bool slow = ns > SV_PU;
	uint32_t h[SV_PU], mu[SV_PU];
	if constexpr (INL) {
		// (2) the group of the first SV_PU lists that hit: the block's rank among the group's occupied blocks, then (3) that slot of
		// the same 32-byte sector (it has just arrived): the offset of the list's posting inside the block — block number + offset is
		// the term id, no posting is read. A block in which the list has several postings, or whose slot did not fit, is left to
		// the general resolution.
		uint32_t rk[SV_PU], jp = 0;
		const uint32_t bps = sieve_slot_bits(shift), nSlot = sieve_slots_per_group(shift);
		{
			uint4 x[SV_PU];
			uint32_t rest = lm;
	#pragma unroll
			for (int u = 0; u < SV_PU; ++u) {
				const uint32_t j = rest ? (uint32_t)__ffs(rest) - 1u : 0u;
				x[u] = rest ? __ldg(gp[j] + GU * g) : make_uint4(0, 0, 0, 0);
				jp |= j << (4 * u);
				rest &= rest - 1;
			}
	#pragma unroll
			for (int u = 0; u < SV_PU; ++u) rk[u] = __popc(x[u].y & my) + __popc(x[u].z & mz) + __popc(x[u].w & mw);
		}
		{
			uint32_t s0[SV_PU], s1[SV_PU];
	#pragma unroll
			for (int u = 0; u < SV_PU; ++u) {
				const uint32_t j = (jp >> (4 * u)) & 15u;
				const bool on = (uint32_t)u < ns && rk[u] < nSlot;
				const uint32_t wd = (rk[u] * bps) >> 5;
				const uint32_t* sp = reinterpret_cast<const uint32_t*>(gp[j] + GU * g + 1);
				s0[u] = on ? __ldg(sp + wd) : 0u;
				s1[u] = on && wd < 3 ? __ldg(sp + wd + 1) : 0u;
				mu[u] = sh.mult[j];
			}
	#pragma unroll
			for (int u = 0; u < SV_PU; ++u) {
				const uint32_t v = __funnelshift_r(s0[u], s1[u], (rk[u] * bps) & 31u) & ((1u << bps) - 1u);
				h[u] = INF;
				if ((uint32_t)u < ns) {
					if (rk[u] < nSlot && !(v >> shift)) h[u] = (blk << shift) | v;
					else slow = true;
				}
			}
		}
	} else {
		uint32_t idx[SV_PU], jp = 0;
		{
			uint4 x[SV_PU];
			uint32_t rest = lm;
	#pragma unroll
			for (int u = 0; u < SV_PU; ++u) {
				const uint32_t j = rest ? (uint32_t)__ffs(rest) - 1u : 0u;
				x[u] = rest ? __ldg(gp[j] + GU * g) : make_uint4(0, 0, 0, 0);
				jp |= j << (4 * u);
				rest &= rest - 1;
			}
	#pragma unroll
			for (int u = 0; u < SV_PU; ++u) idx[u] = x[u].x + __popc(x[u].y & my) + __popc(x[u].z & mz) + __popc(x[u].w & mw);
		}
		// Three postings per list decide the two common cases: the first one is the block's only posting of the list, or one
		// extra posting of a lower block precedes it.
	#pragma unroll
		for (int u = 0; u < SV_PU; ++u) {
			const uint32_t j = (jp >> (4 * u)) & 15u;
			const uint32_t end = (uint32_t)u < ns ? sh.listEnd[j] : 0u;
			mu[u] = sh.mult[j];
			const uint32_t id0 = idx[u] < end ? __ldg(postings + idx[u]) : INF;
			const uint32_t id1 = idx[u] + 1 < end ? __ldg(postings + idx[u] + 1) : INF;
			const uint32_t id2 = idx[u] + 2 < end ? __ldg(postings + idx[u] + 2) : INF;
			const uint32_t b0 = id0 >> shift, b1 = id1 >> shift, b2 = id2 >> shift;
			h[u] = INF;
			if ((uint32_t)u < ns) {
				if (b0 == blk && (id1 == INF || b1 != blk)) h[u] = id0;
				else if (id0 != INF && b0 < blk && id1 != INF && b1 == blk && (id2 == INF || b2 != blk)) h[u] = id1;
				else slow = true;
			}
		}
	}
	if (slow) {
		// compacted and resolved after the pass, with every lane busy (inline, one such block would stall the whole warp)
		// (a full queue — a dynamic pass of a query with many lists — does not cost the pass: the block is resolved on the spot)
		const uint32_t si = atomicAdd(&sh.nSlow, 1u);
		if (si < SV_SLOWCAP) sh.slowQ[si] = blk; else resolve_slow<INL>(sh, blk, onlyCount, tau);
		return ns;
	}
	// integer pre-filter refined by the block's first id (identity mode: kid == term id >= first id of the block)
	uint32_t cRej = sh.cRej;
	if (sh.classMode) {
		const unsigned long long lo = (unsigned long long)blk << shift;
		while (cRej < sh.nGrams && ((((unsigned long long)sh.hiKey[cRej + 1]) << 32) | lo) >= tau) ++cRej;
	}
#pragma unroll
	for (int u = 0; u < SV_PU; ++u) {
		if ((uint32_t)u >= ns) continue;
		uint32_t c = 0;
		bool first = true;
#pragma unroll
		for (int w = 0; w < SV_PU; ++w)
			if ((uint32_t)w < ns && h[w] == h[u]) { c += mu[w]; if (w < u) first = false; }
		if (first && c > cRej && (!onlyCount || c == onlyCount)) sieve_emit(sh, h[u], c, tau);
	}
	return ns;
}

// ---- SIEVE + RESOLVE over groups [g0, g1). A list whose gram occurs m times in the query appears m times in sievePtr
// (hpp:289-298 counts once per window), so the threshold counter needs no multiplicities. Every warp streams 32 consecutive
// groups per step, collects the blocks that pass in a queue of its own and resolves them 32 at a time — while the groups it
// needs again are still in L1, without a CTA-wide barrier, and while the other warps keep streaming.
// one bitmap step of a warp: lane l streams groups gA and gB of every counted list through the bit-sliced threshold counter;
// m = the blocks that at least CREQ lists hit (3 words of group A, 3 of group B)
template <int CREQ, bool INL>
__device__ __forceinline__ void sieve_step(const SieveShared& sh, const uint32_t nS, const uint32_t gA, const uint32_t gB, const uint32_t g1, uint32_t (&m)[6])
{
	constexpr uint32_t GU = INL ? 2 : 1;
	uint32_t ga[CREQ], gb[CREQ], gc[CREQ], ha[CREQ], hb[CREQ], hc[CREQ];
#pragma unroll
	for (int c = 0; c < CREQ; ++c) { ga[c] = 0; gb[c] = 0; gc[c] = 0; ha[c] = 0; hb[c] = 0; hc[c] = 0; }
	const bool onA = gA < g1, onB = gB < g1;
#pragma unroll 4
	for (uint32_t j = 0; j < nS; ++j) {
		const uint4* ptr = (INL ? sh.sievePtr : sh.sievePtr16)[j];
		const uint4 x = onA ? __ldg(ptr + GU * gA) : make_uint4(0, 0, 0, 0);      // { first posting, 3 x 32 block bits }
		const uint4 y = onB ? __ldg(ptr + GU * gB) : make_uint4(0, 0, 0, 0);
#pragma unroll
		for (int c = CREQ - 1; c >= 1; --c) {
			ga[c] |= ga[c - 1] & x.y; gb[c] |= gb[c - 1] & x.z; gc[c] |= gc[c - 1] & x.w;
			ha[c] |= ha[c - 1] & y.y; hb[c] |= hb[c - 1] & y.z; hc[c] |= hc[c - 1] & y.w;
		}
		ga[0] |= x.y; gb[0] |= x.z; gc[0] |= x.w; ha[0] |= y.y; hb[0] |= y.z; hc[0] |= y.w;
	}
	m[0] = ga[CREQ - 1]; m[1] = gb[CREQ - 1]; m[2] = gc[CREQ - 1]; m[3] = ha[CREQ - 1]; m[4] = hb[CREQ - 1]; m[5] = hc[CREQ - 1];
}

// DYN (CREQ == 2, the first phase of a query without threshold): the pass runs to the end of the id range and switches from
// "two lists" to "three lists" (and, on indexes where even three-match terms are plentiful, on to "four") on its own. Every warp publishes the end of the chunk it starts (maxPos); the candidate that
// brings the number of terms with two or more matches to k (sieve_emit) turns maxPos into the bound — k such terms lie below
// it — and every chunk that starts at or behind it is sieved for three. One pass, and the two-list stretch is exactly as long
// as the data needs (a planned pass is either too long: candidate blocks for nothing, or too short: another pass).
template <int CREQ, bool DYN, bool INL>
__device__ __noinline__ void sieve_groups(SieveShared& sh, const uint32_t g0, const uint32_t g1, const uint32_t nS, const uint32_t onlyCount)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned long long tau = sh.tau;
	uint32_t* wq = sh.wq[warp];
	uint32_t qn = 0, nc = 0, np = 0;                             // queue fill (warp-uniform); statistics
	// a warp streams 64 consecutive groups per step, two per lane (all 2 * nS loads of a lane are independent)
	for (uint32_t base = g0 + warp * 64; base < g1; base += 2 * SV_THREADS) {
		const uint32_t gA = base + lane, gB = gA + 32;
#if SV_TIMING
		long long tk = clock64();
#define SV_TICKW(slot) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&sh.tim[slot], (uint32_t)((t_ - tk) >> 6)); tk = t_; } } while (0)
#else
#define SV_TICKW(slot) do { } while (0)
#endif
		uint32_t m[6];                                            // blocks that pass: 3 words of group A, 3 of group B
		if (DYN) {
			if (lane == 0) atomicMax(&sh.maxPos, min(base + 64, g1));
			__syncwarp();
			// (SV_DYN16: the 16-byte layout behind the first rise; default: the kernel's own layout throughout)
			if (SV_DYN_LEVELS >= 3 && base >= *reinterpret_cast<volatile uint32_t*>(&sh.switchGrp[SV_DYN_LEVELS >= 3 ? 2 : 0])) sieve_step<5, INL && !SV_DYN16>(sh, nS, gA, gB, g1, m);
			else if (base >= *reinterpret_cast<volatile uint32_t*>(&sh.switchGrp[1])) sieve_step<4, INL && !SV_DYN16>(sh, nS, gA, gB, g1, m);
			else if (base >= *reinterpret_cast<volatile uint32_t*>(&sh.switchGrp[0])) sieve_step<3, INL && !SV_DYN16>(sh, nS, gA, gB, g1, m);
			else sieve_step<2, INL>(sh, nS, gA, gB, g1, m);
		} else sieve_step<CREQ, INL>(sh, nS, gA, gB, g1, m);
		uint32_t left = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]) + __popc(m[4]) + __popc(m[5]);
		SV_TICKW(0);
		// rounds: every lane pushes up to `quota` of its blocks, then the queue is drained down to less than a warp's worth.
		// The usual case is one round that takes everything (a sparse sieve); dense phases (c = 1) go 4 blocks per lane a time.
#pragma unroll 1
		while (__any_sync(0xffffffffu, left != 0)) {
			uint32_t inc = left;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
			uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
			uint32_t cnt = left;
			if (qn + tot > SV_WQCAP) {
				cnt = min(left, 4u);
				inc = cnt;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
				tot = __shfl_sync(0xffffffffu, inc, 31);
			}
			uint32_t at = qn + inc - cnt;
			left -= cnt;
			nc += cnt;
#pragma unroll
			for (int wd = 0; wd < 6; ++wd) {
				const uint32_t b0 = (wd < 3 ? gA : gB) * SIEVE_GROUP_BLOCKS + (wd % 3) * 32;
#pragma unroll 1
				while (m[wd] && cnt) { wq[at++] = b0 + (__ffs(m[wd]) - 1); m[wd] &= m[wd] - 1; --cnt; }
			}
			qn += tot;
			__syncwarp();
			SV_TICKW(1);
#pragma unroll 1
			while (qn >= 32) {
				qn -= 32;
				const uint32_t blk = wq[qn + lane];
				__syncwarp();
				if (DYN) tau = *reinterpret_cast<volatile unsigned long long*>(&sh.tau);      // (it tightens while the pass runs)
				// (either layout resolves a block exactly; the one whose sectors the warp has just streamed is the cheap one)
				if (DYN && INL && SV_DYN16 && blk / SIEVE_GROUP_BLOCKS >= *reinterpret_cast<volatile uint32_t*>(&sh.switchGrp[0]))
					np += resolve_block<false>(sh, blk, onlyCount, tau);
				else np += resolve_block<INL>(sh, blk, onlyCount, tau);
			}
			SV_TICKW(2);
		}
	}
	if (qn) {
		const uint32_t blk = (uint32_t)lane < qn ? wq[lane] : INF;
		if (DYN) tau = *reinterpret_cast<volatile unsigned long long*>(&sh.tau);
		if (blk != INF) {
			if (DYN && INL && SV_DYN16 && blk / SIEVE_GROUP_BLOCKS >= *reinterpret_cast<volatile uint32_t*>(&sh.switchGrp[0]))
				np += resolve_block<false>(sh, blk, onlyCount, tau);
			else np += resolve_block<INL>(sh, blk, onlyCount, tau);
		}
	}
#pragma unroll
	for (int o = 16; o; o >>= 1) { nc += __shfl_xor_sync(0xffffffffu, nc, o); np += __shfl_xor_sync(0xffffffffu, np, o); }
	if (lane == 0) { atomicAdd(&sh.nCand, nc); atomicAdd(&sh.nProbe, np); }
}

template <bool INL>
__device__ __forceinline__ void sieve_dispatch(SieveShared& sh, uint32_t creq, uint32_t g0, uint32_t g1, uint32_t nS, uint32_t onlyCount)
{
	// thresholds without an instantiation of their own use the next lower one: a superset of the blocks, RESOLVE counts exactly
	if (sh.dyn) sieve_groups<2, true, INL>(sh, g0, g1, nS, onlyCount);
	else if (creq <= 1) sieve_groups<1, false, INL>(sh, g0, g1, nS, onlyCount);
	else if (creq == 2) sieve_groups<2, false, INL>(sh, g0, g1, nS, onlyCount);
	else if (creq == 3) sieve_groups<3, false, INL>(sh, g0, g1, nS, onlyCount);
	else if (creq < 6) sieve_groups<4, false, INL>(sh, g0, g1, nS, onlyCount);
	else if (creq < 8) sieve_groups<6, false, INL>(sh, g0, g1, nS, onlyCount);
	else sieve_groups<8, false, INL>(sh, g0, g1, nS, onlyCount);
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

template <class U, bool INL>
__global__ void __launch_bounds__(SV_THREADS, SV_MIN_CTAS) k_sieve_search(IndexView ix, const QDesc* __restrict__ desc, const U* __restrict__ qnorm,
                                                             const uint32_t* __restrict__ fastList, uint32_t* __restrict__ counters,
                                                             uint32_t* __restrict__ handList, uint32_t limit, uint32_t outStride,
                                                             uint32_t* __restrict__ outCount, uint32_t* __restrict__ outKid,
                                                             float* __restrict__ outScore, unsigned long long* __restrict__ statCounters,
                                                             uint32_t bufCap, float threshold, uint32_t defExt, uint32_t minMeanTotal, HostOut ho)
{
	extern __shared__ __align__(16) unsigned char sv_smem[];
	SieveShared& sh = *reinterpret_cast<SieveShared*>(sv_smem);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t nFast = counters[0];
	// batches of short queries are k_fast_search's (search.cu): the list is handed over as it is (LPT order included)
	if (*reinterpret_cast<const unsigned long long*>(counters + 30) < (unsigned long long)minMeanTotal * nFast) {
		for (uint32_t i = blockIdx.x * SV_THREADS + threadIdx.x; i < nFast; i += gridDim.x * SV_THREADS) handList[i] = fastList[i];
		if (blockIdx.x == 0 && threadIdx.x == 0) counters[5] = nFast;
		return;
	}
	const uint32_t k = limit;
	const uint32_t NG = ix.sieveNGroups;
	constexpr uint32_t GU = INL ? 2 : 1;                                   // 16-byte units per group
	const uint4* const sieveBase = INL ? ix.sieveInline : ix.sieveGroups;
	const int shift = ix.sieveShift;
	const uint32_t nBlocks = (uint32_t)((((unsigned long long)ix.T - 1) >> shift) + 1);
	if (tid == 0) {
		sh.ix = ix;
		sh.buf = reinterpret_cast<unsigned long long*>(sv_smem + ((sizeof(SieveShared) + 15) & ~size_t(15)));
		sh.bufCap = bufCap;
		sh.defList = defExt ? sh.buf + bufCap : sh.deferred;
		sh.defCap = defExt ? defExt : (uint32_t)FS_DEF_CAP;
		sh.shortActive = 0; sh.scanLong = 0; sh.scanLongOnly = 0; sh.threshold = threshold;
	}

#if SV_TIMING
	unsigned long long gt0, gtq = 0; uint32_t nDone = 0;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
#endif
	while (true) {
		__syncthreads();
		if (tid == 0) sh.query = atomicAdd(&counters[2], 1u);
		__syncthreads();
		const uint32_t wi = sh.query;
#if SV_TIMING
		if (wi >= nFast) {
			unsigned long long gt1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
			// when the CTAs run out of work: histogram of (end - start) in steps of 65.5 us in counters[48..59]
			if (tid == 0) atomicAdd(&counters[48 + min((uint32_t)((gt1 - gt0) >> 16), 11u)], 1u);
			(void)gtq; (void)nDone;
		}
#endif
		if (wi >= nFast) break;
#if SV_TIMING
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gtq)); ++nDone;
		long long tick = clock64();
		const long long tQuery = tick;
#endif
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
				sh.grpPtr[tid] = sieveBase + (size_t)d.listSlot[tid] * NG * GU;
				sh.grpPtr16[tid] = ix.sieveGroups + (size_t)d.listSlot[tid] * NG;
				sh.listEnd[tid] = d.listOff[tid] + d.listLen[tid];
				sh.mult[tid] = d.listMult[tid];
			} else { sh.grpPtr[tid] = sieveBase; sh.grpPtr16[tid] = ix.sieveGroups; sh.listEnd[tid] = 0; sh.mult[tid] = 0; }
		} else if (tid >= 32 && tid < 32 + 9) {
			const uint32_t c = tid - 32;
			sh.rateTab[c] = fmaxf(3.0f * sieve_rate(nLists, max(c, 1u), dens), 1e-4f);      // per 96-block group
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
			for (int i = 0; i < 9; ++i) sh.tim[i] = 0;
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
					if (sum < SV_MAXL) {
						sh.sievePtr[sum] = sieveBase + (size_t)d.listSlot[i] * NG * GU;
						sh.sievePtr16[sum] = ix.sieveGroups + (size_t)d.listSlot[i] * NG;
					}
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
		// Class mode: blocks for the general resolution wait in the slow queue across passes and are resolved once per phase
		// (or when the queue is half full) instead of after every pass — their postings and bitmap groups have left the caches
		// by then either way, and each resolution costs the CTA a chain of DRAM round trips. Leaving their terms out for a while
		// only makes tau looser. Nothing here can be rolled back: an overflow hands the query to k_fast_search.
		const bool lazySlow = classMode;
		auto flush_slow = [&](const uint32_t onlyCount) {
			__syncthreads();
			const uint32_t n = min(sh.nSlow, (uint32_t)SV_SLOWCAP);
			if (SV_TIMING && tid == 0) sh.tim[4] += n;
			// (the count a term needs is the query's own lower bound here, refined per block from tau: the queue holds blocks of
			// passes with different requirements)
			if (tid == 0) { sh.defCount = 0; sh.cRej = sh.cmin ? sh.cmin - 1 : 0; }
			__syncthreads();
			SV_TICK(3);
			for (uint32_t i = tid; i < n; i += SV_THREADS) resolve_slow<INL>(sh, sh.slowQ[i], onlyCount, sh.tau);
			SV_TICK(6);
			__syncthreads();
			SV_TICK(7);
			const uint32_t nd = min(sh.defCount, sh.defCap);
			for (uint32_t i = tid; i < nd; i += SV_THREADS) {
				const unsigned long long e = sh.defList[i];
				emit_candidate<U, false>(sh, qn, (uint32_t)e, (uint32_t)(e >> 32));
			}
			__syncthreads();
			if (sh.overflow) failed = true;
			level = sh.bufCount;
			__syncthreads();
			if (tid == 0) { sh.nSlow = 0; sh.defCount = 0; }
			if (!failed && level > bufCap / 2) { select_smallest<SV_THREADS>(sh, level, k); level = k; }
		};
		if (tid == 0) sh.nSlow = 0;
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
			uint32_t gLo = 0, lastG1 = 0;
			uint32_t lastCreq = 0, lastLen = 0, lastCands = 0, forceLen = 0;
			bool hadRedo = false;
			SV_TICK(0);
			while (gLo < NG) {
				__syncthreads();                      // the previous pass is complete (its counters and plan have been read)
				// ---- warp 0: tau from the counts of the pass just finished, the count a term needs from here on, the next pass
				if (warp == 0) {
					if (classMode && lastG1) {
						// k candidates with >= c matches are known, all of them below the end of the last pass: (class of c, end
						// of that pass) bounds the k-th best. One warp scan; the select only runs to make room in the buffer.
						const uint32_t c = lane + 1;
						uint32_t h = c <= nGrams ? sh.geHist[c] : 0u;
#pragma unroll
						for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_down_sync(0xffffffffu, h, o); if (lane + o < 32) h += t; }
						const unsigned m = __ballot_sync(0xffffffffu, h >= k);
						if (lane == 0 && m) {
							const uint32_t cb = 32u - (uint32_t)__clz(m);
							const unsigned long long hiId = min(((unsigned long long)lastG1 * SIEVE_GROUP_BLOCKS) << shift, 0xFFFFFFFFull);
							const unsigned long long t2 = (((unsigned long long)sh.exactKey[cb]) << 32) | hiId;
							if (t2 < sh.tau) sh.tau = t2;
						}
						__syncwarp();
					}
					sh.geSnap[lane + 1] = sh.geHist[lane + 1];
					uint32_t rej = sh.cmin ? sh.cmin - 1 : 0;
					const unsigned long long tau = sh.tau;
					unsigned long long lowKid = 0;
					if (classMode) lowKid = min(((unsigned long long)gLo * SIEVE_GROUP_BLOCKS) << shift, 0xFFFFFFFFull);
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
							len = l >= (float)NG ? NG : max(1u, (uint32_t)l);
							if (len > 8 * SV_THREADS) len &= ~(uint32_t)(SV_THREADS - 1);
						}
						// the first phase of a query without threshold runs as ONE dynamic pass to the end of the range (sieve_groups<2, true>)
						// unless a pass had to be redone (then planned passes, which can be rolled back, take over)
						const bool dyn = SV_DYNAMIC && twoPhase && phase == 0 && creq == 2 && !forceLen && !hadRedo;
						sh.dyn = dyn;
						if (dyn) {
							len = NG - gLo;
							sh.kWant = k; sh.maxPos = gLo;
							for (uint32_t lv = 0; lv < (uint32_t)SV_DYN_LEVELS; ++lv) {
								uint32_t ge = 0;
								for (uint32_t c2 = 2 + lv; c2 <= nGrams; ++c2) ge += sh.geHist[c2];
								sh.nGe[lv] = ge; sh.switchGrp[lv] = ge >= k ? gLo : 0xFFFFFFFFu;
							}
						}
						sh.passG0 = gLo; sh.passG1 = min(NG, gLo + len);
						sh.nCand = 0; sh.defCount = 0;
						if (!sh.classMode) sh.nSlow = 0;
					}
				}
				__syncthreads();
				SV_TICK(1);
				if (sh.done) break;
				const uint32_t w0 = sh.passG0, w1 = sh.passG1, creq = sh.cRej + 1;
				const uint32_t slowSnap = sh.nSlow;                // (lazy mode) entries of earlier passes
				if (tid == 0) atomicAdd(&counters[24], 1u);
				// ---- sieve + resolve, then the blocks left to the general resolution, then what was deferred
				sieve_dispatch<INL>(sh, creq, w0, w1, nSieve, onlyCount);
				__syncthreads();
				SV_TICK(2);
				const uint32_t nCand = sh.nCand;
				bool redo = false;
				uint32_t shorter = 0;
				{
					if (!lazySlow) {
						const uint32_t nSlow = min(sh.nSlow, (uint32_t)SV_SLOWCAP);
						for (uint32_t i = tid; i < nSlow; i += SV_THREADS) resolve_slow<INL>(sh, sh.slowQ[i], onlyCount, sh.tau);
						if (nSlow) __syncthreads();
					}
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
							if (lane == 0) { sh.bufCount = level; sh.overflow = 0; if (lazySlow) sh.nSlow = slowSnap; }
						}
						redo = true;
						shorter = (w1 - w0) >> 1;
					}
				}
				if (redo) {
					if (w1 - w0 <= 1) { failed = true; break; }
					forceLen = max(1u, shorter);
					hadRedo = true;
					if (tid == 0) atomicAdd(&counters[27], 1u);
					continue;
				}
				forceLen = 0;
				SV_TICK(3);
				if (tid == 0) { atomicAdd(&counters[26], nCand); sh.tim[3] += nCand; }
				lastCreq = creq; lastLen = w1 - w0; lastCands = nCand + 1;
				gLo = w1; lastG1 = w1;
				level = sh.bufCount;
				// ---- make room / establish tau where the counts cannot (weights: values are not classes)
				if (classMode) {
					if (level > bufCap / 2) { select_smallest<SV_THREADS>(sh, level, k); level = k; }
				} else if (level > k && (!tauReal || level > selectAt)) {
					select_smallest<SV_THREADS>(sh, level, k); level = k; tauReal = true;
					selectAt = bufCap / 2;
				}
				SV_TICK(4);
				if (lazySlow && sh.nSlow > (uint32_t)SV_SLOWCAP / 2) { flush_slow(onlyCount); if (failed) break; }
			}
			if (lazySlow && !failed) flush_slow(onlyCount);
			SV_TICK(3);
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
		SV_TICK(5);
#if SV_TIMING
		if (tid == 0) {
			atomicAdd(&counters[60], sh.tim[5]); atomicAdd(&counters[61], sh.tim[6]); atomicAdd(&counters[62], sh.tim[7]); atomicAdd(&counters[63], sh.tim[8]);
			atomicAdd(&counters[40], sh.tim[0]); atomicAdd(&counters[41], sh.tim[1]); atomicAdd(&counters[42], sh.tim[2]);
			// the slowest query of the batch, the number of queries over 1 and 2 megacycles, and when the last query started / ended
			const uint32_t kc = (uint32_t)((clock64() - tQuery) >> 10);
			atomicMax(&counters[43], kc);
			if (kc > 1000) atomicAdd(&counters[44], 1u);
			if (kc > 2000) atomicAdd(&counters[45], 1u);
			if (kc > 500) atomicAdd(&counters[46], 1u);
			if (SV_TIMING >= 2 && (kc > 700 || (q & 127u) == 0))
				printf("SVQ q=%u kc=%u order=%u nLists=%u nGrams=%u total=%u cand=%u probes=%u slow=%u sieve=%u push=%u resolve=%u\n", q, kc, wi, nLists, nGrams,
				       d.total, sh.tim[3], sh.nProbe, sh.tim[4], sh.tim[0] >> 4, sh.tim[1] >> 4, sh.tim[2] >> 4);
		}
#endif
	}
}

}   // namespace
