// hash_path.cuh — k_hash_search: posting lists -> top-k without merging them (included by search.cu after fast_path.cuh).
//
// Observation behind it: once the running k-th best (tau) says that a term matched by ONE query gram can no longer make the
// top-k — because of the threshold (nGramSearch.hpp:313-316) or because k better candidates are already known — the only
// postings that matter are the terms that occur in AT LEAST TWO of the query's posting lists. Finding those does not need
// the lists merged or counted (hpp:289-298 increments a counter per posting): it needs a duplicate detector.
//
// One CTA = 8 consumer warps + 1 producer warp, one query at a time, posting lists consumed in passes aligned on term-id
// ranges (every list contributes a window, as in k_fast_search).
//   producer: stages the windows of pass p+1 with TMA bulk copies into the second stage buffer while the consumers work on
//     pass p; when a buffer has landed it computes the pass bound, how much of every list the pass consumes (binary search),
//     replaces every slot of the touched 16-byte groups that holds no posting of the pass by a DEAD value, lays the groups
//     out in one flat index space and publishes the plan (mbarrier). Nothing of this is on the consumers' critical path.
//   consumers, FILTER pass:
//     level 1: every staged posting (rel = term - passLo, list) is written to slot hash(rel) of a shared-memory table with a
//       PLAIN store — no atomics, colliding stores overwrite each other — and read back after a barrier:
//         slot == own key              -> the posting owns the slot: nothing to do (the common case)
//         slot == same rel, other list -> the term is in two lists: the posting is FLAGGED
//         slot == another term         -> collision: the posting is re-hashed with another multiplier into a second table
//                                         (level 2); what still collides there is counted like a flagged posting
//       A false duplicate is impossible (keys hold the full rel), a missed one too: copies of a term share their slot on
//       every level, so either one copy owns the slot and all others see it, or all lose to another term and move on together.
//     flagged postings (a few per cent) are counted exactly in a small open-addressing table rel -> match count (atomicCAS
//       insert, atomic add; the owner's multiplicity is added by whoever creates the entry). Terms whose count survives the
//       integer pre-filter get the reference's score / weight / promotion code shared with k_fast_search (hpp:299-341).
//   consumers, PROBE pass: the head of the id range, where single-list terms still matter (kid below tau's kid in the count-1
//     class): every posting goes through such a count table, exact for any input.
// Lists are read once (algorithmic bytes unchanged). Queries the kernel cannot take (edit-distance path, repeated query
// grams that keep single-list terms alive, limit > 512, table overflow) are handed to k_fast_search through fastList2.
#pragma once

namespace {

constexpr int HS_CONSUMERS = 256;
constexpr int HS_CWARPS = HS_CONSUMERS / 32;
constexpr int HS_THREADS = HS_CONSUMERS + 32;                    // + the producer warp
constexpr int HS_RMAX = 4;                                       // 16-byte posting groups per consumer thread per pass
constexpr int HS_ITEMS = HS_RMAX * 4;                            // posting slots per thread per pass (kept in registers)
constexpr int HS_GROUPS = HS_CONSUMERS * HS_RMAX;                // 1024 groups = 4096 posting slots per pass
constexpr int HS_WIN_ALLOC = HS_GROUPS * 4 + MAX_QGRAMS * FS_REGION_PAD + 16;
constexpr int HS_T1_BITS = 13, HS_T1 = 1 << HS_T1_BITS;          // level-1 table (also the PROBE count table)
constexpr int HS_T2_BITS = 12, HS_T2 = 1 << HS_T2_BITS;          // level-2 table
constexpr int HS_CT_BITS = 9, HS_CT = 1 << HS_CT_BITS;           // count table of the flagged postings
constexpr int HS_LIST_CAP = 5 * HS_CONSUMERS;                    // level-1 losers
constexpr int HS_REL_BITS = 26;                                  // key = rel << 6 | list
constexpr uint32_t HS_DEADREL = (1u << HS_REL_BITS) - 1;         // rel of a slot that holds no posting of the pass
constexpr int HS_KLIM = 512;                                     // limit <= 512: candidate buffer of 1024 entries
constexpr uint32_t HS_EMPTY = 0xFFFFFFFFu;
constexpr int HS_BAR = 1;                                        // named barrier of the consumer warps

struct Stage {                                                   // one per stage buffer: producer -> consumers
	PassPlan plan;                                               // layout of the staged windows (producer only)
	uint32_t segStart[MAX_QGRAMS + 2];                           // first group (flat index) of list i; [nLists] = gTot, then INF
	uint32_t segAddr[MAX_QGRAMS + 1];                            // shared byte address of flat group 0 as seen from list i
	uint32_t warpSeg[HS_CWARPS];                                 // list that holds the first group of consumer warp w
	uint32_t passLo, passHi, gTot, mode, lastPass;
};

struct HashShared {
	alignas(16) uint32_t win[2][HS_WIN_ALLOC];
	alignas(16) uint32_t tab[HS_T1];
	uint32_t tab2[HS_T2];
	uint32_t ct[HS_CT];                                          // rel << 6 | count; HS_EMPTY when free (kept clean by the scan)
	uint32_t list[HS_LIST_CAP];                                  // keys of the level-1 losers
	Stage st[2];
	unsigned long long* buf;                                     // candidate buffer: bufCap entries behind this struct
	uint32_t bufCap;
	alignas(8) unsigned long long full[2];                       // TMA completion of stage buffer b
	alignas(8) unsigned long long planned[2];                    // producer: the plan of stage buffer b is published
	alignas(8) unsigned long long consumed[2];                   // consumers: stage buffer b has been read (one arrival per warp)
	uint32_t hist[256];
	uint32_t mult[MAX_QGRAMS], listOff[MAX_QGRAMS];
	uint64_t qcode[MAX_QGRAMS];
	uint32_t hiKey[MAX_QGRAMS + 1], exactKey[MAX_QGRAMS + 1];
	float sOf[MAX_QGRAMS + 1];
	unsigned long long tau;
	uint32_t promotedKey;
	uint32_t bufCount, overflow, query, selPrefixLo, selK, listCount, noneWord;
	volatile uint32_t passesDone;                                // passes (counted over the whole kernel) the consumers have finished
	uint32_t qlen, nGrams, cmin, cprom, nLists, total, maxMult;
	uint32_t shortActive, scanLong, scanLongOnly;                // read by the shared scoring helpers; always 0 here
	float threshold;
	int topShift, kidBits;
	uint32_t classMode;
	IndexView ix;
};

__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32m(uint32_t addr)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr)
{
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ uint32_t atom_cas_shared(uint32_t addr, uint32_t cmp, uint32_t val)
{
	uint32_t old;
	asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(addr), "r"(cmp), "r"(val) : "memory");
	return old;
}
__device__ __forceinline__ void red_add_shared(uint32_t addr, uint32_t val)
{
	asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(val) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync() { group_sync<HS_BAR, HS_CONSUMERS>(); }

__device__ __forceinline__ uint32_t hs_slot1(uint32_t tabAddr, uint32_t rel) { return tabAddr + (((rel * 0x9E3779B1u) >> (32 - HS_T1_BITS)) << 2); }
__device__ __forceinline__ uint32_t hs_slot2(uint32_t tab2Addr, uint32_t rel) { return tab2Addr + (((rel * 0x85EBCA6Bu) >> (32 - HS_T2_BITS)) << 2); }

// count table (open addressing, linear probing): adds `m` matches to term `rel`; whoever creates the entry also adds `extra`
// (the multiplicity of the copy that owns the hash slot and therefore never gets here).
template <int BITS>
__device__ __forceinline__ void ct_add(HashShared& sh, uint32_t ctAddr, uint32_t rel, uint32_t m, uint32_t extra)
{
	uint32_t h = (rel * 0xC2B2AE35u) >> (32 - BITS);
	for (int step = 0; step < (1 << BITS); ++step) {
		const uint32_t a = ctAddr + (h << 2);
		const uint32_t old = atom_cas_shared(a, HS_EMPTY, (rel << 6) | (m + extra));
		if (old == HS_EMPTY) return;
		if ((old >> 6) == rel) { red_add_shared(a, m); return; }
		h = (h + 1) & ((1u << BITS) - 1);
	}
	sh.overflow = 3;               // table full: k_fast_search takes the query
}

// a term with exact match count c: the common case (uniform weights, kid == term id, no exact-match promotion possible) is
// decided right here, everything else goes through the reference's scoring code (hpp:310-341)
template <class U>
__device__ __forceinline__ void emit_term(HashShared& sh, const U* __restrict__ qn, uint32_t term, uint32_t c)
{
	if (sh.classMode && c < sh.cprom) {
		const unsigned long long key = (((unsigned long long)sh.exactKey[c]) << 32) | term;
		if (key < sh.tau) append_key(sh, key);
	} else emit_candidate<U, false>(sh, qn, term, c);
}

// the 16-byte groups of the stage buffer a consumer thread owns: warp w owns the flat groups [w * 128, w * 128 + 128), lane l
// of it the groups l, l + 32, l + 64, l + 96 of those; `fn(x, list, slot)` is called for each posting slot of a staged group,
// `none(slot)` for the slots of groups past the end of the pass
template <class FN, class NONE>
__device__ __forceinline__ void for_each_slot(const Stage& st, FN fn, NONE none)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t gTot = st.gTot;
	uint32_t seg = st.warpSeg[warp], nb = st.segStart[seg + 1], sa = st.segAddr[seg];
#pragma unroll
	for (int r = 0; r < HS_RMAX; ++r) {
		const uint32_t f = warp * (HS_RMAX * 32) + r * 32 + lane;
		if (f < gTot) {
			while (f >= nb) { ++seg; nb = st.segStart[seg + 1]; sa = st.segAddr[seg]; }
			const uint4 v = lds_v4(sa + f * 16);
			fn(v.x, seg, r * 4 + 0); fn(v.y, seg, r * 4 + 1); fn(v.z, seg, r * 4 + 2); fn(v.w, seg, r * 4 + 3);
		} else { none(r * 4 + 0); none(r * 4 + 1); none(r * 4 + 2); none(r * 4 + 3); }
	}
}

// FILTER pass (consumers): see the file comment.
template <class U>
__device__ __forceinline__ void hash_pass(HashShared& sh, const Stage& st, unsigned long long* consumedBar, const U* __restrict__ qn, uint32_t cRej)
{
	const int tid = threadIdx.x, lane = tid & 31;
	const uint32_t passLo = st.passLo;
	const uint32_t tabAddr = smem_u32(&sh.tab[0]), tab2Addr = smem_u32(&sh.tab2[0]), ctAddr = smem_u32(&sh.ct[0]);
	uint32_t key[HS_ITEMS], adr[HS_ITEMS];
	// ---- level 1, write (slots without a posting carry rel == HS_DEADREL: they hash like a term and are dropped on level 2)
	const uint32_t noneAddr = smem_u32(&sh.noneWord);            // holds HS_EMPTY: "owns its slot" for slots without a group
	for_each_slot(st, [&](uint32_t x, uint32_t seg, int i) {
		const uint32_t rel = x - passLo;
		adr[i] = hs_slot1(tabAddr, rel);
		key[i] = rel * 64u + seg;
		sts_u32(adr[i], key[i]);
	}, [&](int i) { adr[i] = noneAddr; key[i] = HS_EMPTY; });
	__syncwarp();
	if (lane == 0) mbar_arrive(consumedBar);                     // this warp no longer needs the stage buffer
	consumer_sync();
	// ---- level 1, read: postings that do not own their slot are queued
	{
		uint32_t mis = 0;
#pragma unroll
		for (int i = 0; i < HS_ITEMS; ++i) {
			const uint32_t seen = lds_u32m(adr[i]);
			if (seen != key[i]) mis |= 1u << i;
		}
		const uint32_t cnt = __popc(mis);
		uint32_t inc = cnt;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
		uint32_t base = 0;
		if (lane == 31 && inc) base = atomicAdd(&sh.listCount, inc);
		base = __shfl_sync(0xffffffffu, base, 31) + inc - cnt;
		if (base + cnt > (uint32_t)HS_LIST_CAP) { if (cnt) sh.overflow = 2; mis = 0; }
#pragma unroll
		for (int i = 0; i < HS_ITEMS; ++i)
			if (mis & (1u << i)) sh.list[base++] = key[i];
	}
	consumer_sync();
	if (sh.overflow) return;
	// ---- level 2: flagged postings go to the count table, colliders are re-hashed into the second table
	const uint32_t n = sh.listCount;
	for (uint32_t i = tid; i < n; i += HS_CONSUMERS) {
		const uint32_t k = sh.list[i], rel = k >> 6;
		if (rel == HS_DEADREL) { sh.list[i] = HS_EMPTY; continue; }
		const uint32_t v = lds_u32m(hs_slot1(tabAddr, rel));
		if ((k ^ v) < 64u) {
			ct_add<HS_CT_BITS>(sh, ctAddr, rel, sh.mult[k & 63u], sh.mult[v & 63u]);
			sh.list[i] = HS_EMPTY;
		} else sts_u32(hs_slot2(tab2Addr, rel), k);
	}
	consumer_sync();
	for (uint32_t i = tid; i < n; i += HS_CONSUMERS) {
		const uint32_t k = sh.list[i];
		if (k != HS_EMPTY) {
			const uint32_t rel = k >> 6;
			const uint32_t v = lds_u32m(hs_slot2(tab2Addr, rel));
			// the owner of the level-2 slot is done; copies that see it are flagged; what still collides is counted as it is
			// (all copies of such a term are in the same situation, so its count is complete either way)
			if (v != k) ct_add<HS_CT_BITS>(sh, ctAddr, rel, sh.mult[k & 63u], ((k ^ v) < 64u) ? sh.mult[v & 63u] : 0u);
		}
	}
	consumer_sync();
	// ---- exact counts of the flagged terms: the ones that survive the integer pre-filter are scored (hpp:299-341)
#pragma unroll
	for (int r = 0; r < HS_CT / HS_CONSUMERS; ++r) {
		const uint32_t v = sh.ct[tid + r * HS_CONSUMERS];
		if (v != HS_EMPTY) {
			sh.ct[tid + r * HS_CONSUMERS] = HS_EMPTY;
			const uint32_t c = v & 63u;
			if (c > cRej) emit_term<U>(sh, qn, passLo + (v >> 6), c);
		}
		__syncwarp();
	}
}

// PROBE pass (consumers): every posting is a candidate (single-list terms can still make the top-k in this id range), so
// the exact match count of every term is needed: all postings go through a count table (hpp:289-298 per pass).
template <class U>
__device__ __forceinline__ void probe_pass(HashShared& sh, const Stage& st, unsigned long long* consumedBar, const U* __restrict__ qn, uint32_t cRej)
{
	const int tid = threadIdx.x, lane = tid & 31;
	const uint32_t passLo = st.passLo;
	const uint32_t tabAddr = smem_u32(&sh.tab[0]);
	for (int i = tid; i < HS_T1; i += HS_CONSUMERS) sh.tab[i] = HS_EMPTY;
	consumer_sync();
	for_each_slot(st, [&](uint32_t x, uint32_t seg, int) {
		const uint32_t rel = x - passLo;
		if (rel != HS_DEADREL) ct_add<HS_T1_BITS>(sh, tabAddr, rel, sh.mult[seg], 0u);
	}, [](int) {});
	__syncwarp();
	if (lane == 0) mbar_arrive(consumedBar);
	consumer_sync();
	for (int i = tid; i < HS_T1; i += HS_CONSUMERS) {
		const uint32_t v = sh.tab[i];
		if (v != HS_EMPTY) {
			const uint32_t c = v & 63u;
			if (c > cRej) emit_term<U>(sh, qn, passLo + (v >> 6), c);
		}
		__syncwarp();
	}
}

// producer: lays out the windows of a pass in stage buffer b from the cursors (cur, rem: lane = list) and issues the TMA copies
__device__ __forceinline__ void stage_issue(HashShared& sh, int b, uint32_t cur, uint32_t rem, uint32_t cap, uint32_t nLists,
                                            const uint32_t* __restrict__ postings, int lane)
{
	const uint32_t wmax = max(1u, cap / nLists);
	const uint32_t region = (wmax + 10) & ~3u;                     // <= 3 shift + (wmax + 1) data + round-up to 16 bytes
	PassPlan& pl = sh.st[b].plan;
	uint32_t bytes = 0;
	const uint32_t* src = nullptr;
	uint32_t* dst = nullptr;
	if ((uint32_t)lane < nLists) {
		const uint32_t e0 = sh.listOff[lane] + cur;                   // absolute posting index of the cursor
		const uint32_t shift = e0 & 3u;                                // 16-byte alignment of the global source
		const uint32_t take = min(rem, wmax + 1);
		pl.wbase[lane] = lane * region + shift;
		pl.avail[lane] = min(rem, wmax);
		pl.more[lane] = rem > wmax;
		if (take) {
			bytes = ((shift + take + 3) & ~3u) * 4;
			src = postings + (e0 - shift);
			dst = &sh.win[b][lane * region];
		}
	}
	uint32_t total = bytes;
	for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
	if (lane == 0) {
		pl.wmax = wmax;
		// the buffer was read / written through the generic proxy by an earlier pass: order that before the async writes
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		mbar_expect_tx(&sh.full[b], total);                           // total == 0: the phase completes right here
	}
	__syncwarp();
	if (bytes) tma_load_1d(dst, src, bytes, &sh.full[b]);
}

template <class U>
__global__ void __launch_bounds__(HS_THREADS, 2) k_hash_search(IndexView ix, const QDesc* __restrict__ desc, const U* __restrict__ qnorm,
                                                            const uint32_t* __restrict__ fastList, uint32_t* __restrict__ counters,
                                                            uint32_t* __restrict__ fastList2, uint32_t limit, uint32_t outStride,
                                                            uint32_t* __restrict__ outCount, uint32_t* __restrict__ outKid,
                                                            float* __restrict__ outScore, unsigned long long* __restrict__ statCounters,
                                                            uint32_t bufCap, float threshold)
{
	extern __shared__ __align__(16) unsigned char hs_smem[];
	HashShared& sh = *reinterpret_cast<HashShared*>(hs_smem);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const bool producer = warp == HS_CWARPS;
	const uint32_t nFast = counters[0];
	const uint32_t k = limit;
	for (int i = tid; i < HS_CT; i += HS_THREADS) sh.ct[i] = HS_EMPTY;
	if (tid == 0) {
		sh.ix = ix;
		sh.buf = reinterpret_cast<unsigned long long*>(hs_smem + ((sizeof(HashShared) + 15) & ~size_t(15)));
		sh.bufCap = bufCap;
		sh.passesDone = 0;
		sh.noneWord = HS_EMPTY;
		for (int b = 0; b < 2; ++b) {
			mbar_init(&sh.full[b], 1);
			mbar_init(&sh.planned[b], 1);
			mbar_init(&sh.consumed[b], HS_CWARPS);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	uint32_t gp = 0;                        // passes started since the kernel began (same value in every thread): pass gp uses
	                                        // stage buffer gp & 1, phase parity (gp >> 1) & 1 of that buffer's barriers

	while (true) {
		__syncthreads();
		if (tid == 0) sh.query = atomicAdd(&counters[2], 1u);
		__syncthreads();
		const uint32_t wi = sh.query;
		if (wi >= nFast) break;
		const uint32_t q = fastList[wi];
		const QDesc& d = desc[q];
		const U* qn = qnorm + d.qbeg;
		const uint32_t nLists = d.nLists, nGrams = d.nGrams;
		const bool classMode = ix.identity && ix.uniformW;
		uint32_t maxMult = 0;
		for (uint32_t i = 0; i < nLists; ++i) maxMult = max(maxMult, (uint32_t)d.listMult[i]);
		// FILTER passes need "a term found in one list only is out": by the threshold (caseA) or, with uniform weights and
		// kid == term id, behind the kid of a seeded tau (caseB; the id range in front of it is covered by PROBE passes)
		const bool caseA = (d.cmin ? d.cmin - 1 : 0u) >= maxMult;
		const bool caseB = classMode && maxMult == 1 && d.cprom > 1 && d.cmin <= 1;
		bool eligible = nLists >= 1 && !(d.flags & 1u) && k <= (uint32_t)HS_KLIM && (caseA || caseB);
		if (eligible) {
			if (tid < MAX_QGRAMS) {
				sh.qcode[tid] = (uint32_t)tid < nGrams ? d.qcode[tid] : ~0ull;
				sh.listOff[tid] = (uint32_t)tid < nLists ? d.listOff[tid] : 0u;
				sh.mult[tid] = (uint32_t)tid < nLists ? d.listMult[tid] : 0u;
			}
			if (tid <= MAX_QGRAMS) {
				const uint32_t c = tid;
				const float s = (c <= nGrams && nGrams) ? term_score(c, nGrams) : 0.0f;
				sh.sOf[c] = s;
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
				sh.bufCount = 0; sh.overflow = 0; sh.tau = ~0ull; sh.listCount = 0;
				sh.qlen = d.qlen; sh.nGrams = nGrams; sh.cmin = d.cmin; sh.cprom = d.cprom; sh.nLists = nLists; sh.total = d.total;
				sh.maxMult = maxMult;
				sh.classMode = classMode;
				sh.shortActive = 0; sh.scanLong = 0; sh.scanLongOnly = 0; sh.threshold = threshold;
				sh.promotedKey = classMode ? 0u : ~float_to_ordered(100.0f);
				sh.topShift = classMode ? 32 : 56;
				int kb = 1;
				while (kb < 32 && (1ull << kb) < ix.K) ++kb;
				sh.kidBits = kb;
			}
			__syncthreads();
			if (!caseA) {
				// seed tau: the first k postings of one list are k distinct keys with at least one match each (see k_fast_search)
				if (warp == 0) {
					uint32_t x = INF;
					if ((uint32_t)lane < nLists && d.listLen[lane] >= k) x = __ldg(ix.postings + d.listOff[lane] + (k - 1));
					for (int o = 16; o; o >>= 1) x = min(x, __shfl_xor_sync(0xffffffffu, x, o));
					if (lane == 0 && x != INF) sh.tau = ((((unsigned long long)sh.exactKey[1]) << 32) | x) + 1;
				}
				__syncthreads();
				if (sh.tau == ~0ull) eligible = false;           // short lists: k_fast_search takes the query
			}
		}
		if (!eligible) {
			if (tid == 0) fastList2[atomicAdd(&counters[5], 1u)] = q;
			continue;
		}

		if (producer) {
			// =================================================================================================== producer
			const bool on = (uint32_t)lane < nLists;
			uint32_t cur = 0, rem = on ? d.listLen[lane] : 0u;      // lane = list
			uint32_t passLo = 0;
			const uint32_t capFilter = 4u * (uint32_t)(HS_GROUPS - 2 * (int)nLists);
			// a PROBE posting appends at most one candidate (kid == term id) and the buffer is at most half full when a
			// pass starts; the first one only has to find k terms to replace the seeded tau
			const uint32_t capProbe = min(capFilter, min(bufCap / 2, max(3u * k, 128u)));
			// single-list terms still matter below tau's kid while tau sits in the count-1 class: PROBE pass
			auto wants_probe = [&](uint32_t nextLo) {
				const unsigned long long tau = sh.tau;
				return !caseA && (uint32_t)(tau >> 32) == sh.exactKey[1] && nextLo < (uint32_t)tau;
			};
			bool probe = wants_probe(0);
			stage_issue(sh, gp & 1, cur, rem, probe ? capProbe : capFilter, nLists, ix.postings, lane);
			while (true) {
				const int b = gp & 1;
				const uint32_t ph = (gp >> 1) & 1;
				Stage& st = sh.st[b];
				const PassPlan& pl = st.plan;
				uint32_t* wb = sh.win[b];
				mbar_wait(&sh.full[b], ph);
				// ---- pass bound: the smallest posting that is NOT staged; consumed counts
				const uint32_t wmax = pl.wmax;
				uint32_t wbase = 0, avail = 0, more = 0, bound = INF, last = 0;
				if (on) {
					wbase = pl.wbase[lane]; avail = pl.avail[lane]; more = pl.more[lane];
					if (more) bound = wb[wbase + wmax];
					else if (avail) last = wb[wbase + avail - 1];
				}
				for (int o = 16; o; o >>= 1) {
					bound = min(bound, __shfl_xor_sync(0xffffffffu, bound, o));
					last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
				}
				// keys hold rel in HS_REL_BITS bits (HS_DEADREL excluded); a PROBE pass ends where single-list terms stop
				// mattering (tau's kid)
				unsigned long long clip = (unsigned long long)passLo + HS_DEADREL;
				if (probe) {
					const unsigned long long tau = sh.tau;
					if ((uint32_t)(tau >> 32) == sh.exactKey[1] && (uint32_t)tau > passLo) clip = min(clip, (unsigned long long)(uint32_t)tau);
				}
				uint32_t passHi = bound == INF ? last + 1 : bound;
				if (passHi < passLo) passHi = passLo;                 // nothing staged at all
				bool clipped = false;
				if (clip < (unsigned long long)passHi) { passHi = (uint32_t)clip; clipped = true; }
				uint32_t nt = 0;
				if (on) nt = (bound == INF && !clipped) ? avail : lower_bound_smem(wb + wbase, avail, passHi);
				const uint32_t curN = cur + nt, remN = rem - nt;
				uint32_t best = nt;
				for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
				bool lastPass = !__any_sync(0xffffffffu, remN > 0);
				if (best == 0 && !clipped && !lastPass) {            // no list can advance (corrupt postings): never spin
					if (lane == 0) sh.overflow = 5;
					lastPass = true;
				}
				// ---- FILTER pass: the next pass is a FILTER pass too (tau only improves) and can be staged right away
				if (!probe && !lastPass) {
					if (gp >= 1) mbar_wait(&sh.consumed[b ^ 1], ((gp - 1) >> 1) & 1);
					stage_issue(sh, b ^ 1, curN, remN, capFilter, nLists, ix.postings, lane);
				}
				// ---- every slot of the touched groups that holds no posting of this pass becomes DEAD: the alignment slots in
				// front of the window and whatever follows the consumed part
				const uint32_t deadX = passLo + HS_DEADREL;
				const uint32_t shift = wbase & 3u;
				if (on && nt) {
					for (uint32_t j = wbase - shift; j < wbase; ++j) wb[j] = deadX;
					wb[wbase + nt] = deadX; wb[wbase + nt + 1] = deadX; wb[wbase + nt + 2] = deadX;
				}
				// ---- flat group index space over the windows: list i owns groups [segStart[i], segStart[i+1])
				const uint32_t ng = (on && nt) ? (shift + nt + 3) >> 2 : 0u;
				uint32_t inc = ng;
				for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
				const uint32_t gTot = __shfl_sync(0xffffffffu, inc, 31);
				if (on) {
					st.segStart[lane] = inc - ng;
					st.segAddr[lane] = smem_u32(&wb[wbase - shift]) - (inc - ng) * 16;
				}
				{
					// the list that holds flat group w * 128 = the number of lists that end at or before it (lanes 0..7: w = lane)
					const uint32_t f0 = (uint32_t)lane * (HS_RMAX * 32);
					uint32_t sgi = 0;
					for (uint32_t i = 0; i < nLists; ++i) sgi += __shfl_sync(0xffffffffu, inc, i) <= f0;
					if (lane < HS_CWARPS) st.warpSeg[lane] = min(sgi, nLists);
				}
				if (lane == 0) {
					st.segStart[nLists] = gTot; st.segStart[nLists + 1] = INF; st.segAddr[nLists] = 0;
					st.gTot = gTot; st.passLo = passLo; st.passHi = passHi; st.mode = probe; st.lastPass = lastPass;
				}
				__syncwarp();
				if (lane == 0) mbar_arrive(&sh.planned[b]);          // release: the plan and the DEAD slots are visible
				// ---- after a PROBE pass the kind of the next pass depends on the tau its candidates establish: wait for them
				if (probe && !lastPass) {
					for (uint32_t spin = 0; sh.passesDone <= gp; ++spin)
						if (spin > (1u << 26)) __trap();                // a lost hand-over must never hang the GPU
					__threadfence_block();
					__syncwarp();
					probe = wants_probe(passHi);
					if (gp >= 1) mbar_wait(&sh.consumed[b ^ 1], ((gp - 1) >> 1) & 1);
					stage_issue(sh, b ^ 1, curN, remN, probe ? capProbe : capFilter, nLists, ix.postings, lane);
				}
				cur = curN; rem = remN; passLo = passHi;
				++gp;
				if (lastPass) break;
			}
			continue;                                              // joins the consumers at the barrier on top of the loop
		}

		// ======================================================================================================= consumers
		bool failed = false;
		while (true) {
			const int b = gp & 1;
			const uint32_t ph = (gp >> 1) & 1;
			const Stage& st = sh.st[b];
			mbar_wait(&sh.planned[b], ph);
			const bool lastPass = st.lastPass != 0;
			if (sh.overflow) failed = true;
			if (!failed) {
				// integer pre-filter: a count c <= cRej can neither pass the threshold nor beat the running k-th best; with
				// kid == term id >= passLo the bound is exact for the tie class of tau
				uint32_t cRej = sh.cmin ? sh.cmin - 1 : 0;
				{
					const unsigned long long tau = sh.tau;
					const unsigned long long lowKid = classMode ? st.passLo : 0u;
					const uint32_t c = cRej + 1 + lane;
					const bool rejc = c <= nGrams && ((((unsigned long long)sh.hiKey[min(c, (uint32_t)MAX_QGRAMS)]) << 32) | lowKid) >= tau;
					const unsigned m = __ballot_sync(0xffffffffu, rejc);
					cRej += m == 0xffffffffu ? 32u : (uint32_t)__ffs(~m) - 1;   // leading run of rejectable counts
				}
				if (st.mode) probe_pass<U>(sh, st, &sh.consumed[b], qn, cRej);
				else if (cRej < maxMult) {                            // cannot happen; a FILTER pass would miss terms
					if (lane == 0) { sh.overflow = 4; mbar_arrive(&sh.consumed[b]); }
				} else hash_pass<U>(sh, st, &sh.consumed[b], qn, cRej);
			} else if (lane == 0) mbar_arrive(&sh.consumed[b]);        // given up: keep the pipeline's barriers in step
			consumer_sync();
			if (sh.overflow) failed = true;
			uint32_t level = sh.bufCount;
			if (!failed && level > k && (sh.tau == ~0ull || level > bufCap / 2)) select_smallest<HS_CONSUMERS, HS_BAR>(sh, level, k);
			if (tid == 0) {
				sh.listCount = 0;
				__threadfence_block();
				sh.passesDone = gp + 1;
			}
			++gp;
			if (lastPass) break;
		}
		consumer_sync();
		if (failed) {
			for (int i = tid; i < HS_CT; i += HS_CONSUMERS) sh.ct[i] = HS_EMPTY;
			if (tid == 0) {
				fastList2[atomicAdd(&counters[5], 1u)] = q;
				atomicAdd(&counters[7], 1u);
				atomicAdd(&counters[13 + min(sh.overflow, 3u) - 1], 1u);   // why: [13] candidate buffer [14] loser queue [15] other
			}
			continue;
		}
		// ---- final: top-k, sorted
		uint32_t level = sh.bufCount;
		if (level > k) { select_smallest<HS_CONSUMERS, HS_BAR>(sh, level, k); level = k; }
		sort_buffer<HS_CONSUMERS, HS_BAR>(sh, level);
		for (uint32_t i = tid; i < level; i += HS_CONSUMERS) {
			const unsigned long long v = sh.buf[i];
			outKid[(size_t)q * outStride + i] = (uint32_t)v;
			const uint32_t hi = (uint32_t)(v >> 32);
			outScore[(size_t)q * outStride + i] = classMode ? (hi == 0 ? 100.0f : edge_value(1.0f, sh.sOf[nGrams - hi + 1])) : rank_key_score(v);
		}
		if (tid == 0) {
			outCount[q] = level;
			atomicAdd(&statCounters[0], (unsigned long long)d.total);
			atomicAdd(&counters[6], 1u);
		}
	}
}

}   // namespace
