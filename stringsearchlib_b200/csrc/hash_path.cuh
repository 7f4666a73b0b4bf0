// hash_path.cuh — k_hash_search: posting lists -> top-k without merging them (included by search.cu after fast_path.cuh).
//
// Observation behind it: once the running k-th best (tau) says that a term matched by ONE query gram can no longer make the
// top-k — because of the threshold (nGramSearch.hpp:313-316) or because k better candidates are already known — the only
// postings that matter are the terms that occur in AT LEAST TWO of the query's posting lists. Finding those does not need
// the lists merged or counted (hpp:289-298 increments a counter per posting): it needs a duplicate detector. Per pass
// (a term-id range, staged by TMA as in k_fast_search, the next pass's copies already in flight):
//   * level 1: every staged posting (rel = term - passLo, list) is written to slot hash(rel) of a shared-memory table with a
//     PLAIN store — no atomics, colliding stores overwrite each other — and read back after a barrier:
//       slot == own key              -> the posting owns the slot: nothing to do (the common case)
//       slot == same rel, other list -> the term is in two lists: the posting is FLAGGED
//       slot == another term         -> collision: the posting is re-hashed with another multiplier into a second table
//                                       (level 2); what still collides there is treated like a flagged posting
//     A false duplicate is impossible (keys hold the full rel), a missed one too: copies of a term share their slot on every
//     level, so either one copy owns the slot and all others see it, or all lose to another term and move on together.
//   * flagged postings (a few per cent) are counted exactly in a small open-addressing table rel -> match count (atomicCAS
//     insert, atomicAdd; the owner's multiplicity is added by whoever creates the entry). Terms whose count survives the
//     integer pre-filter get the reference's score / weight / promotion code shared with k_fast_search (hpp:299-341).
//   * the head of the id range, where single-list terms still matter (kid below tau's kid in the count-1 class), is handled by
//     PROBE passes: every posting goes through such a count table, exact for any input.
// Lists are read once (algorithmic bytes unchanged). Queries the kernel cannot take (edit-distance path, repeated query
// grams that keep single-list terms alive, limit > 512, table overflow) are handed to k_fast_search through fastList2.
#pragma once

namespace {

constexpr int HS_THREADS = 256;
constexpr int HS_RMAX = 4;                                       // 16-byte posting groups per thread per pass
constexpr int HS_GROUPS = HS_THREADS * HS_RMAX;                  // 1024 groups = 4096 posting slots per pass
constexpr int HS_ITEMS = HS_RMAX * 4;                            // postings per thread per pass (kept in registers)
constexpr int HS_WIN_ALLOC = HS_GROUPS * 4 + MAX_QGRAMS * FS_REGION_PAD + 16;
constexpr int HS_TAB_BITS = 14, HS_TAB = 1 << HS_TAB_BITS;       // level-1 table
constexpr int HS_TAB2_BITS = 11, HS_TAB2 = 1 << HS_TAB2_BITS;    // level-2 table
constexpr int HS_PROBE_BITS = 12, HS_PROBE = 1 << HS_PROBE_BITS; // PROBE passes: count table inside the level-1 table
constexpr int HS_CT_BITS = 9, HS_CT = 1 << HS_CT_BITS;           // FILTER passes: count table of the flagged postings
constexpr int HS_LIST_CAP = 3 * HS_THREADS;                      // level-1 losers (3 per thread on level 2)
constexpr int HS_REL_BITS = 26;                                  // key = rel << 6 | list; a pass spans < 2^26 term ids
constexpr int HS_KLIM = 512;                                     // limit <= 512: candidate buffer of 1024 entries
constexpr uint32_t HS_EMPTY = 0xFFFFFFFFu;

static_assert(HS_LIST_CAP == 3 * HS_THREADS, "queue entries are processed three per thread");

struct HashShared {
	alignas(16) uint32_t win[HS_WIN_ALLOC];
	alignas(16) uint32_t tab[HS_TAB + 4];                        // [HS_TAB] = slot of masked (out-of-range) postings
	uint32_t tab2[HS_TAB2];
	uint32_t ct[HS_CT];                                          // rel << 6 | count; HS_EMPTY when free (kept clean by the scan)
	unsigned long long list[HS_LIST_CAP];                        // level-1 losers: (slot value seen << 32) | own key
	unsigned long long* buf;                                     // candidate buffer: bufCap entries behind this struct
	uint32_t bufCap;
	alignas(8) unsigned long long bar;
	uint32_t hist[256];
	PassPlan plan;
	uint32_t cur[MAX_QGRAMS], rem[MAX_QGRAMS], nTake[MAX_QGRAMS], mult[MAX_QGRAMS], listOff[MAX_QGRAMS];
	uint32_t winByte[MAX_QGRAMS];
	uint32_t segStart[MAX_QGRAMS + 2];                           // first group (flat index) of list i; [nLists] = gTot, then INF
	uint32_t segAddr[MAX_QGRAMS + 1];                            // shared byte address of flat group 0 as seen from list i
	uint64_t qcode[MAX_QGRAMS];
	uint32_t hiKey[MAX_QGRAMS + 1], exactKey[MAX_QGRAMS + 1];
	float sOf[MAX_QGRAMS + 1];
	unsigned long long tau;
	uint32_t promotedKey;
	uint32_t bufCount, overflow, query, selPrefixLo, selK;
	uint32_t passLo, passHi, cRej, lastPass, prefetched, issuedNext, mode, nextMode, gTot;
	uint32_t listCount;
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

// count table (open addressing, linear probing): adds `m` matches to term `rel`; whoever creates the entry also adds `extra`
// (the multiplicity of the copy that owns the hash slot and therefore never gets here). Returns the slot address when this
// call created the entry, 0 otherwise.
template <int BITS>
__device__ __forceinline__ uint32_t ct_add(HashShared& sh, uint32_t ctAddr, uint32_t rel, uint32_t m, uint32_t extra)
{
	uint32_t h = (rel * 0xC2B2AE35u) >> (32 - BITS);
	for (int step = 0; step < 64; ++step) {
		const uint32_t a = ctAddr + (h << 2);
		const uint32_t old = atom_cas_shared(a, HS_EMPTY, (rel << 6) | (m + extra));
		if (old == HS_EMPTY) return a;
		if ((old >> 6) == rel) { red_add_shared(a, m); return 0; }
		h = (h + 1) & ((1u << BITS) - 1);
	}
	sh.overflow = 1;               // table (nearly) full: k_fast_search takes the query
	return 0;
}

// FILTER pass over the staged windows (see the file comment). CTA-wide. `issueNext` (warp 0) stages the next pass as soon as
// the windows have been consumed into registers.
template <class U, class ISSUE>
__device__ __forceinline__ void hash_pass(HashShared& sh, const U* __restrict__ qn, ISSUE issueNext)
{
	const int tid = threadIdx.x, lane = tid & 31;
	const uint32_t passLo = sh.passLo, span = sh.passHi - passLo, gTot = sh.gTot;
	const uint32_t tabAddr = smem_u32(&sh.tab[0]), tab2Addr = smem_u32(&sh.tab2[0]), ctAddr = smem_u32(&sh.ct[0]);
	const uint32_t dummy = tabAddr + HS_TAB * 4, infAddr = smem_u32(&sh.win[HS_WIN_ALLOC - 4]);
	uint32_t key[HS_ITEMS], adr[HS_ITEMS];
	// ---- level 1, write: thread t owns the 16-byte groups t, t + 256, ... of the concatenated windows
	{
		uint32_t seg = 0, nb = sh.segStart[1], sa = sh.segAddr[0];
#pragma unroll
		for (int r = 0; r < HS_RMAX; ++r) {
			const uint32_t f = tid + r * HS_THREADS;
			while (f >= nb) { ++seg; nb = sh.segStart[seg + 1]; sa = sh.segAddr[seg]; }
			const uint4 v = lds_v4(f < gTot ? sa + f * 16 : infAddr);
			const uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const uint32_t rel = x[j] - passLo;
				const bool ok = rel < span;                       // leading alignment slots, the INF tail and the next pass's postings drop out
				const uint32_t a = tabAddr + (((rel * 0x9E3779B1u) >> (32 - HS_TAB_BITS)) << 2);
				adr[r * 4 + j] = ok ? a : dummy;
				key[r * 4 + j] = ok ? ((rel << 6) | seg) : 0xFFFFFFFFu;
				sts_u32(adr[r * 4 + j], key[r * 4 + j]);
			}
		}
	}
	__syncthreads();
	issueNext();                                                 // the windows now live in registers: stage the next pass
	// ---- level 1, read: postings that do not own their slot are queued together with what they saw
	{
		uint32_t seen[HS_ITEMS], cnt = 0;
#pragma unroll
		for (int i = 0; i < HS_ITEMS; ++i) {
			seen[i] = lds_u32m(adr[i]);
			cnt += seen[i] != key[i];
		}
		uint32_t inc = cnt;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
		uint32_t base = 0;
		if (lane == 31 && inc) base = atomicAdd(&sh.listCount, inc);
		base = __shfl_sync(0xffffffffu, base, 31) + inc - cnt;
		if (cnt) {
			if (base + cnt > HS_LIST_CAP) sh.overflow = 1;
			else {
#pragma unroll
				for (int i = 0; i < HS_ITEMS; ++i)
					if (seen[i] != key[i]) sh.list[base++] = ((unsigned long long)seen[i] << 32) | key[i];
			}
		}
	}
	__syncthreads();
	if (sh.overflow) return;
	// ---- level 2: flagged postings go to the count table, colliders are re-hashed into the second table
	const uint32_t n = sh.listCount;
	uint32_t k2[3], a2[3];
#pragma unroll
	for (int r = 0; r < 3; ++r) {
		const uint32_t i = tid + r * HS_THREADS;
		k2[r] = HS_EMPTY; a2[r] = 0;
		if (i < n) {
			const unsigned long long e = sh.list[i];
			const uint32_t k = (uint32_t)e, v = (uint32_t)(e >> 32), rel = k >> 6;
			if ((k ^ v) < 64u) ct_add<HS_CT_BITS>(sh, ctAddr, rel, sh.mult[k & 63u], sh.mult[v & 63u]);
			else {
				k2[r] = k;
				a2[r] = tab2Addr + (((rel * 0x85EBCA6Bu) >> (32 - HS_TAB2_BITS)) << 2);
				sts_u32(a2[r], k);
			}
		}
	}
	__syncthreads();
#pragma unroll
	for (int r = 0; r < 3; ++r) {
		if (k2[r] != HS_EMPTY) {
			const uint32_t v = lds_u32m(a2[r]);
			// the owner of the level-2 slot is done; copies that see it are flagged; what still collides is counted as it is
			// (all copies of such a term are in the same situation, so its count is complete either way)
			if (v != k2[r]) ct_add<HS_CT_BITS>(sh, ctAddr, k2[r] >> 6, sh.mult[k2[r] & 63u], ((k2[r] ^ v) < 64u) ? sh.mult[v & 63u] : 0u);
		}
	}
	__syncthreads();
	// ---- exact counts of the flagged terms: the ones that survive the integer pre-filter are scored (hpp:299-341)
	{
		const uint32_t cRej = sh.cRej;
#pragma unroll
		for (int r = 0; r < HS_CT / HS_THREADS; ++r) {
			const uint32_t v = sh.ct[tid + r * HS_THREADS];
			if (v != HS_EMPTY) {
				sh.ct[tid + r * HS_THREADS] = HS_EMPTY;
				const uint32_t c = v & 63u;
				if (c > cRej) emit_candidate<U, false>(sh, qn, passLo + (v >> 6), c);
			}
		}
	}
	__syncthreads();
}

// PROBE pass: every staged posting is a candidate (single-list terms can still make the top-k in this id range), so the
// exact match count of every term is needed: all postings go through a count table (hpp:289-298 per pass); the posting that
// created a term's entry evaluates it. At most ~1000 postings per pass.
template <class U>
__device__ __forceinline__ void probe_pass(HashShared& sh, const U* __restrict__ qn)
{
	const int tid = threadIdx.x;
	const uint32_t passLo = sh.passLo, span = sh.passHi - passLo, gTot = sh.gTot;
	const uint32_t tabAddr = smem_u32(&sh.tab[0]), infAddr = smem_u32(&sh.win[HS_WIN_ALLOC - 4]);
	for (int i = tid; i < HS_PROBE; i += HS_THREADS) sh.tab[i] = HS_EMPTY;
	__syncthreads();
	uint32_t made[HS_ITEMS];                     // slot address of the entries this thread created (0 = none)
	{
		uint32_t seg = 0, nb = sh.segStart[1], sa = sh.segAddr[0];
#pragma unroll
		for (int r = 0; r < HS_RMAX; ++r) {
			const uint32_t f = tid + r * HS_THREADS;
			while (f >= nb) { ++seg; nb = sh.segStart[seg + 1]; sa = sh.segAddr[seg]; }
			const uint4 v = lds_v4(f < gTot ? sa + f * 16 : infAddr);
			const uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const uint32_t rel = x[j] - passLo;
				made[r * 4 + j] = rel < span ? ct_add<HS_PROBE_BITS>(sh, tabAddr, rel, sh.mult[seg], 0u) : 0u;
			}
		}
	}
	__syncthreads();
	const uint32_t cRej = sh.cRej;
#pragma unroll
	for (int i = 0; i < HS_ITEMS; ++i) {
		if (made[i]) {
			const uint32_t v = lds_u32m(made[i]);
			const uint32_t c = v & 63u;
			if (c > cRej) emit_candidate<U, false>(sh, qn, passLo + (v >> 6), c);
		}
	}
	__syncthreads();
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
	const uint32_t nFast = counters[0];
	const uint32_t k = limit;
	if (tid < 4) sh.win[HS_WIN_ALLOC - 4 + tid] = INF;
	for (int i = tid; i < HS_CT; i += HS_THREADS) sh.ct[i] = HS_EMPTY;
	if (tid == 0) {
		sh.ix = ix;
		sh.buf = reinterpret_cast<unsigned long long*>(hs_smem + ((sizeof(HashShared) + 15) & ~size_t(15)));
		sh.bufCap = bufCap;
		mbar_init(&sh.bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	uint32_t parity = 0;

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
				if ((uint32_t)tid < nLists) {
					sh.listOff[tid] = d.listOff[tid];
					sh.cur[tid] = 0; sh.rem[tid] = d.listLen[tid]; sh.mult[tid] = d.listMult[tid]; sh.nTake[tid] = 0;
				} else { sh.rem[tid] = 0; sh.mult[tid] = 0; sh.nTake[tid] = 0; sh.cur[tid] = 0; sh.listOff[tid] = 0; }
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
				sh.bufCount = 0; sh.overflow = 0; sh.tau = ~0ull; sh.passLo = 0; sh.passHi = 0; sh.prefetched = 0; sh.issuedNext = 0; sh.lastPass = 0;
				sh.listCount = 0;
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

		bool failed = false;
		uint32_t level = 0;
		// warp 0: advances the cursors past the pass that was just planned (lists consumed: nTake) and stages the next pass
		// starting at term id sh.passHi: pass kind, window layout, TMA copies. `early` = called from inside a FILTER pass.
		auto issue = [&](bool early) {
			if (warp != 0) return;
			uint32_t c = 0, r = 0;
			if ((uint32_t)lane < nLists) {
				c = sh.cur[lane] + sh.nTake[lane]; r = sh.rem[lane] - sh.nTake[lane];
				sh.cur[lane] = c; sh.rem[lane] = r; sh.nTake[lane] = 0;
			}
			const unsigned long long tau = sh.tau;
			const uint32_t nextLo = sh.passHi;
			// single-list terms still matter below tau's kid while tau sits in the count-1 class: PROBE pass. Never true after
			// a FILTER pass (tau only improves).
			const bool probe = !early && !caseA && (uint32_t)(tau >> 32) == sh.exactKey[1] && nextLo < (uint32_t)tau;
			uint32_t cap = 4u * (uint32_t)(HS_GROUPS - 2 * (int)nLists);
			if (probe) cap = min(cap, bufCap - level);           // a PROBE posting appends at most one candidate (kid == term id)
			const bool any = sh.lastPass ? false : plan_and_issue(sh, c, r, cap, ix.postings, lane);
			if (lane == 0) { sh.prefetched = any; sh.issuedNext = any; sh.nextMode = probe; }
		};
		issue(false);
		while (true) {
			__syncthreads();
			if (!sh.prefetched) break;
			mbar_wait(&sh.bar, parity);
			parity ^= 1;
			// ---- warp 0: pass bound, consumed counts, sentinels, group segments, pre-filter
			if (warp == 0) {
				const PassPlan& pl = sh.plan;
				const uint32_t wmax = pl.wmax;
				uint32_t* wb = sh.win;
				const uint32_t passLo = sh.passLo;
				const uint32_t mode = sh.nextMode;
				uint32_t bound = INF, wbase = 0, avail = 0;
				const bool on = (uint32_t)lane < nLists;
				if (on) {
					wbase = pl.wbase[lane]; avail = pl.avail[lane];
					if (pl.more[lane]) bound = wb[wbase + wmax];
				}
				for (int o = 16; o; o >>= 1) bound = min(bound, __shfl_xor_sync(0xffffffffu, bound, o));
				// keys hold rel in HS_REL_BITS bits; a PROBE pass ends where single-list terms stop mattering (tau's kid)
				unsigned long long clip = (unsigned long long)passLo + (1ull << HS_REL_BITS);
				if (mode) clip = min(clip, (unsigned long long)(uint32_t)sh.tau);
				bool clipped = false;
				if (clip < (unsigned long long)bound) { bound = (uint32_t)clip; clipped = true; }
				uint32_t nt = 0, last = 0;
				if (on) {
					nt = bound == INF ? avail : lower_bound_smem(wb + wbase, avail, bound);
					sh.nTake[lane] = nt;
					sh.winByte[lane] = wbase * 4;
					if (nt) last = wb[wbase + nt - 1];
					// the rest of the last group: the stage buffer is refilled every pass, so overwriting is harmless
					wb[wbase + nt] = INF; wb[wbase + nt + 1] = INF; wb[wbase + nt + 2] = INF; wb[wbase + nt + 3] = INF;
					// alignment slots in front of a list's first posting hold another gram's postings
					if (sh.cur[lane] == 0) for (uint32_t j = wbase & ~3u; j < wbase; ++j) wb[j] = INF;
				}
				const bool remains = __any_sync(0xffffffffu, on && sh.rem[lane] > nt);
				uint32_t best = nt;
				for (int o = 16; o; o >>= 1) {
					best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
					last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
				}
				// flat group index space over the windows: list i owns groups [segStart[i], segStart[i+1])
				const uint32_t shift = wbase & 3u;
				const uint32_t ng = (on && nt) ? (shift + nt + 3) >> 2 : 0u;
				uint32_t inc = ng;
				for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
				const uint32_t gTot = __shfl_sync(0xffffffffu, inc, 31);
				if (on) {
					sh.segStart[lane] = inc - ng;
					sh.segAddr[lane] = smem_u32(&wb[wbase - shift]) - (inc - ng) * 16;
				}
				if (lane == 0) {
					sh.segStart[nLists] = gTot; sh.segStart[nLists + 1] = INF; sh.segAddr[nLists] = 0; sh.gTot = gTot;
					sh.passHi = remains ? bound : (best ? last + 1 : passLo);
					sh.lastPass = !remains;
					sh.mode = mode;
					sh.issuedNext = 0;
					uint32_t rej = sh.cmin ? sh.cmin - 1 : 0;
					const unsigned long long tau = sh.tau;
					const unsigned long long lowKid = classMode ? passLo : 0u;
					for (uint32_t c = rej + 1; c <= nGrams; ++c) {
						if (((((unsigned long long)sh.hiKey[c]) << 32) | lowKid) >= tau) rej = c; else break;
					}
					sh.cRej = rej;
					if (best == 0 && remains && !clipped) sh.overflow = 2;   // no list can advance (corrupt postings): never spin
					if (!mode && rej < sh.maxMult) sh.overflow = 2;            // cannot happen; a FILTER pass would miss terms
				}
			}
			__syncthreads();
			if (sh.overflow) { failed = true; break; }
			const bool probe = sh.mode != 0;
			if (probe) probe_pass<U>(sh, qn);
			else hash_pass<U>(sh, qn, [&] { issue(true); });
			if (sh.overflow) { failed = true; break; }
			level = sh.bufCount;
			const uint32_t nextLo = sh.passHi;
			const bool lastPass = sh.lastPass != 0;
			__syncthreads();
			if (tid == 0) { sh.passLo = nextLo; sh.listCount = 0; }
			if (level > k && (sh.tau == ~0ull || level > bufCap / 2)) { select_smallest<HS_THREADS>(sh, level, k); level = k; }
			if (lastPass) break;
			if (probe) { __syncthreads(); issue(false); }
		}
		if (failed) {
			// a FILTER pass that gives up may already have staged its successor: let the copies land before the buffer is reused
			__syncthreads();
			if (sh.issuedNext) { mbar_wait(&sh.bar, parity); parity ^= 1; }
			for (int i = tid; i < HS_CT; i += HS_THREADS) sh.ct[i] = HS_EMPTY;
			if (tid == 0) {
				fastList2[atomicAdd(&counters[5], 1u)] = q;
				atomicAdd(&counters[7], 1u);
			}
			continue;
		}
		// ---- final: top-k, sorted
		__syncthreads();
		level = sh.bufCount;
		if (level > k) { select_smallest<HS_THREADS>(sh, level, k); level = k; }
		sort_buffer<HS_THREADS>(sh, level);
		for (uint32_t i = tid; i < level; i += HS_THREADS) {
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
