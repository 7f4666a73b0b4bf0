// fast_path.cuh — k_fast_search: the fused posting-merge / score / top-k kernel (included by search.cu).
//
// One persistent CTA per SM slot; queries are handed out by an atomic cursor. For one query:
//   * the K posting lists of its distinct grams are consumed in PASSES aligned on term-id ranges: every list
//     contributes a window of at most W = cap/K postings, the pass covers the term ids below
//     b = min over lists of (first posting NOT in its window), so every occurrence of a term falls in one pass;
//   * windows are staged in shared memory by TMA 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx) issued by
//     warp 0 (one copy per list, 16-byte aligned source, completion on one mbarrier); two CTAs per SM overlap one
//     CTA's copies with the other's merge, which is why the windows are not double buffered (a second buffer would
//     halve the window and raise the per-pass search overhead for a ~3 % latency gain);
//   * each thread owns an equal slice of the pass's term-id range, locates it in all K windows with branch-free
//     binary searches run in lock step, and K-way merges the windows in registers -> (term, match count), ascending,
//     without atomics or per-term counters (the reference increments a hash map per posting, nGramSearch.hpp:289-298);
//   * counts are filtered by an integer bound derived from the running k-th best (tau); survivors are scored with the
//     reference's float operations and appended to a shared-memory buffer that an in-CTA radix select keeps at <= k;
//   * the final <= limit entries are bitonic-sorted by (score desc, kid asc) and written out.
#pragma once

namespace {

constexpr uint32_t INF = 0xFFFFFFFFu;
constexpr int FS_THREADS = 128;
constexpr int FS_WIN = 6144;                            // posting slots per staged pass
constexpr int FS_DEF_CAP = 128;                         // deferred (term, count) pairs per pass: candidates whose evaluation needs edges / string compares
constexpr int FS_DEF_EXT = 1024;                        // ... on indexes with descriptions / weights, where every survivor is deferred (own array)
constexpr int FS_REGION_PAD = 12;                        // per-list slack: 16-byte alignment shift (<=3) + bound slot + rounding
constexpr int FS_WIN_ALLOC = FS_WIN + MAX_QGRAMS * FS_REGION_PAD + 16;
constexpr int FS_BUF_MAX = 2 * FAST_KLIM;                // candidate buffer entries: 2 * limit rounded up to a power of two, chosen at launch
constexpr int FS_FIRST_PASS = 2048;                      // small first pass: establishes tau early

// Pinned result arrays of the caller (device-accessible host memory): with a fixed number of results per query the CSR layout
// of an ssb_batch is the fixed-stride layout, so every CTA can store a finished query straight to the host while the others
// keep searching — the result copy overlaps the kernel instead of following it. kids == nullptr: off.
struct HostOut {
	uint32_t* kids;
	float* scores;
	uint64_t* rows;
	uint64_t* strs;
	uint64_t strBase;
};

struct PassPlan {
	uint32_t wbase[MAX_QGRAMS];     // element offset of list i's window inside the stage buffer
	uint32_t avail[MAX_QGRAMS];     // postings of list i that may be consumed in this pass: min(rem, wmax)
	uint32_t more[MAX_QGRAMS];      // 1 when the list continues past its window (then window[wmax] is its bound)
	uint32_t wmax;
};

struct FastShared {
	alignas(16) uint32_t win[FS_WIN_ALLOC];
	unsigned long long* buf;                            // candidate buffer: bufCap entries behind this struct in dynamic smem
	uint32_t bufCap;
	uint32_t defCount, defCap;
	unsigned long long* defList;                        // = deferred (aliases hist) or FS_DEF_EXT entries behind the candidate buffer
	alignas(8) unsigned long long bar;                  // mbarrier of the stage buffer (TMA completion)
	union {
		unsigned long long deferred[FS_DEF_CAP];        // (count << 32 | term) pairs: live from the merge loop of a pass to the evaluation right after it
		uint32_t hist[256];                             // live inside select_smallest
	};
	PassPlan plan;
	uint32_t cur[MAX_QGRAMS], rem[MAX_QGRAMS], nTake[MAX_QGRAMS], mult[MAX_QGRAMS], listOff[MAX_QGRAMS];
	uint32_t winByte[MAX_QGRAMS];                        // byte offset of list i's window from &win[0] (current pass)
	uint64_t qcode[MAX_QGRAMS];
	uint32_t hiKey[MAX_QGRAMS + 1];    // ~ordered(upper bound of the value of any edge of a term with count c) incl. promotion
	uint32_t exactKey[MAX_QGRAMS + 1]; // ~ordered(value) of a weight-1, unpromoted candidate with count c
	float sOf[MAX_QGRAMS + 1];         // term score for count c
	unsigned long long tau;
	uint32_t promotedKey;              // key high word of an exact match (score 100)
	uint32_t bufCount, overflow, query, selPrefixLo, selK;
	uint32_t nGe2;                     // candidates appended so far with a match count >= 2 (uniform-weight mode)
	uint32_t passLo, passHi, searchTop, cRej, lastPass, prefetched;
	uint32_t qlen, nGrams, cmin, cprom, allOnes, nLists, total;
	uint32_t shortActive, scanLong;    // the edit-distance path runs for this query / it also covers every long term
	uint32_t scanLongOnly;             // internal: value_for_term() evaluates only the edit-distance path (unused, 0)
	float threshold;
	int topShift, kidBits;             // radix-select digit range (see select_smallest)
	uint32_t classMode;                // 1: key high word = score-class rank (uniform weights), 0: ~ordered(score)
	IndexView ix;
};

// ---- mbarrier / TMA bulk copy (sm_90+ PTX; SASS: SYNCS.*, UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, unsigned long long* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
	             "l"(src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
	uint32_t ok = 0;
	for (uint32_t spin = 0; !ok; ++spin) {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(ok)
		             : "r"(smem_u32(bar)), "r"(parity)
		             : "memory");
		if (spin > (1u << 24)) __trap();      // a lost TMA completion must never hang the GPU
	}
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}

// if (head == m) { pos += 4; head = *pos; } as two predicated instructions that update the cursor in place (the C++ form
// made ptxas compute pos + 4 into a temporary and copy it back: one extra MOV per list and term)
__device__ __forceinline__ void advance_if_head(uint32_t& pos, uint32_t& head, uint32_t m)
{
	asm("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %1, %2;\n\t@p add.u32 %0, %0, 4;\n\t@p ld.shared.u32 %1, [%0];\n\t}" : "+r"(pos), "+r"(head) : "r"(m) : "memory");
}

__device__ __forceinline__ uint32_t lower_bound_smem(const uint32_t* a, uint32_t n, uint32_t x)
{
	uint32_t lo = 0, hi = n;
	while (lo < hi) {
		uint32_t mid = (lo + hi) >> 1;
		if (a[mid] < x) lo = mid + 1; else hi = mid;
	}
	return lo;
}

// barrier of the NT threads that call it: the whole CTA (BAR == 0) or a named barrier shared by a subset of its warps
template <int BAR, int NT>
__device__ __forceinline__ void group_sync()
{
	if (BAR == 0) __syncthreads();
	else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NT) : "memory");
}

// keeps the k smallest entries of buf[0..n) (all distinct), sets tau to (an upper bound without any entry in between of) the k-th
// smallest. CTA-wide. MSD radix select over 8-bit digits; digits known to be zero in every entry are skipped: the high word
// is a score-class rank < 256 in uniform-weight mode (topShift = 32), and kids have kidBits significant bits. Two barriers
// per digit (warp 0 re-zeroes the bins it has just scanned); the select stops as soon as the bin that holds the k-th entry is
// needed as a whole.
template <int NT, int BAR = 0, class SH>
__device__ void select_smallest(SH& sh, uint32_t n, uint32_t k)
{
	const int tid = threadIdx.x;
	unsigned long long prefix = 0;
	uint32_t kk = k;
	const int topShift = sh.topShift, kidBits = sh.kidBits;
	for (int i = tid; i < 256; i += NT) sh.hist[i] = 0;
	group_sync<BAR, NT>();
	int shift = topShift;
	for (; shift >= 0; shift -= 8) {
		if (shift < 32 && shift >= kidBits) { prefix <<= 8; continue; }      // this kid byte is zero everywhere
		for (uint32_t i = tid; i < n; i += NT) {
			unsigned long long v = sh.buf[i];
			if (shift == topShift || (v >> (shift + 8)) == prefix) atomicAdd(&sh.hist[(uint32_t)(v >> shift) & 255u], 1u);
		}
		group_sync<BAR, NT>();
		if (tid < 32) {
			uint32_t loc[8], sum = 0;
#pragma unroll
			for (int j = 0; j < 8; ++j) { loc[j] = sh.hist[tid * 8 + j]; sh.hist[tid * 8 + j] = 0; sum += loc[j]; }
			uint32_t inc = sum;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += t; }
			uint32_t exc = inc - sum;
			if (kk > exc && kk <= inc) {
				uint32_t run = exc;
#pragma unroll
				for (int j = 0; j < 8; ++j) {
					if (kk > run && kk <= run + loc[j]) {
						sh.selPrefixLo = tid * 8 + j;
						sh.selK = (kk - run == loc[j]) ? 0u : kk - run;     // 0: every entry of this bin is among the k smallest
					}
					run += loc[j];
				}
			}
		}
		group_sync<BAR, NT>();
		prefix = (prefix << 8) | sh.selPrefixLo;
		kk = sh.selK;
		if (kk == 0) break;
	}
	// stopped early: everything that starts with `prefix` is kept -> the bound is prefix followed by ones
	if (shift > 0) prefix = (prefix << shift) | ((1ull << shift) - 1);
	unsigned long long keep[FS_BUF_MAX / NT];
	uint32_t nk = 0;
	for (uint32_t i = tid; i < n; i += NT) {
		unsigned long long v = sh.buf[i];
		if (v <= prefix) keep[nk++] = v;
	}
	if (tid == 0) { sh.bufCount = 0; sh.tau = prefix; }
	group_sync<BAR, NT>();
	for (uint32_t j = 0; j < nk; ++j) sh.buf[atomicAdd(&sh.bufCount, 1u)] = keep[j];
	group_sync<BAR, NT>();
}

// ascending bitonic sort of buf[0..n), n <= bufCap. CTA-wide.
template <int NT, int BAR = 0, class SH>
__device__ void sort_buffer(SH& sh, uint32_t n)
{
	if (n <= 2u * NT) {
		// short lists (the usual top-100): every thread ranks its (at most two) entries against all others — the entries are
		// distinct, so the rank is the position — two barriers instead of one per bitonic stage
		const uint32_t i0 = threadIdx.x, i1 = threadIdx.x + NT;
		const unsigned long long k0 = i0 < n ? sh.buf[i0] : ~0ull, k1 = i1 < n ? sh.buf[i1] : ~0ull;
		uint32_t r0 = 0, r1 = 0;
		for (uint32_t j = 0; j < n; ++j) {
			const unsigned long long v = sh.buf[j];
			r0 += v < k0; r1 += v < k1;
		}
		group_sync<BAR, NT>();
		if (i0 < n) sh.buf[r0] = k0;
		if (i1 < n) sh.buf[r1] = k1;
		group_sync<BAR, NT>();
		return;
	}
	uint32_t p = 1;
	while (p < n) p <<= 1;
	for (uint32_t i = n + threadIdx.x; i < p; i += NT) sh.buf[i] = ~0ull;
	group_sync<BAR, NT>();
	for (uint32_t size = 2; size <= p; size <<= 1)
		for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
			for (uint32_t i = threadIdx.x; i < (p >> 1); i += NT) {
				uint32_t lo = 2 * i - (i & (stride - 1));
				uint32_t hi = lo + stride;
				bool up = (lo & size) == 0;
				unsigned long long a = sh.buf[lo], b = sh.buf[hi];
				if ((a > b) == up) { sh.buf[lo] = b; sh.buf[hi] = a; }
			}
			group_sync<BAR, NT>();
		}
}

// appends one key per calling lane. Appends are rare events by design (a handful of lanes at a time once tau is established),
// so one shared-memory atomic per lane is cheaper than electing a leader and shuffling offsets around.
template <class SH>
__device__ __forceinline__ void append_key(SH& sh, unsigned long long key)
{
	const uint32_t idx = atomicAdd(&sh.bufCount, 1u);
	if (idx < sh.bufCap) sh.buf[idx] = key; else sh.overflow = 1;
}

// Value that term t2 contributes to key `kid` through an edge of weight w2, over every path that applies to the term for
// this query: the gram path for long terms (hpp:278-301) and the edit-distance path for short terms — and for long terms too
// when |q| <= gSize (hpp:240-253). Returns -1 when the term does not reach the key (no hit / below the threshold).
template <class U, bool SHORT, class SH>
__device__ float value_for_term(SH& sh, const U* __restrict__ qn, uint32_t t2, uint32_t kid, float w2, int& keyEq)
{
	const IndexView& ix = sh.ix;
	const uint32_t len2 = ix.termOff[t2 + 1] - ix.termOff[t2];
	const bool isLong = len2 >= 2u * ix.g;
	const float th = sh.threshold;
	float best = -1.0f;
	if (isLong && sh.nGrams >= 1 && !sh.scanLongOnly) {
		const uint32_t c2 = count_for_term<U>(ix, t2, sh.qcode, sh.nGrams);
		if (c2) {
			const float s2 = sh.sOf[c2];
			if (!(s2 < th)) {
				float v = edge_value(w2, s2);
				if (c2 >= sh.cprom) {
					if (keyEq < 0) keyEq = key_equals_query<U>(ix, kid, qn, sh.qlen);
					if (keyEq) v = 100.0f;
				}
				best = fmaxf(best, v);
			}
		}
	}
	if (SHORT && sh.shortActive && (!isLong || sh.scanLong)) {
		const U* tc = reinterpret_cast<const U*>(ix.termChars) + ix.termOff[t2];
		const uint32_t m2 = string_match<U>(qn, sh.qlen, tc, len2);
		const float s2 = __fdiv_rn((float)m2, (float)sh.qlen);
		if (!(s2 < th)) {
			float v = edge_value(w2, s2);
			if ((double)s2 > 0.999) {
				if (keyEq < 0) keyEq = key_equals_query<U>(ix, kid, qn, sh.qlen);
				if (keyEq) v = 100.0f;
			}
			best = fmaxf(best, v);
		}
	}
	return best;
}

// Emits (kid, v) for the edge (term m -> kid) unless another edge of the same key gives it a better value: a key reachable
// through several terms appears once, with its best value (per-key max-merge, hpp:326; ties -> smallest term id).
template <class U, bool SHORT, class SH>
__device__ void emit_edge(SH& sh, const U* __restrict__ qn, uint32_t m, uint32_t kid, float v, int keyEq)
{
	const IndexView& ix = sh.ix;
	const unsigned long long key = rank_key(v, kid);
	if (key >= sh.tau) return;
	const uint32_t kb = ix.keyEdgeOff[kid], ke = ix.keyEdgeOff[kid + 1];
	if (ke - kb > 1) {
		for (uint32_t x = kb; x < ke; ++x) {
			const uint32_t t2 = ix.keyEdgeTerm[x];
			if (t2 == m) continue;
			const float v2 = value_for_term<U, SHORT>(sh, qn, t2, kid, ix.keyEdgeW[x], keyEq);
			if (v2 > v || (v2 == v && t2 < m)) return;           // dominated
		}
	}
	append_key(sh, key);
}

// A long term reached `c` gram matches: evaluate its keys (hpp:310-341) and append what can still make the top-k.
// Called only for terms that survive the integer pre-filter, i.e. rarely once tau is established -> not inlined.
template <class U, bool SHORT, class SH>
__device__ __noinline__ void emit_candidate(SH& sh, const U* __restrict__ qn, uint32_t m, uint32_t c)
{
	const IndexView& ix = sh.ix;
	const unsigned long long tau = sh.tau;
	const uint32_t qlen = sh.qlen, cprom = sh.cprom;
	if (ix.identity && ix.uniformW) {
		const uint32_t kid = m;
		uint32_t hk = sh.exactKey[c];
		if (c >= cprom && key_equals_query<U>(ix, kid, qn, qlen)) hk = sh.promotedKey;              // hpp:333-335
		const unsigned long long key = ((unsigned long long)hk << 32) | kid;
		if (key < tau) append_key(sh, key);
		return;
	}
	const float s = sh.sOf[c];
	for (uint32_t e = ix.termEdgeOff[m]; e < ix.termEdgeOff[m + 1]; ++e) {           // hpp:321-338
		const uint32_t kid = ix.edgeKid[e];
		float v = edge_value(ix.edgeW[e], s);
		int keyEq = -1;
		if (c >= cprom) {
			keyEq = key_equals_query<U>(ix, kid, qn, qlen);
			if (keyEq) v = 100.0f;
		}
		emit_edge<U, SHORT>(sh, qn, m, kid, v, keyEq);
	}
}

// Edit-distance path (hpp:232-254) for one term: every scanned term scores match/|q| (zeros included); when the scan covers
// long terms (|q| <= gSize) their gram-path score is folded in here, so that a (term, key) pair is emitted exactly once.
template <class U>
__device__ __noinline__ void scan_term(FastShared& sh, const U* __restrict__ qn, uint32_t t)
{
	const IndexView& ix = sh.ix;
	for (uint32_t e = ix.termEdgeOff[t]; e < ix.termEdgeOff[t + 1]; ++e) {
		const uint32_t kid = ix.edgeKid[e];
		int keyEq = -1;
		const float v = value_for_term<U, true>(sh, qn, t, kid, ix.edgeW[e], keyEq);
		if (v < 0.0f) continue;
		emit_edge<U, true>(sh, qn, t, kid, v, keyEq);
	}
}

// One pass over the staged windows: see the file comment.
template <class U, int KT, bool ONES, bool SHORT>
__device__ __forceinline__ void merge_pass(FastShared& sh, const U* __restrict__ qn, uint32_t nLists)
{
	const int tid = threadIdx.x;
	const unsigned long long span = (unsigned long long)sh.passHi - sh.passLo;
	const uint32_t lo = sh.passLo + (uint32_t)((span * (unsigned)tid) / FS_THREADS);
	const uint32_t hi = (tid == FS_THREADS - 1) ? sh.passHi : sh.passLo + (uint32_t)((span * (unsigned)(tid + 1)) / FS_THREADS);
	if (lo >= hi) return;
	const uint32_t winAddr = smem_u32(&sh.win[0]);
	uint32_t pos[KT], head[KT], mu[ONES ? 1 : KT];       // pos: absolute shared-memory byte addresses (one LDS, no address math)
	{
		uint32_t end[KT];
#pragma unroll
		for (int i = 0; i < KT; ++i) {
			const bool on = (uint32_t)i < nLists;
			// template slots beyond K are empty windows that read the INF slot
			pos[i] = winAddr + (on ? sh.winByte[i] : (uint32_t)((FS_WIN_ALLOC - 1) * 4));
			end[i] = pos[i] + (on ? sh.nTake[i] * 4 : 0u);
			if (!ONES) mu[i] = sh.mult[i];
		}
		// lower_bound(lo) in every window: branch-free, all K searches advance in lock step (K independent LDS per step)
		for (uint32_t st = sh.searchTop * 4; st >= 4; st >>= 1) {
#pragma unroll
			for (int i = 0; i < KT; ++i) {
				const uint32_t q = pos[i] + st;
				if (q <= end[i] && lds_u32(q - 4) < lo) pos[i] = q;
			}
		}
#pragma unroll
		for (int i = 0; i < KT; ++i) head[i] = lds_u32(pos[i]);
	}
	// integer pre-filter, refined per thread in identity mode: every kid of this slice is >= lo, so a whole score class
	// that ties with tau but lies behind tau's kid can be rejected without evaluating its candidates
	uint32_t cRej = sh.cRej;
	if (sh.ix.identity && sh.ix.uniformW) {
		const unsigned long long tau = sh.tau;
		const uint32_t nG = sh.nGrams;
		while (cRej < nG && ((((unsigned long long)sh.hiKey[cRej + 1]) << 32) | lo) >= tau) ++cRej;
	}
	// The hot loop makes no calls: a surviving term is either appended right here (identity + uniform weights: kid == term id,
	// value known from the count) or parked in the deferred list and evaluated after the loop (edges, weights, exact-match
	// string compare, per-key max-merge). A call inside the loop costs the merge its registers (spilled heads).
	const bool identMode = sh.ix.identity && sh.ix.uniformW;
	const unsigned long long tauNow = sh.tau;
	const uint32_t cpromNow = sh.cprom;
	uint32_t posSumPrev = 0;
#pragma unroll
	for (int i = 0; i < KT; ++i) posSumPrev += pos[i];
	while (true) {
		uint32_t m = head[0];
#pragma unroll
		for (int i = 1; i < KT; ++i) m = min(m, head[i]);
		if (m >= hi) break;
		uint32_t c = 0;
		if (ONES) {
			// count = number of cursors that advance = (sum of the cursor addresses now - after the previous term) / 4: the
			// adds fold into 3-input IADD3s (cheaper than one predicated increment per list), and carrying the sum from term
			// to term means the cursors are updated in place (a before/after pair made the compiler keep two copies of
			// every cursor: ~2 MOVs per list and term)
#pragma unroll
			for (int i = 0; i < KT; ++i) advance_if_head(pos[i], head[i], m);
			uint32_t posSum = 0;
#pragma unroll
			for (int i = 0; i < KT; ++i) posSum += pos[i];
			c = (posSum - posSumPrev) >> 2;
			posSumPrev = posSum;
		} else {
#pragma unroll
			for (int i = 0; i < KT; ++i) {
				if (head[i] == m) {
					c += mu[i];
					pos[i] += 4;
					head[i] = lds_u32(pos[i]);
				}
			}
		}
		if (c > cRej) {
			if (identMode && c < cpromNow) {
				const unsigned long long key = (((unsigned long long)sh.exactKey[c]) << 32) | m;
				if (key < tauNow) {
					append_key(sh, key);
					if (c >= 2) atomicAdd(&sh.nGe2, 1u);
				}
			} else {
				const uint32_t di = atomicAdd(&sh.defCount, 1u);
				if (di < sh.defCap) sh.defList[di] = (((unsigned long long)c) << 32) | m; else sh.overflow = 1;
			}
		}
	}
}

template <class U, bool ONES, bool SHORT>
__device__ __forceinline__ void merge_dispatch(FastShared& sh, const U* __restrict__ qn, uint32_t nLists)
{
	switch (nLists) {
	case 1: merge_pass<U, 1, ONES, SHORT>(sh, qn, nLists); break;
	case 2: merge_pass<U, 2, ONES, SHORT>(sh, qn, nLists); break;
	case 3: merge_pass<U, 3, ONES, SHORT>(sh, qn, nLists); break;
	case 4: merge_pass<U, 4, ONES, SHORT>(sh, qn, nLists); break;
	case 5: merge_pass<U, 5, ONES, SHORT>(sh, qn, nLists); break;
	case 6: merge_pass<U, 6, ONES, SHORT>(sh, qn, nLists); break;
	case 7: merge_pass<U, 7, ONES, SHORT>(sh, qn, nLists); break;
	case 8: merge_pass<U, 8, ONES, SHORT>(sh, qn, nLists); break;
	case 9: merge_pass<U, 9, ONES, SHORT>(sh, qn, nLists); break;
	case 10: merge_pass<U, 10, ONES, SHORT>(sh, qn, nLists); break;
	case 11: merge_pass<U, 11, ONES, SHORT>(sh, qn, nLists); break;
	case 12: merge_pass<U, 12, ONES, SHORT>(sh, qn, nLists); break;
	case 13: merge_pass<U, 13, ONES, SHORT>(sh, qn, nLists); break;
	case 14: merge_pass<U, 14, ONES, SHORT>(sh, qn, nLists); break;
	case 15: case 16: merge_pass<U, 16, ONES, SHORT>(sh, qn, nLists); break;
	case 17: case 18: case 19: case 20: merge_pass<U, 20, ONES, SHORT>(sh, qn, nLists); break;
	case 21: case 22: case 23: case 24: merge_pass<U, 24, ONES, SHORT>(sh, qn, nLists); break;
	default: merge_pass<U, 32, ONES, SHORT>(sh, qn, nLists); break;
	}
}

// warp 0: lays out the windows of one pass in the stage buffer from the cursors (cur, rem) and issues the TMA copies.
// Returns (uniformly over the warp) whether anything was issued.
template <class SH>
__device__ __forceinline__ bool plan_and_issue(SH& sh, uint32_t cur, uint32_t rem, uint32_t cap, const uint32_t* postings, int lane)
{
	const uint32_t nLists = sh.nLists;
	const uint32_t wmax = max(1u, cap / nLists);
	const uint32_t region = (wmax + 10) & ~3u;                     // <= 3 shift + (wmax + 1) data + round-up to 16 bytes
	PassPlan& pl = sh.plan;
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
			dst = &sh.win[lane * region];
		}
	}
	uint32_t total = bytes;
	for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
	if (lane == 0) pl.wmax = wmax;
	if (total == 0) return false;
	if (lane == 0) {
		// the buffer was read/written through the generic proxy by the previous pass: order that before the async writes
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		mbar_expect_tx(&sh.bar, total);
	}
	__syncwarp();
	if (bytes) tma_load_1d(dst, src, bytes, &sh.bar);
	return true;
}

// SHORT = the index has short terms, so queries shorter than 3*gSize also run the edit-distance scan; the lean variant
// (no such terms, e.g. the benchmark vocabulary) is compiled without that code to keep the merge loops' registers.
template <class U, bool SHORT>
__global__ void __launch_bounds__(FS_THREADS, 6) k_fast_search(IndexView ix, const QDesc* __restrict__ desc, const U* __restrict__ qnorm,
                                                            const uint32_t* __restrict__ fastList, uint32_t* __restrict__ counters,
                                                            uint32_t* __restrict__ denseList, uint32_t limit, uint32_t outStride,
                                                            uint32_t* __restrict__ outCount, uint32_t* __restrict__ outKid,
                                                            float* __restrict__ outScore, unsigned long long* __restrict__ statCounters,
                                                            uint32_t bufCap, float threshold, uint32_t nIdx, uint32_t curIdx, uint32_t defExt,
                                                            HostOut ho)
{
	// nIdx / curIdx: which counters hold this launch's list length and work cursor (the list is either every fast query or
	// the ones k_sieve_search handed over)
	extern __shared__ __align__(16) unsigned char fs_smem[];
	FastShared& sh = *reinterpret_cast<FastShared*>(fs_smem);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t nFast = counters[nIdx];
	const uint32_t k = limit;
	if (tid == 0) {
		sh.win[FS_WIN_ALLOC - 1] = INF;
		sh.ix = ix;
		sh.buf = reinterpret_cast<unsigned long long*>(fs_smem + ((sizeof(FastShared) + 15) & ~size_t(15)));
		sh.bufCap = bufCap;
		sh.defList = defExt ? sh.buf + bufCap : sh.deferred;
		sh.defCap = defExt ? defExt : (uint32_t)FS_DEF_CAP;
		mbar_init(&sh.bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	uint32_t parity = 0;                    // phase parity of the stage barrier (uniform across the CTA)

	while (true) {
		__syncthreads();
		if (tid == 0) sh.query = atomicAdd(&counters[curIdx], 1u);
		__syncthreads();
		const uint32_t wi = sh.query;
		if (wi >= nFast) break;
		const uint32_t q = fastList[wi];
		const QDesc& d = desc[q];
		const U* qn = qnorm + d.qbeg;
		const uint32_t nLists = d.nLists, nGrams = d.nGrams;
		if (tid < MAX_QGRAMS) {
			sh.qcode[tid] = (uint32_t)tid < nGrams ? d.qcode[tid] : ~0ull;
			if ((uint32_t)tid < nLists) {
				sh.listOff[tid] = d.listOff[tid];
				sh.cur[tid] = 0; sh.rem[tid] = d.listLen[tid]; sh.mult[tid] = d.listMult[tid];
			} else { sh.rem[tid] = 0; sh.mult[tid] = 0; sh.nTake[tid] = 0; sh.cur[tid] = 0; sh.listOff[tid] = 0; }
		}
		// identity + uniform weights: a candidate's value is fully determined by its match count, so the key's high word is the
		// RANK of the score class (0 = exact match promoted to 100, then nGrams matches, ..., 1 match) instead of the float bits
		// — same order, 8 significant bits for the radix select. Otherwise: ~ordered(float value).
		const bool shortActive = SHORT && (d.flags & 1u) != 0, scanLong = SHORT && (d.flags & 2u) != 0;
		const bool classMode = ix.identity && ix.uniformW && !shortActive;
		if (tid <= MAX_QGRAMS) {
			const uint32_t c = tid;
			float s = (c <= nGrams && nGrams) ? term_score(c, nGrams) : 0.0f;
			sh.sOf[c] = s;
			if (classMode) {
				sh.exactKey[c] = c <= nGrams ? nGrams - c + 1 : 255u;
				sh.hiKey[c] = c >= d.cprom ? 0u : sh.exactKey[c];
			} else {
				// an upper bound over all edges of any term with this count
				float ub = ix.uniformW ? edge_value(1.0f, s) : edge_value(ix.wmax, s);
				sh.exactKey[c] = ~float_to_ordered(edge_value(1.0f, s));
				if (c >= d.cprom) ub = fmaxf(ub, 100.0f);
				sh.hiKey[c] = ~float_to_ordered(ub);
			}
		}
		if (tid == 0) {
			sh.bufCount = 0; sh.overflow = 0; sh.tau = ~0ull; sh.passLo = 0; sh.prefetched = 0; sh.nGe2 = 0;
			sh.qlen = d.qlen; sh.nGrams = nGrams; sh.cmin = d.cmin; sh.cprom = d.cprom; sh.nLists = nLists; sh.total = d.total;
			sh.classMode = classMode;
			sh.shortActive = shortActive; sh.scanLong = scanLong; sh.scanLongOnly = 0; sh.threshold = threshold;
			sh.promotedKey = classMode ? 0u : ~float_to_ordered(100.0f);
			sh.topShift = classMode ? 32 : 56;
			int kb = 1;
			while (kb < 32 && (1ull << kb) < ix.K) ++kb;
			sh.kidBits = kb;
			uint32_t ones = 1;
			for (uint32_t i = 0; i < nLists; ++i) ones &= d.listMult[i] == 1;
			sh.allOnes = ones;
		}
		__syncthreads();

		// |q| <= gSize: the term scan below covers every long term and folds the gram path in, so no posting passes run
		bool failed = false, safeMode = false, more = nLists != 0 && !scanLong;
		bool tauReal = false;          // tau comes from a select over real candidates (not merely from the seed below)
		// the second select comes as soon as the buffer holds ~2k spare candidates: the sooner tau reaches the score class the
		// k-th best lives in, the sooner the integer pre-filter stops whole classes. Later selects (long posting lists keep
		// feeding better candidates) wait for a half-full buffer again, a select costs ~10 barriers.
		uint32_t selectAt = min(bufCap / 2, 2 * k + 64);
		uint32_t level = 0;
		uint32_t levelGe2 = 0;         // sh.nGe2 at the start of the pass in flight (restored together with bufCount when a pass is redone)
		// small first pass when tau cannot be seeded: it only has to establish tau, and every survivor of it is a candidate
		// (identity mode: at most bufCap appends; otherwise at most defCap deferred evaluations)
		const uint32_t firstCap = (ix.identity && ix.uniformW) ? min((uint32_t)FS_FIRST_PASS, bufCap) : sh.defCap;
		uint32_t cap = d.total > firstCap ? firstCap : FS_WIN;
		if (ix.identity && ix.uniformW && d.cmin <= 1 && k <= FAST_KLIM && more) {
			// Seed tau without a pass: the first k postings of any one list are k distinct keys (kid == term id) with at least
			// one match each, i.e. k candidates ranking at or before (score(1), kid = list[k-1]). The final k-th best cannot
			// be worse, so everything behind the smallest such bound can be dropped from the start.
			if (warp == 0) {
				uint32_t x = INF;
				if ((uint32_t)lane < nLists && d.listLen[lane] >= k) x = __ldg(ix.postings + d.listOff[lane] + (k - 1));
				for (int o = 16; o; o >>= 1) x = min(x, __shfl_xor_sync(0xffffffffu, x, o));
				if (lane == 0 && x != INF) sh.tau = ((((unsigned long long)sh.exactKey[1]) << 32) | x) + 1;   // "< tau" keeps the bound itself
			}
			__syncthreads();
			// a seeded tau admits ~k single-list terms per list; a short first pass (about 3k postings) finds k real candidates
			// among the smallest ids, the select that follows it replaces the seed by a bound that is ~nLists times tighter
			if (sh.tau != ~0ull) cap = min((uint32_t)FS_WIN, max(3u * k, 256u));
		}
		while (more) {
			// ---- warp 0: lay out the windows from the cursors and issue the TMA copies
			if (warp == 0) {
				const bool any = plan_and_issue(sh, sh.cur[lane & (MAX_QGRAMS - 1)], (uint32_t)lane < nLists ? sh.rem[lane] : 0, cap,
				                                ix.postings, lane);
				if (lane == 0) sh.prefetched = any;
			}
			__syncthreads();
			if (!sh.prefetched) break;
			mbar_wait(&sh.bar, parity);
			parity ^= 1;
			// ---- warp 0: pass bound b = min over lists of the first posting NOT covered; consumed counts; sentinels; pre-filter
			if (warp == 0) {
				const PassPlan& pl = sh.plan;
				const uint32_t wmax = pl.wmax;
				uint32_t* wb = sh.win;
				uint32_t bound = INF, wbase = 0, avail = 0;
				if ((uint32_t)lane < nLists) {
					wbase = pl.wbase[lane]; avail = pl.avail[lane];
					if (pl.more[lane]) bound = wb[wbase + wmax];
				}
				for (int o = 16; o; o >>= 1) bound = min(bound, __shfl_xor_sync(0xffffffffu, bound, o));
				uint32_t nt = 0, last = 0;
				if ((uint32_t)lane < nLists) {
					nt = bound == INF ? avail : lower_bound_smem(wb + wbase, avail, bound);
					sh.nTake[lane] = nt;
					sh.winByte[lane] = wbase * 4;
					if (nt) last = wb[wbase + nt - 1];
					wb[wbase + nt] = INF;
				}
				uint32_t best = nt;
				for (int o = 16; o; o >>= 1) {
					best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
					last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
				}
				// integer pre-filter: a count c <= cRej can neither pass the threshold nor beat the running k-th best. In identity
				// mode kid == term id >= passLo, which makes the bound exact for the tie class of tau. Lane j tests count
				// rej0 + 1 + j (nGrams <= 32 = one warp); cRej = rej0 + the leading run of rejectable counts.
				uint32_t rej = sh.cmin ? sh.cmin - 1 : 0;
				{
					const unsigned long long tau = sh.tau;
					const unsigned long long lowKid = (sh.ix.identity && sh.ix.uniformW) ? sh.passLo : 0u;
					const uint32_t c = rej + 1 + lane;
					const bool out = c <= nGrams && ((((unsigned long long)sh.hiKey[min(c, (uint32_t)MAX_QGRAMS)]) << 32) | lowKid) >= tau;
					const unsigned mOut = __ballot_sync(0xffffffffu, out);
					rej += mOut == 0xffffffffu ? 32u : (uint32_t)__ffs(~mOut) - 1;
				}
				if (lane == 0) {
					sh.passHi = bound == INF ? last + 1 : bound;
					sh.searchTop = best ? 1u << (31 - __clz(best)) : 0u;        // largest power of two <= the longest window
					sh.cRej = rej;
					sh.defCount = 0;
					sh.lastPass = bound == INF;
					if (best == 0) sh.overflow = 2;          // no list can advance (corrupt postings): never spin, fall back
				}
			}
			__syncthreads();
			if (sh.overflow == 2) { failed = true; break; }
			if (tid == 0) atomicAdd(&counters[16], 1u);
			// ---- merge
			if (sh.allOnes) merge_dispatch<U, true, SHORT>(sh, qn, nLists);
			else merge_dispatch<U, false, SHORT>(sh, qn, nLists);
			__syncthreads();
			{
				const uint32_t nd = min(sh.defCount, sh.defCap);
				if (tid == 0 && nd) atomicAdd(&counters[19], nd);
				for (uint32_t i = tid; i < nd; i += FS_THREADS) {
					const unsigned long long e = sh.defList[i];
					emit_candidate<U, SHORT>(sh, qn, (uint32_t)e, (uint32_t)(e >> 32));
				}
				if (nd) __syncthreads();
			}
			if (sh.overflow) {
				// the optimistic pass overflowed the candidate buffer: roll it back, make room and redo it with a window small
				// enough to be safe for single-edge terms; if that overflows too -> dense path
				if (safeMode && level <= k) { failed = true; break; }
				if (tid == 0) atomicAdd(&counters[17], 1u);
				__syncthreads();
				// the redone pass re-merges the same id range: its two-match candidates must not be counted twice
				if (tid == 0) { sh.bufCount = level; sh.nGe2 = levelGe2; sh.overflow = 0; }
				__syncthreads();
				if (level > k) { select_smallest<FS_THREADS>(sh, level, k); level = k; tauReal = true; }
				safeMode = true;
				// (uniform weights + kid == term id: survivors are appended directly, only exact-match suspects are deferred)
				cap = min((uint32_t)FS_WIN, bufCap - level);
				if (!(ix.identity && ix.uniformW)) cap = min(cap, sh.defCap);
				continue;
			}
			safeMode = false;
			cap = FS_WIN;
			// descriptions / weights: while single-gram matches cannot be rejected yet, (nearly) every posting of the pass is
			// deferred -> keep the pass within the deferred list instead of overflowing it and redoing the pass
			if (!(ix.identity && ix.uniformW) && sh.cmin <= 1 && (((unsigned long long)sh.hiKey[1]) << 32) < sh.tau) cap = min(cap, sh.defCap);
			more = sh.lastPass == 0;
			if (tid < MAX_QGRAMS && (uint32_t)tid < nLists) { sh.cur[tid] += sh.nTake[tid]; sh.rem[tid] -= sh.nTake[tid]; }
			level = sh.bufCount;
			levelGe2 = sh.nGe2;
			const uint32_t nextLo = sh.passHi;
			__syncthreads();
			if (tid == 0) sh.passLo = nextLo;
			// establish tau as soon as k candidates exist; afterwards the exact integer pre-filter keeps appends rare, so the
			// (barrier-heavy) select only runs again when the buffer is half full
			if (level > k && !tauReal && sh.classMode && (uint32_t)(sh.tau >> 32) == sh.exactKey[1] && level <= bufCap / 4) {
				// first candidates after a seeded tau (the short first pass): all `level` >= k of them have a kid below this
				// pass's end, so (count-1 class, end of the pass) bounds the k-th best — as good as a select for stopping the
				// single-list terms of the passes to come, and free
				if (tid == 0) {
					const unsigned long long t2 = (((unsigned long long)sh.exactKey[1]) << 32) | nextLo;
					if (t2 < sh.tau) sh.tau = t2;
				}
				tauReal = true;
			} else if (tauReal && sh.classMode && nGrams >= 2 && sh.nGe2 >= k && (uint32_t)(sh.tau >> 32) > sh.exactKey[2] &&
			           level <= bufCap / 2) {
				// k candidates with two or more matches are known, all below this pass's end: (count-2 class, end of the pass)
				// bounds the k-th best — tau moves into the class it will end in without a select, and the passes to come
				// stop their two-match terms in the integer pre-filter
				__syncthreads();                    // every thread has read tau for the test above before it changes class
				if (tid == 0) sh.tau = (((unsigned long long)sh.exactKey[2]) << 32) | nextLo;
				selectAt = bufCap / 2;
			} else if (level > k && (!tauReal || level > selectAt)) {
				if (tauReal) selectAt = bufCap / 2;
				select_smallest<FS_THREADS>(sh, level, k); level = k; tauReal = true;
				if (tid == 0) atomicAdd(&counters[18], 1u);
			}
		}
		// ---- edit-distance path: brute-force scan of the short terms (and of all terms for |q| <= gSize), hpp:232-254
		if (SHORT && !failed && shortActive) {
			const uint32_t nScan = scanLong ? ix.T : ix.nShort;
			for (uint32_t base = 0; base < nScan; base += FS_THREADS) {
				const uint32_t i = base + tid;
				if (i < nScan) scan_term<U>(sh, qn, scanLong ? i : ix.shortTerms[i]);
				__syncthreads();
				if (sh.overflow) { failed = true; break; }
				level = sh.bufCount;
				__syncthreads();
				if (level > k && (!tauReal || level > selectAt)) { select_smallest<FS_THREADS>(sh, level, k); level = k; tauReal = true; }
			}
			if (!failed) level = sh.bufCount;
		}
		if (failed) {
			if (tid == 0) {
				denseList[atomicAdd(&counters[1], 1u)] = q;
				atomicAdd(&counters[3], 1u);
			}
			continue;
		}
		// ---- final: top-k, sorted
		if (level > k) { select_smallest<FS_THREADS>(sh, level, k); level = k; }
		sort_buffer<FS_THREADS>(sh, level);
		for (uint32_t i = tid; i < level; i += FS_THREADS) {
			unsigned long long v = sh.buf[i];
			const uint32_t kid = (uint32_t)v;
			const uint32_t hi = (uint32_t)(v >> 32);
			const float score = sh.classMode ? (hi == 0 ? 100.0f : edge_value(1.0f, sh.sOf[nGrams - hi + 1])) : rank_key_score(v);
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
			if (level != outStride) atomicAdd(&counters[22], 1u);      // a short result list: the host arrays need compaction
		}
	}
}

}   // namespace
