// tile_path.cuh — the tile pipeline: k_sieve_search's algorithm (sieve_path.cuh) re-cut so that the unit of work is a TILE of
// a query's pass instead of a whole query (included by search.cu after sieve_path.cuh, whose device functions it reuses).
//
// k_sieve_search keeps one CTA on one query from set-up to output: passes, their plans, the final select and sort all run
// one after the other behind CTA barriers, and with a few thousand queries per batch a GPU holds only ~1 query per warp
// slot group, so nothing hides those serial stretches (ncu: issue slots 40 % busy, barrier + scoreboard stalls). Here a batch
// advances in ROUNDS, each a pair of launches over every query of the batch at once:
//   k_tile_plan   one warp per query: tau from the counts of the round just finished (k candidates with >= c matches are
//                 known below the end of the pass: the same free bound as in k_sieve_search), the count a term needs from
//                 here on, the next pass, cut into tiles of TILE_G groups that are appended to the round's tile list
//   k_tile_sieve  persistent CTAs take tiles from the list: SIEVE + RESOLVE over the tile's groups exactly as in
//                 k_sieve_search (sieve_groups / resolve_block / resolve_slow), candidates collected in shared memory and
//                 flushed to the query's candidate list in global memory with one atomic per tile
// and after the last round
//   k_tile_final  one CTA per finished query: top-k of its candidate list (radix select), sorted, written out.
// A query's tiles of one pass run concurrently on different SMs under the same tau, so a batch is balanced tile by tile and
// the per-query serial work shrinks to one warp of k_tile_plan per round. Queries the pipeline does not take (weights /
// descriptions, a single gram, the edit-distance path) or does not finish (fewer than k terms with two matches, an
// overflowing buffer, more rounds needed than launched) are handed to k_sieve_search through handList: every path is exact.
#pragma once

namespace {

constexpr int TP_ROUNDS = 3;                            // sieve rounds launched per batch
// slots in SearchWorkspace::counters
constexpr uint32_t TP_COUNT = 44;                       // [44..46] tiles of round r
constexpr uint32_t TP_CURSOR = 47;                      // [47..49] work cursor of round r
constexpr uint32_t TP_HAND = 50;                        // queries handed to k_sieve_search
constexpr uint32_t TP_DONE = 51;                        // queries answered by the pipeline
constexpr uint32_t TP_SVCUR = 52;                       // k_sieve_search's cursor over the handed-over queries
constexpr uint32_t TP_FAILED = 53;                      // of TP_HAND: a tile failed / out of rounds / not enough two-match terms
constexpr uint32_t TP_WANT0 = 700;                      // candidate blocks the first pass of a query aims at
constexpr uint32_t TP_WANT = 8192;                      // ... a later pass

struct FinalShared {
	uint32_t hist[256];
	unsigned long long* buf;
	unsigned long long tau;
	uint32_t bufCount, selPrefixLo, selK;
	int topShift, kidBits;
};

// ---- one warp per fast query: state, tau, required count, next pass -> tiles
__global__ void __launch_bounds__(256) k_tile_plan(IndexView ix, const QDesc* __restrict__ desc, const uint32_t* __restrict__ fastList,
                                                   uint32_t* __restrict__ counters, TileQ* __restrict__ tq, uint4* __restrict__ tiles,
                                                   uint32_t round, uint32_t k, uint32_t tileG, uint32_t* __restrict__ handList)
{
	const uint32_t wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t lane = threadIdx.x & 31;
	if (wi >= counters[0]) return;
	const uint32_t q = fastList[wi];
	const QDesc& d = desc[q];
	TileQ& s = tq[q];
	const uint32_t NG = ix.sieveNGroups;
	const int shift = ix.sieveShift;
	const uint32_t nGrams = d.nGrams, nLists = d.nLists, cmin = d.cmin, cprom = d.cprom;
	const bool twoPhase = cmin <= 1 && nGrams >= 2;
	uint32_t sumMult = lane < nLists ? d.listMult[lane] : 0u;
#pragma unroll
	for (int o = 16; o; o >>= 1) sumMult += __shfl_xor_sync(0xffffffffu, sumMult, o);
	uint32_t gLo = 0, lastG1 = 0, lastLen = 0, lastCreq = 0, nCandLast = 0, level = 0;
	unsigned long long tau;
	if (round == 0) {
		// class mode is a property of the index (the host launches the pipeline only there); per query: every list has a bitmap,
		// no edit-distance path (flags bit 2), at most SV_MAXL counted lists, and a first phase that is sparse — terms with >= 2
		// matches (no threshold) or >= cmin >= 2 matches
		const bool ok = (d.flags & 4u) && sumMult <= SV_MAXL && (twoPhase || cmin >= 2);
		if (!ok) {
			if (lane == 0) { s.status = 2; handList[atomicAdd(&counters[TP_HAND], 1u)] = q; }
			return;
		}
		s.geHist[lane] = 0;
		if (lane == 0) { s.geHist[32] = 0; s.status = 0; s.candCount = 0; s.failed = 0; }
		tau = twoPhase ? ((unsigned long long)nGrams << 32) : ~0ull;      // "< tau": two or more matches (exactKey[1] = nGrams)
	} else {
		if (s.status != 0) return;
		if (s.failed) {
			if (lane == 0) { s.status = 2; handList[atomicAdd(&counters[TP_HAND], 1u)] = q; atomicAdd(&counters[TP_FAILED], 1u); }
			return;
		}
		gLo = s.gLo; lastG1 = s.lastG1; lastLen = s.lastLen; lastCreq = s.lastCreq; nCandLast = s.nCand; level = s.candCount;
		tau = s.tau;
	}
	auto exactKey = [&](uint32_t c) { return c <= nGrams ? nGrams - c + 1 : 255u; };
	auto hiKey = [&](uint32_t c) { return c >= cprom ? 0u : exactKey(c); };
	if (lastG1) {
		// k candidates with >= c matches are known, all of them below the end of the last pass
		const uint32_t c = lane + 1;
		uint32_t h = c <= nGrams ? s.geHist[c] : 0u;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_down_sync(0xffffffffu, h, o); if (lane + o < 32) h += t; }
		const unsigned m = __ballot_sync(0xffffffffu, h >= k);
		if (m) {
			const uint32_t cb = 32u - (uint32_t)__clz(m);
			const unsigned long long hiId = min(((unsigned long long)lastG1 * SIEVE_GROUP_BLOCKS) << shift, 0xFFFFFFFFull);
			const unsigned long long t2 = (((unsigned long long)exactKey(cb)) << 32) | hiId;
			if (t2 < tau) tau = t2;
		}
	}
	uint32_t rej = cmin ? cmin - 1 : 0;
	{
		const unsigned long long lowKid = min(((unsigned long long)gLo * SIEVE_GROUP_BLOCKS) << shift, 0xFFFFFFFFull);
		const uint32_t c = rej + 1 + lane;
		const bool out = c <= nGrams && ((((unsigned long long)hiKey(min(c, (uint32_t)MAX_QGRAMS))) << 32) | lowKid) >= tau;
		const unsigned mOut = __ballot_sync(0xffffffffu, out);
		rej += mOut == 0xffffffffu ? 32u : (uint32_t)__ffs(~mOut) - 1;
	}
	const uint32_t creq = rej + 1;
	if (creq > sumMult || gLo >= NG) {
		// the whole id range is covered (or nothing further can matter). Without a threshold the pipeline only looked for terms
		// with two or more matches: fewer than k of them -> k_sieve_search starts over and fills the list with single matches
		if (lane == 0) {
			if (twoPhase && level < k) { s.status = 2; handList[atomicAdd(&counters[TP_HAND], 1u)] = q; atomicAdd(&counters[TP_FAILED], 1u); }
			else { s.status = 1; s.tau = tau; }
		}
		return;
	}
	if (round >= (uint32_t)TP_ROUNDS) {
		if (lane == 0) { s.status = 2; handList[atomicAdd(&counters[TP_HAND], 1u)] = q; atomicAdd(&counters[TP_FAILED], 1u); }
		return;
	}
	// ---- the next pass: long enough for the candidate blocks it aims at; the last round takes whatever is left
	const uint32_t rest = NG - gLo;
	uint32_t len = rest;
	if (lane == 0 && round + 1 < (uint32_t)TP_ROUNDS) {
		const uint32_t nBlocks = (uint32_t)((((unsigned long long)ix.T - 1) >> shift) + 1);
		const float dens = fminf(1.0f, (float)d.total / ((float)nLists * (float)nBlocks));
		float rate = fmaxf(3.0f * sieve_rate(nLists, creq, dens), 1e-4f);                     // candidate blocks per 96-block group
		if (creq == lastCreq && lastLen) rate = fmaxf(1.25f * (float)(nCandLast + 1) / (float)lastLen, 1e-4f);
		float l = (float)(round == 0 ? TP_WANT0 : TP_WANT) / rate;
		if (twoPhase && creq <= 2 && lastLen) {
			// the two-match class is still open: `level` terms were found in the groups so far, k are needed
			const float need = (float)(k > level ? k - level : 1u) + 8.0f;
			l = fminf(l, 1.5f * (float)gLo * need / ((float)level + 4.0f));
			l = fmaxf(l, 0.5f * (float)lastLen);
		}
		len = l >= (float)rest ? rest : max(1u, (uint32_t)l);
		if (rest - len < 64) len = rest;                                                       // no tiny remainders
	}
	len = __shfl_sync(0xffffffffu, len, 0);
	const uint32_t g1 = gLo + len;
	uint32_t nT = (len + tileG - 1) / tileG;
	const uint32_t per = (((len + nT - 1) / nT) + 63u) & ~63u;                                 // equal tiles, whole warp steps
	nT = (len + per - 1) / per;
	uint32_t base = 0;
	if (lane == 0) base = atomicAdd(&counters[TP_COUNT + round], nT);
	base = __shfl_sync(0xffffffffu, base, 0);
	for (uint32_t i = lane; i < nT; i += 32) tiles[base + i] = make_uint4(q, gLo + i * per, min(g1, gLo + (i + 1) * per), creq);
	if (lane == 0) {
		s.tau = tau; s.gLo = g1; s.lastG1 = g1; s.lastLen = len; s.lastCreq = creq; s.nCand = 0; s.level0 = level;
	}
}

// ---- k_tile_sieve: every WARP takes tiles on its own — no CTA barrier anywhere, a warp's queue, candidate buffer and query
// tables live in its slice of shared memory.
constexpr int WT_WARPS = 4;                             // warps per CTA
constexpr int WT_BUF = 256;                             // candidates a warp collects before it flushes them to the query's list
constexpr int WT_WQCAP = 160;                           // blocks that passed the sieve: < 32 left over + <= 128 pushed
constexpr int WT_SLOWCAP = 64;                          // blocks waiting for the general resolution: < 32 left over + <= 32 pushed

struct WarpTile {
	const uint4* sievePtr[SV_MAXL];                     // group arrays as the sieve counts them: list j repeated mult[j] times
	const uint4* grpPtr[SV_MAXL];                       // group array of list j
	uint32_t listEnd[SV_MAXL], mult[SV_MAXL];
	uint32_t hiKey[MAX_QGRAMS + 1], exactKey[MAX_QGRAMS + 1], geHist[MAX_QGRAMS + 1];
	uint32_t wq[WT_WQCAP];
	uint32_t slowQ[WT_SLOWCAP];
	unsigned long long buf[WT_BUF];
	unsigned long long tau;
	uint32_t bufCount, nSlow, overflow, nGrams, nLists, cprom, qlen, cRej;
};

// where a warp's candidates go
struct TileSink {
	TileQ* q;
	unsigned long long* cand;
	uint32_t candCap;
};

// a term with exact match count c > cRej. Works under divergence (shared-memory atomics on the warp's own counters).
template <class U>
__device__ __forceinline__ void wt_emit(WarpTile& w, const IndexView& ix, const U* __restrict__ qn, uint32_t term, uint32_t c)
{
	uint32_t hk = w.exactKey[c];
	if (c >= w.cprom) {
		if (key_equals_query<U>(ix, term, qn, w.qlen)) hk = 0u;                              // promoted to 100 (hpp:333-335)
	}
	const unsigned long long key = (((unsigned long long)hk) << 32) | term;
	if (key < w.tau) {
		const uint32_t at = atomicAdd(&w.bufCount, 1u);
		if (at < (uint32_t)WT_BUF) w.buf[at] = key; else w.overflow = 1;
		if (c < w.cprom) atomicAdd(&w.geHist[c], 1u);
	}
}

// the warp's candidates -> the query's list; its counts -> the query's histogram
__device__ __forceinline__ void wt_flush(WarpTile& w, const TileSink& sink, uint32_t lane)
{
	__syncwarp();
	const uint32_t n = min(w.bufCount, (uint32_t)WT_BUF);
	uint32_t base = 0;
	if (lane == 0 && n) base = atomicAdd(&sink.q->candCount, n);
	base = __shfl_sync(0xffffffffu, base, 0);
	if (base + n > sink.candCap || w.overflow) { if (lane == 0) sink.q->failed = 1; }
	else for (uint32_t i = lane; i < n; i += 32) sink.cand[base + i] = w.buf[i];
	for (uint32_t c = lane; c <= (uint32_t)MAX_QGRAMS; c += 32)
		if (w.geHist[c]) { atomicAdd(&sink.q->geHist[c], w.geHist[c]); w.geHist[c] = 0; }
	__syncwarp();
	if (lane == 0) w.bufCount = 0;
	__syncwarp();
}

// general resolution of one block per lane (more than four lists hit it, or a list has several postings in or just below
// it): local arrays, counts by comparing all pairs
template <class U>
__device__ __noinline__ void wt_resolve_slow(WarpTile& w, const IndexView& ix, const U* __restrict__ qn, const uint32_t blk)
{
	if (blk == INF) return;
	constexpr int CAP = 48;
	const int shift = ix.sieveShift;
	const uint32_t* __restrict__ postings = ix.postings;
	const uint32_t v = blk >> 5, g = v / 3u, k = v - 3u * g, bit = blk & 31u;
	const uint32_t nL = w.nLists, cRej = w.cRej;
	uint32_t lm = 0;
#pragma unroll 4
	for (uint32_t j = 0; j < nL; ++j) lm |= ((__ldg(reinterpret_cast<const uint32_t*>(w.grpPtr[j] + g) + 1 + k) >> bit) & 1u) << j;
	uint32_t ids[CAP], mus[CAP], m = 0;
	while (lm) {
		const uint32_t j = (uint32_t)__ffs(lm) - 1u;
		lm &= lm - 1;
		uint32_t p;
		sieve_locate(__ldg(w.grpPtr[j] + g), k, bit, p);
		const uint32_t end = w.listEnd[j], mu = w.mult[j];
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
				if (m < CAP) { ids[m] = id[t]; mus[m] = mu; ++m; } else w.overflow = 1;
				any = true;
			}
			p += 4;
		}
		if (!any) w.overflow = 1;                            // bitmap and postings disagree: never trust, hand the query over
	}
	for (uint32_t a = 0; a < m; ++a) {
		uint32_t c = 0;
		bool first = true;
		for (uint32_t b = 0; b < m; ++b)
			if (ids[b] == ids[a]) { if (b < a) first = false; c += mus[b]; }
		if (first && c > cRej) wt_emit<U>(w, ix, qn, ids[a], c);
	}
}

// ---- RESOLVE one block per lane (blk == INF: lane idle). Two waits: (1) the group of every list (streamed by this warp
// moments ago: L1 / L2 hits) -> which lists hit the block and where the block's first posting is — the group's first posting
// + the occupied blocks below (exact unless one of those holds several); (2) two postings per list: the first one must be the
// block's only posting of the list, everything else goes to the general resolution.
template <class U, int NLMAX>
__device__ __noinline__ void wt_resolve_block(WarpTile& w, const IndexView& ix, const U* __restrict__ qn, const uint32_t blk, const uint32_t lane)
{
	const int shift = ix.sieveShift;
	const uint32_t* __restrict__ postings = ix.postings;
	const bool act = blk != INF;
	const uint32_t bq = act ? blk : 0u;
	const uint32_t v = bq >> 5, g = v / 3u, k = v - 3u * g, bit = bq & 31u;
	const uint32_t nL = w.nLists;
	const uint32_t low = (1u << bit) - 1u;
	const uint32_t my = k == 0 ? low : 0xFFFFFFFFu, mz = k == 1 ? low : (k > 1 ? 0xFFFFFFFFu : 0u), mw = k == 2 ? low : 0u;
	uint32_t idx[SV_PU] = {0, 0, 0, 0}, jp = 0, ns = 0;
	{
		uint4 x[NLMAX];
#pragma unroll
		for (int j = 0; j < NLMAX; ++j) x[j] = (uint32_t)j < nL ? __ldg(w.grpPtr[j] + g) : make_uint4(0, 0, 0, 0);
#pragma unroll
		for (int j = NLMAX - 1; j >= 0; --j) {
			const uint32_t wk = k == 0 ? x[j].y : (k == 1 ? x[j].z : x[j].w);
			const bool hit = ((wk >> bit) & 1u) != 0;
			const uint32_t p = x[j].x + __popc(x[j].y & my) + __popc(x[j].z & mz) + __popc(x[j].w & mw);
			if (hit) { idx[3] = idx[2]; idx[2] = idx[1]; idx[1] = idx[0]; idx[0] = p; jp = (jp << 4) | (uint32_t)j; ++ns; }
		}
	}
	if (!act) ns = 0;
	bool slow = ns > (uint32_t)SV_PU;
	uint32_t h[SV_PU], mu[SV_PU];
#pragma unroll
	for (int u = 0; u < SV_PU; ++u) {
		const uint32_t j = (jp >> (4 * u)) & 15u;
		const bool on = (uint32_t)u < ns;
		const uint32_t end = on ? w.listEnd[j] : 0u;
		mu[u] = on ? w.mult[j] : 0u;
		const uint32_t id0 = idx[u] < end ? __ldg(postings + idx[u]) : INF;
		const uint32_t id1 = idx[u] + 1 < end ? __ldg(postings + idx[u] + 1) : INF;
		h[u] = 0xFFFFFFF0u + (uint32_t)u;                    // no posting: a value of its own
		if (on) {
			if (id0 != INF && (id0 >> shift) == blk && (id1 == INF || (id1 >> shift) != blk)) h[u] = id0;
			else slow = true;
		}
	}
	// blocks for the general resolution are compacted and resolved a full warp at a time
	{
		const unsigned ms = __ballot_sync(0xffffffffu, slow);
		if (ms) {
			const uint32_t base = w.nSlow;
			if (slow) w.slowQ[base + __popc(ms & ((1u << lane) - 1u))] = blk;
			__syncwarp();
			if (lane == 0) w.nSlow = base + __popc(ms);
			__syncwarp();
		}
	}
	if (slow || ns == 0) return;
	const uint32_t cRej = w.cRej;
#pragma unroll
	for (int u = 0; u < SV_PU; ++u) {
		uint32_t c = mu[u];
		bool first = true;
#pragma unroll
		for (int x2 = 0; x2 < SV_PU; ++x2)
			if (x2 != u && h[x2] == h[u]) { c += mu[x2]; if (x2 < u) first = false; }
		if (first && c > cRej) wt_emit<U>(w, ix, qn, h[u], c);
	}
}

template <class U>
__device__ __forceinline__ void wt_resolve_round(WarpTile& w, const IndexView& ix, const U* __restrict__ qn, const uint32_t blk, const uint32_t lane,
                                                 const TileSink& sink)
{
	const uint32_t nL = w.nLists;
	if (nL <= 8) wt_resolve_block<U, 8>(w, ix, qn, blk, lane);
	else if (nL <= 12) wt_resolve_block<U, 12>(w, ix, qn, blk, lane);
	else wt_resolve_block<U, 16>(w, ix, qn, blk, lane);
	__syncwarp();
	if (w.nSlow >= 32) {
		const uint32_t n = w.nSlow - 32;
		const uint32_t sb = w.slowQ[n + lane];
		__syncwarp();
		if (lane == 0) w.nSlow = n;
		wt_resolve_slow<U>(w, ix, qn, sb);
		__syncwarp();
	}
	if (w.bufCount > (uint32_t)WT_BUF / 4) wt_flush(w, sink, lane);     // keeps room for the next rounds (<= 4 emits per block)
}

// ---- SIEVE + RESOLVE over the groups [g0, g1) of one tile by one warp (see sieve_groups in sieve_path.cuh)
template <class U, int CREQ>
__device__ __noinline__ uint32_t wt_sieve(WarpTile& w, const IndexView& ix, const U* __restrict__ qn, const uint32_t g0, const uint32_t g1,
                                         const uint32_t nS, const uint32_t lane, const TileSink& sink)
{
	uint32_t* wq = w.wq;
	uint32_t qn_ = 0, nc = 0;                                    // queue fill (warp-uniform); candidate blocks
	for (uint32_t base = g0; base < g1; base += 64) {
		const uint32_t gA = base + lane, gB = gA + 32;
		uint32_t m[6];
		{
			uint32_t ga[CREQ], gb[CREQ], gc[CREQ], ha[CREQ], hb[CREQ], hc[CREQ];
#pragma unroll
			for (int c = 0; c < CREQ; ++c) { ga[c] = 0; gb[c] = 0; gc[c] = 0; ha[c] = 0; hb[c] = 0; hc[c] = 0; }
			const bool onA = gA < g1, onB = gB < g1;
#pragma unroll 4
			for (uint32_t j = 0; j < nS; ++j) {
				const uint4* ptr = w.sievePtr[j];
				const uint4 x = onA ? __ldg(ptr + gA) : make_uint4(0, 0, 0, 0);
				const uint4 y = onB ? __ldg(ptr + gB) : make_uint4(0, 0, 0, 0);
#pragma unroll
				for (int c = CREQ - 1; c >= 1; --c) {
					ga[c] |= ga[c - 1] & x.y; gb[c] |= gb[c - 1] & x.z; gc[c] |= gc[c - 1] & x.w;
					ha[c] |= ha[c - 1] & y.y; hb[c] |= hb[c - 1] & y.z; hc[c] |= hc[c - 1] & y.w;
				}
				ga[0] |= x.y; gb[0] |= x.z; gc[0] |= x.w; ha[0] |= y.y; hb[0] |= y.z; hc[0] |= y.w;
			}
			m[0] = ga[CREQ - 1]; m[1] = gb[CREQ - 1]; m[2] = gc[CREQ - 1]; m[3] = ha[CREQ - 1]; m[4] = hb[CREQ - 1]; m[5] = hc[CREQ - 1];
		}
		uint32_t left = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]) + __popc(m[4]) + __popc(m[5]);
#pragma unroll 1
		while (__any_sync(0xffffffffu, left != 0)) {
			uint32_t inc = left;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += t; }
			uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
			uint32_t cnt = left;
			if (qn_ + tot > (uint32_t)WT_WQCAP) {
				cnt = min(left, 4u);
				inc = cnt;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += t; }
				tot = __shfl_sync(0xffffffffu, inc, 31);
			}
			uint32_t at = qn_ + inc - cnt;
			left -= cnt;
			nc += cnt;
#pragma unroll
			for (int wd = 0; wd < 6; ++wd) {
				const uint32_t b0 = (wd < 3 ? gA : gB) * SIEVE_GROUP_BLOCKS + (wd % 3) * 32;
#pragma unroll 1
				while (m[wd] && cnt) { wq[at++] = b0 + (__ffs(m[wd]) - 1); m[wd] &= m[wd] - 1; --cnt; }
			}
			qn_ += tot;
			__syncwarp();
#pragma unroll 1
			while (qn_ >= 32) {
				qn_ -= 32;
				const uint32_t blk = wq[qn_ + lane];
				__syncwarp();
				wt_resolve_round<U>(w, ix, qn, blk, lane, sink);
			}
		}
	}
	if (qn_) {
		const uint32_t blk = lane < qn_ ? wq[lane] : INF;
		__syncwarp();
		wt_resolve_round<U>(w, ix, qn, blk, lane, sink);
	}
	if (w.nSlow) {
		const uint32_t sb = lane < w.nSlow ? w.slowQ[lane] : INF;
		__syncwarp();
		if (lane == 0) w.nSlow = 0;
		wt_resolve_slow<U>(w, ix, qn, sb);
		__syncwarp();
	}
#pragma unroll
	for (int o = 16; o; o >>= 1) nc += __shfl_xor_sync(0xffffffffu, nc, o);
	return nc;
}

template <class U>
__global__ void __launch_bounds__(WT_WARPS * 32, SV_MIN_CTAS) k_tile_sieve(IndexView ix, const QDesc* __restrict__ desc, const U* __restrict__ qnorm,
                                                                          TileQ* __restrict__ tq, const uint4* __restrict__ tiles,
                                                                          uint32_t* __restrict__ counters, uint32_t round,
                                                                          unsigned long long* __restrict__ cand, uint32_t candCap)
{
	__shared__ WarpTile wts[WT_WARPS];
	WarpTile& w = wts[threadIdx.x >> 5];
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t nTiles = counters[TP_COUNT + round];
	const uint32_t NG = ix.sieveNGroups;
	uint32_t nTilesDone = 0, nCandAll = 0;
	while (true) {
		uint32_t ti = 0;
		if (lane == 0) ti = atomicAdd(&counters[TP_CURSOR + round], 1u);
		ti = __shfl_sync(0xffffffffu, ti, 0);
		if (ti >= nTiles) break;
		const uint4 t = tiles[ti];
		const uint32_t q = t.x, g0 = t.y, g1 = t.z, creq = t.w;
		const QDesc& d = desc[q];
		const U* qn = qnorm + d.qbeg;
		const uint32_t nLists = d.nLists, nGrams = d.nGrams, cprom = d.cprom;
		// the query's tables: lane j = list j
		uint32_t mu = 0;
		const uint4* gp = ix.sieveGroups;
		if (lane < nLists) { mu = d.listMult[lane]; gp = ix.sieveGroups + (size_t)d.listSlot[lane] * NG; }
		uint32_t inc = mu;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += x; }
		const uint32_t nS = __shfl_sync(0xffffffffu, inc, 31);
		if (lane < (uint32_t)SV_MAXL) {
			w.grpPtr[lane] = gp;
			w.listEnd[lane] = lane < nLists ? d.listOff[lane] + d.listLen[lane] : 0u;
			w.mult[lane] = mu;
		}
		for (uint32_t r = 0; r < mu; ++r) { const uint32_t at = inc - mu + r; if (at < (uint32_t)SV_MAXL) w.sievePtr[at] = gp; }
		for (uint32_t c = lane; c <= (uint32_t)MAX_QGRAMS; c += 32) {
			const uint32_t ek = c <= nGrams ? nGrams - c + 1 : 255u;
			w.exactKey[c] = ek;
			w.hiKey[c] = c >= cprom ? 0u : ek;
			w.geHist[c] = 0;
		}
		if (lane == 0) {
			w.tau = tq[q].tau;
			w.bufCount = 0; w.nSlow = 0; w.overflow = 0;
			w.nGrams = nGrams; w.nLists = nLists; w.cprom = cprom; w.qlen = d.qlen; w.cRej = creq - 1;
		}
		__syncwarp();
		TileSink sink;
		sink.q = tq + q; sink.cand = cand + (size_t)q * candCap; sink.candCap = candCap;
		uint32_t nc;
		// thresholds without an instantiation of their own use the next lower one: a superset of the blocks, RESOLVE counts exactly
		if (creq <= 1) nc = wt_sieve<U, 1>(w, ix, qn, g0, g1, nS, lane, sink);
		else if (creq == 2) nc = wt_sieve<U, 2>(w, ix, qn, g0, g1, nS, lane, sink);
		else if (creq == 3) nc = wt_sieve<U, 3>(w, ix, qn, g0, g1, nS, lane, sink);
		else if (creq < 6) nc = wt_sieve<U, 4>(w, ix, qn, g0, g1, nS, lane, sink);
		else if (creq < 8) nc = wt_sieve<U, 6>(w, ix, qn, g0, g1, nS, lane, sink);
		else nc = wt_sieve<U, 8>(w, ix, qn, g0, g1, nS, lane, sink);
		wt_flush(w, sink, lane);
		if (lane == 0) {
			if (w.overflow) tq[q].failed = 1;
			atomicAdd(&tq[q].nCand, nc);
		}
		++nTilesDone; nCandAll += nc;
		__syncwarp();
	}
	if (lane == 0 && nTilesDone) { atomicAdd(&counters[24], nTilesDone); atomicAdd(&counters[26], nCandAll); }
}

// ---- one CTA per fast query that the pipeline completed: top-k of the candidate list, sorted, written out
__global__ void __launch_bounds__(128) k_tile_final(IndexView ix, const QDesc* __restrict__ desc, const uint32_t* __restrict__ fastList,
                                                    uint32_t* __restrict__ counters, const TileQ* __restrict__ tq,
                                                    const unsigned long long* __restrict__ cand, uint32_t candCap, uint32_t limit,
                                                    uint32_t outStride, uint32_t* __restrict__ outCount, uint32_t* __restrict__ outKid,
                                                    float* __restrict__ outScore, unsigned long long* __restrict__ statCounters, HostOut ho)
{
	extern __shared__ __align__(16) unsigned char tf_smem[];
	FinalShared& sh = *reinterpret_cast<FinalShared*>(tf_smem);
	const int tid = threadIdx.x;
	if (blockIdx.x >= counters[0]) return;
	const uint32_t q = fastList[blockIdx.x];
	const TileQ& s = tq[q];
	if (s.status != 1) return;
	const QDesc& d = desc[q];
	const uint32_t nGrams = d.nGrams, k = limit;
	if (tid == 0) {
		sh.buf = reinterpret_cast<unsigned long long*>(tf_smem + ((sizeof(FinalShared) + 15) & ~size_t(15)));
		sh.topShift = 32;
		int kb = 1;
		while (kb < 32 && (1ull << kb) < ix.K) ++kb;
		sh.kidBits = kb;
		sh.tau = ~0ull;
	}
	uint32_t level = min(s.candCount, candCap);
	__syncthreads();
	const unsigned long long* src = cand + (size_t)q * candCap;
	for (uint32_t i = tid; i < level; i += 128) sh.buf[i] = src[i];
	__syncthreads();
	if (level > k) { select_smallest<128>(sh, level, k); level = k; }
	sort_buffer<128>(sh, level);
	for (uint32_t i = tid; i < level; i += 128) {
		const unsigned long long v = sh.buf[i];
		const uint32_t kid = (uint32_t)v;
		const uint32_t hi = (uint32_t)(v >> 32);
		const float score = hi == 0 ? 100.0f : edge_value(1.0f, term_score(nGrams - hi + 1, nGrams));
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
		atomicAdd(&counters[TP_DONE], 1u);
		if (level != outStride) atomicAdd(&counters[22], 1u);
	}
}

}   // namespace
