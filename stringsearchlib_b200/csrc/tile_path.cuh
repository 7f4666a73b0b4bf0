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
//   k_tile_stream  persistent warps take tiles from the list: the SIEVE over the tile's groups as in k_sieve_search, and for
//                 every block that passes one record (which lists hit it, where their postings are) in a queue in HBM
//   k_tile_resolve persistent warps take records: the postings, exact counts, candidates appended to the query's list in HBM
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
constexpr uint32_t TP_TILE_CANDS = 128;                 // candidate blocks one tile aims at

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
	uint32_t len = rest, tileLen = tileG;
	if (lane == 0) {
		const uint32_t nBlocks = (uint32_t)((((unsigned long long)ix.T - 1) >> shift) + 1);
		const float dens = fminf(1.0f, (float)d.total / ((float)nLists * (float)nBlocks));
		float rate = fmaxf(3.0f * sieve_rate(nLists, creq, dens), 1e-4f);                     // candidate blocks per 96-block group
		if (creq == lastCreq && lastLen) rate = fmaxf(1.25f * (float)(nCandLast + 1) / (float)lastLen, 1e-4f);
		if (round + 1 < (uint32_t)TP_ROUNDS) {
			float l = (float)(round == 0 ? TP_WANT0 : TP_WANT) / rate;
			if (twoPhase && creq <= 2 && lastLen) {
				// the two-match class is still open: `level` terms were found in the groups so far, k are needed
				const float need = (float)(k > level ? k - level : 1u) + 8.0f;
				l = fminf(l, 1.5f * (float)gLo * need / ((float)level + 4.0f));
				l = fmaxf(l, 0.5f * (float)lastLen);
			}
			len = l >= (float)rest ? rest : max(1u, (uint32_t)l);
			if (rest - len < 64) len = rest;                                                   // no tiny remainders
		}
		// a tile = one warp's work: about TP_TILE_CANDS candidate blocks, between one bitmap step (64 groups) and tileG groups
		const float tl = (float)TP_TILE_CANDS / rate;
		tileLen = tl >= (float)tileG ? tileG : max(64u, ((uint32_t)tl + 63u) & ~63u);
	}
	len = __shfl_sync(0xffffffffu, len, 0);
	tileLen = __shfl_sync(0xffffffffu, tileLen, 0);
	const uint32_t g1 = gLo + len;
	uint32_t nT = (len + tileLen - 1) / tileLen;
	const uint32_t per = (((len + nT - 1) / nT) + 63u) & ~63u;                                 // equal tiles, whole bitmap steps
	nT = (len + per - 1) / per;
	uint32_t base = 0;
	if (lane == 0) base = atomicAdd(&counters[TP_COUNT + round], nT);
	base = __shfl_sync(0xffffffffu, base, 0);
	for (uint32_t i = lane; i < nT; i += 32) tiles[base + i] = make_uint4(q, gLo + i * per, min(g1, gLo + (i + 1) * per), creq);
	if (lane == 0) {
		s.tau = tau; s.gLo = g1; s.lastG1 = g1; s.lastLen = len; s.lastCreq = creq; s.nCand = 0; s.level0 = level;
	}
}

// ---- k_tile_stream + k_tile_resolve: the dependent chain of a candidate block (which lists hit it -> where their postings
// are -> the postings themselves -> counts) is cut where it leaves the cache. k_tile_stream streams the bitmaps and, for the
// blocks that pass, re-reads what it has just streamed (L1 / L2 hits) to find the lists that hit the block and the position of
// the block's first posting in each; it writes one 32-byte RECORD per block. k_tile_resolve — every warp of the GPU at once,
// nothing else to wait for — gathers the postings the records point at (the DRAM misses), counts the terms exactly as the
// reference does (hpp:289-298, with multiplicities) and appends what can still matter to the query's candidate list.
constexpr int TS_WARPS = 4;                             // warps per CTA (independent of each other: no CTA barrier)
#ifndef TS_TIMING
#define TS_TIMING 0
#endif
#ifndef TS_STREAM_CTAS
#define TS_STREAM_CTAS 10
#endif
constexpr int TS_WQCAP = 160;                           // (query, block) pairs that passed the sieve: < 32 left over + <= 128 pushed
constexpr uint32_t TS_STRIPES = 64;                     // record queues per round (a warp appends to queue warp-id % 64)
constexpr uint32_t TP_RCUR = 56;                        // [56..58] k_tile_resolve's work cursor of round r
constexpr uint32_t TP_STRIPE0 = 64;                     // [64 + 64 r + s] records in queue s of round r
constexpr uint32_t TP_COUNTERS = TP_STRIPE0 + TS_STRIPES * TP_ROUNDS;

// record of one candidate block: a = { query, block, list numbers (4 bits each) | lists << 16, 0 }, b = position of the
// block's first posting in each of up to four lists. lists == 15: more than four lists hit the block (general resolution).
struct TileRec { uint4 a, b; };

// one bitmap step of a warp: lane l streams groups base + l and base + 32 + l of every counted list through the bit-sliced
// threshold counter (sieve_groups in sieve_path.cuh); m = the blocks that at least CREQ lists hit
template <int CREQ>
__device__ __forceinline__ void ts_step(const uint4* __restrict__ groups, const uint32_t NG, const uint32_t mySlot, const uint32_t nS,
                                        const uint32_t gA, const uint32_t gB, const uint32_t g1, uint32_t (&m)[6])
{
	uint32_t ga[CREQ], gb[CREQ], gc[CREQ], ha[CREQ], hb[CREQ], hc[CREQ];
#pragma unroll
	for (int c = 0; c < CREQ; ++c) { ga[c] = 0; gb[c] = 0; gc[c] = 0; ha[c] = 0; hb[c] = 0; hc[c] = 0; }
	const bool onA = gA < g1, onB = gB < g1;
#pragma unroll 4
	for (uint32_t j = 0; j < nS; ++j) {
		const uint4* ptr = groups + (size_t)__shfl_sync(0xffffffffu, mySlot, j) * NG;
		const uint4 x = onA ? __ldg(ptr + gA) : make_uint4(0, 0, 0, 0);      // { first posting, 3 x 32 block bits }
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

// the same for one group per lane (high thresholds: half the accumulators)
template <int CREQ>
__device__ __forceinline__ void ts_step1(const uint4* __restrict__ groups, const uint32_t NG, const uint32_t mySlot, const uint32_t nS,
                                         const uint32_t gA, const uint32_t g1, uint32_t& m0, uint32_t& m1, uint32_t& m2)
{
	uint32_t ga[CREQ], gb[CREQ], gc[CREQ];
#pragma unroll
	for (int c = 0; c < CREQ; ++c) { ga[c] = 0; gb[c] = 0; gc[c] = 0; }
	const bool onA = gA < g1;
#pragma unroll 4
	for (uint32_t j = 0; j < nS; ++j) {
		const uint4* ptr = groups + (size_t)__shfl_sync(0xffffffffu, mySlot, j) * NG;
		const uint4 x = onA ? __ldg(ptr + gA) : make_uint4(0, 0, 0, 0);
#pragma unroll
		for (int c = CREQ - 1; c >= 1; --c) { ga[c] |= ga[c - 1] & x.y; gb[c] |= gb[c - 1] & x.z; gc[c] |= gc[c - 1] & x.w; }
		ga[0] |= x.y; gb[0] |= x.z; gc[0] |= x.w;
	}
	m0 = ga[CREQ - 1]; m1 = gb[CREQ - 1]; m2 = gc[CREQ - 1];
}

// one (query, block) pair per lane (e == ~0: lane idle) -> one record per lane, 32 records appended to the warp's queue
// curQ / curNL / curSlot: the tile in flight (lane j holds the bitmap slot of its list j) — entries left over from earlier
// tiles read their query's descriptor instead
__device__ __noinline__ void ts_round(const IndexView& ix, const QDesc* __restrict__ desc, TileQ* __restrict__ tq, const unsigned long long e,
                                      const uint32_t lane, TileRec* __restrict__ recs, const uint32_t recCap, uint32_t* __restrict__ stripeCount,
                                      const uint32_t curQ, const uint32_t curNL, const uint32_t curSlot)
{
	const bool act = e != ~0ull;
	const uint32_t q = act ? (uint32_t)(e >> 32) : curQ, blk = act ? (uint32_t)e : 0u;
	const QDesc& d = desc[q];
	const uint32_t NG = ix.sieveNGroups;
	const uint32_t v = blk >> 5, g = v / 3u, k = v - 3u * g, bit = blk & 31u;
	const bool cur = q == curQ;
	const uint32_t nL = act ? (cur ? curNL : d.nLists) : 0u;
	// wait 1: bitmap word of every list (streamed moments ago) -> which lists hit the block
	uint32_t lm = 0;
	{
		uint32_t wv[SV_MAXL];
#pragma unroll
		for (int j = 0; j < SV_MAXL; ++j) {
			uint32_t sl = __shfl_sync(0xffffffffu, curSlot, j);
			if (!cur && (uint32_t)j < nL) sl = d.listSlot[j];
			wv[j] = (uint32_t)j < nL ? __ldg(reinterpret_cast<const uint32_t*>(ix.sieveGroups + (size_t)sl * NG + g) + 1 + k) : 0u;
		}
#pragma unroll
		for (int j = 0; j < SV_MAXL; ++j) lm |= ((wv[j] >> bit) & 1u) << j;
	}
	const uint32_t ns = __popc(lm);
	const uint32_t low = (1u << bit) - 1u;
	const uint32_t my = k == 0 ? low : 0xFFFFFFFFu, mz = k == 1 ? low : (k > 1 ? 0xFFFFFFFFu : 0u), mw = k == 2 ? low : 0u;
	// wait 2: the whole group of the first four of them -> the group's first posting + the occupied blocks below
	uint32_t idx[SV_PU], jp = 0;
	{
		uint4 x[SV_PU];
		uint32_t rest = lm;
#pragma unroll
		for (int u = 0; u < SV_PU; ++u) {
			const uint32_t j = rest ? (uint32_t)__ffs(rest) - 1u : 0u;
			uint32_t sl = __shfl_sync(0xffffffffu, curSlot, j);
			if (!cur && rest) sl = d.listSlot[j];
			x[u] = rest ? __ldg(ix.sieveGroups + (size_t)sl * NG + g) : make_uint4(0, 0, 0, 0);
			jp |= j << (4 * u);
			rest &= rest - 1;
		}
#pragma unroll
		for (int u = 0; u < SV_PU; ++u) idx[u] = x[u].x + __popc(x[u].y & my) + __popc(x[u].z & mz) + __popc(x[u].w & mw);
	}
	uint32_t base = 0;
	const uint32_t stripe = ((blockIdx.x * TS_WARPS) + (threadIdx.x >> 5)) & (TS_STRIPES - 1);
	if (lane == 0) base = atomicAdd(&stripeCount[stripe], 32u);
	base = __shfl_sync(0xffffffffu, base, 0);
	if (base + 32 > recCap) {                                // the queue is full: these queries go to k_sieve_search
		if (act) tq[q].failed = 1;
		return;
	}
	TileRec* r = recs + (size_t)stripe * recCap + base + lane;
	r->a = make_uint4(q, act ? blk : INF, jp | ((ns > (uint32_t)SV_PU ? 15u : ns) << 16), 0u);
	r->b = make_uint4(idx[0], idx[1], idx[2], idx[3]);
}

__global__ void __launch_bounds__(TS_WARPS * 32, TS_STREAM_CTAS) k_tile_stream(IndexView ix, const QDesc* __restrict__ desc, TileQ* __restrict__ tq,
                                                                   const uint4* __restrict__ tiles, uint32_t* __restrict__ counters, uint32_t round,
                                                                   TileRec* __restrict__ recs, uint32_t recCap)
{
	__shared__ unsigned long long wqs[TS_WARPS][TS_WQCAP];
	unsigned long long* wq = wqs[threadIdx.x >> 5];
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t nTiles = counters[TP_COUNT + round];
	const uint32_t NG = ix.sieveNGroups;
	uint32_t* stripeCount = counters + TP_STRIPE0 + TS_STRIPES * round;
	uint32_t qn = 0;                                          // queue fill (warp-uniform); entries survive from tile to tile
	uint32_t nTilesDone = 0, nCandAll = 0;
	// -DTS_TIMING=1: lane 0 of every warp accumulates the cycles (>> 6) it spends per phase into counters[32..35]
	// (0 tile fetch + set-up, 1 bitmap steps, 2 queue pushes, 3 record rounds)
#if TS_TIMING
	long long tick = clock64();
	uint32_t tacc[4] = {0, 0, 0, 0};
#define TS_TICK(slot) do { const long long t_ = clock64(); tacc[slot] += (uint32_t)((t_ - tick) >> 6); tick = t_; } while (0)
#else
#define TS_TICK(slot) do { } while (0)
#endif
	while (true) {
		uint32_t ti = 0;
		if (lane == 0) ti = atomicAdd(&counters[TP_CURSOR + round], 1u);
		ti = __shfl_sync(0xffffffffu, ti, 0);
		if (ti >= nTiles) break;
		const uint4 t = tiles[ti];
		const uint32_t q = t.x, g0 = t.y, g1 = t.z, creq = t.w;
		const QDesc& d = desc[q];
		const uint32_t nLists = d.nLists;
		// counted lists: list j repeated mult[j] times (hpp:289-298 counts once per window); lane i holds the slot of counted list i
		uint32_t mu = 0, slot = 0;
		if (lane < nLists) { mu = d.listMult[lane]; slot = d.listSlot[lane]; }
		uint32_t inc = mu;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += x; }
		const uint32_t nS = __shfl_sync(0xffffffffu, inc, 31);
		uint32_t mySlot = slot;
		if (nS != nLists) {
			mySlot = 0;
			for (uint32_t j = 0; j < nLists; ++j) {
				const uint32_t hi = __shfl_sync(0xffffffffu, inc, j), m_ = __shfl_sync(0xffffffffu, mu, j), sl = __shfl_sync(0xffffffffu, slot, j);
				if (lane >= hi - m_ && lane < hi) mySlot = sl;
			}
		}
		uint32_t nc = 0;
		TS_TICK(0);
		for (uint32_t base = g0; base < g1; base += 64) {
			const uint32_t gA = base + lane, gB = gA + 32;
			uint32_t m[6];
			// thresholds without an instantiation of their own use the next lower one: a superset of the blocks, k_tile_resolve counts exactly
			if (creq <= 1) ts_step<1>(ix.sieveGroups, NG, mySlot, nS, gA, gB, g1, m);
			else if (creq == 2) ts_step<2>(ix.sieveGroups, NG, mySlot, nS, gA, gB, g1, m);
			else if (creq == 3) ts_step<3>(ix.sieveGroups, NG, mySlot, nS, gA, gB, g1, m);
			else if (creq < 6) { ts_step1<4>(ix.sieveGroups, NG, mySlot, nS, gA, g1, m[0], m[1], m[2]); ts_step1<4>(ix.sieveGroups, NG, mySlot, nS, gB, g1, m[3], m[4], m[5]); }
			else if (creq < 8) { ts_step1<6>(ix.sieveGroups, NG, mySlot, nS, gA, g1, m[0], m[1], m[2]); ts_step1<6>(ix.sieveGroups, NG, mySlot, nS, gB, g1, m[3], m[4], m[5]); }
			else { ts_step1<8>(ix.sieveGroups, NG, mySlot, nS, gA, g1, m[0], m[1], m[2]); ts_step1<8>(ix.sieveGroups, NG, mySlot, nS, gB, g1, m[3], m[4], m[5]); }
			uint32_t left = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]) + __popc(m[4]) + __popc(m[5]);
			TS_TICK(1);
			// rounds: every lane pushes up to `quota` of its blocks, then the queue is drained down to less than a warp's worth
#pragma unroll 1
			while (__any_sync(0xffffffffu, left != 0)) {
				uint32_t incl = left;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += x; }
				uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
				uint32_t cnt = left;
				if (qn + tot > (uint32_t)TS_WQCAP) {
					cnt = min(left, 4u);
					incl = cnt;
#pragma unroll
					for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += x; }
					tot = __shfl_sync(0xffffffffu, incl, 31);
				}
				uint32_t at = qn + incl - cnt;
				left -= cnt;
				nc += cnt;
				const unsigned long long qhi = (unsigned long long)q << 32;
#pragma unroll
				for (int wd = 0; wd < 6; ++wd) {
					const uint32_t b0 = (wd < 3 ? gA : gB) * SIEVE_GROUP_BLOCKS + (wd % 3) * 32;
#pragma unroll 1
					while (m[wd] && cnt) { wq[at++] = qhi | (b0 + (__ffs(m[wd]) - 1)); m[wd] &= m[wd] - 1; --cnt; }
				}
				qn += tot;
				__syncwarp();
				TS_TICK(2);
#pragma unroll 1
				while (qn >= 32) {
					qn -= 32;
					const unsigned long long e = wq[qn + lane];
					__syncwarp();
					ts_round(ix, desc, tq, e, lane, recs, recCap, stripeCount, q, nLists, slot);
				}
				TS_TICK(3);
			}
		}
#pragma unroll
		for (int o = 16; o; o >>= 1) nc += __shfl_xor_sync(0xffffffffu, nc, o);
		if (lane == 0) atomicAdd(&tq[q].nCand, nc);
		++nTilesDone; nCandAll += nc;
	}
	if (qn) {
		const unsigned long long e = lane < qn ? wq[lane] : ~0ull;
		__syncwarp();
		ts_round(ix, desc, tq, e, lane, recs, recCap, stripeCount, 0xFFFFFFFFu, 0u, 0u);
	}
	if (lane == 0 && nTilesDone) { atomicAdd(&counters[24], nTilesDone); atomicAdd(&counters[26], nCandAll); }
#if TS_TIMING
	TS_TICK(3);
	if (lane == 0) { atomicAdd(&counters[32], tacc[0]); atomicAdd(&counters[33], tacc[1]); atomicAdd(&counters[34], tacc[2]); atomicAdd(&counters[35], tacc[3]); }
#endif
}

// rank key of a term of query d with exact match count c (class mode: the high word is the rank of the score class)
template <class U>
__device__ __forceinline__ unsigned long long tr_key(const IndexView& ix, const QDesc& d, const U* __restrict__ qnorm, const uint32_t term, const uint32_t c)
{
	const uint32_t nGrams = d.nGrams;
	uint32_t hk = c <= nGrams ? nGrams - c + 1 : 255u;
	if (c >= d.cprom && key_equals_query<U>(ix, term, qnorm + d.qbeg, d.qlen)) hk = 0u;          // promoted to 100 (hpp:333-335)
	return (((unsigned long long)hk) << 32) | term;
}

// a term of query q with exact match count c > cRej -> the query's candidate list (works under divergence: one atomic per term)
template <class U>
__device__ __forceinline__ void tr_emit(const IndexView& ix, const QDesc& d, const U* __restrict__ qnorm, TileQ& s, unsigned long long* __restrict__ cand,
                                        const uint32_t candCap, const uint32_t q, const uint32_t term, const uint32_t c)
{
	const unsigned long long key = tr_key<U>(ix, d, qnorm, term, c);
	if (key < s.tau) {
		const uint32_t at = atomicAdd(&s.candCount, 1u);
		if (at < candCap) cand[(size_t)q * candCap + at] = key; else s.failed = 1;
		if (c < d.cprom) atomicAdd(&s.geHist[c], 1u);
	}
}

// the same for a whole warp (every lane calls): lanes with `has` append their key, one atomic per distinct query
__device__ __forceinline__ void tr_append(TileQ* __restrict__ tq, unsigned long long* __restrict__ cand, const uint32_t candCap, const bool has,
                                          const uint32_t q, const unsigned long long key, const uint32_t lane)
{
	unsigned rem = __ballot_sync(0xffffffffu, has);
	while (rem) {
		const uint32_t leader = (uint32_t)__ffs(rem) - 1u;
		const uint32_t ql = __shfl_sync(0xffffffffu, q, leader);
		const unsigned grp = __ballot_sync(0xffffffffu, has && q == ql);
		uint32_t base = 0;
		if (lane == leader) base = atomicAdd(&tq[ql].candCount, (uint32_t)__popc(grp));
		base = __shfl_sync(0xffffffffu, base, leader);
		if (has && q == ql) {
			const uint32_t at = base + __popc(grp & ((1u << lane) - 1u));
			if (at < candCap) cand[(size_t)q * candCap + at] = key; else tq[q].failed = 1;
		}
		rem &= ~grp;
	}
}

// general resolution of one block per lane: a list has several postings in or below the block (the record's positions are
// scanned further), or more than four lists hit it (their positions are found from the groups first)
template <class U>
__device__ __noinline__ void tr_slow(const IndexView& ix, const QDesc* __restrict__ desc, const U* __restrict__ qnorm, TileQ* __restrict__ tq,
                                     unsigned long long* __restrict__ cand, const uint32_t candCap, const TileRec* __restrict__ recs, const uint32_t ri)
{
	if (ri == INF) return;
	constexpr int CAP = 24;
	const uint4 ra = __ldg(&recs[ri].a), rb = __ldg(&recs[ri].b);
	const uint32_t q = ra.x, blk = ra.y, jp = ra.z & 0xFFFFu, nsr = ra.z >> 16;
	const QDesc& d = desc[q];
	TileQ& s = tq[q];
	const int shift = ix.sieveShift;
	const uint32_t NG = ix.sieveNGroups;
	const uint32_t* __restrict__ postings = ix.postings;
	const uint32_t v = blk >> 5, g = v / 3u, k = v - 3u * g, bit = blk & 31u;
	const uint32_t cRej = s.lastCreq - 1;
	const bool fromRec = nsr <= (uint32_t)SV_PU;
	const uint32_t nIter = fromRec ? nsr : d.nLists;
	uint32_t ids[CAP], mus[CAP], m = 0;
	for (uint32_t it = 0; it < nIter; ++it) {
		uint32_t j, p;
		if (fromRec) { j = (jp >> (4 * it)) & 15u; p = it == 0 ? rb.x : (it == 1 ? rb.y : (it == 2 ? rb.z : rb.w)); }
		else {
			j = it;
			if (!sieve_locate(__ldg(ix.sieveGroups + (size_t)d.listSlot[j] * NG + g), k, bit, p)) continue;
		}
		const uint32_t end = d.listOff[j] + d.listLen[j], mu = d.listMult[j];
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
				if (m < CAP) { ids[m] = id[t]; mus[m] = mu; ++m; } else s.failed = 1;
				any = true;
			}
			p += 4;
		}
		if (!any) s.failed = 1;                              // bitmap and postings disagree: never trust, hand the query over
	}
	for (uint32_t a = 0; a < m; ++a) {
		uint32_t c = 0;
		bool first = true;
		for (uint32_t b = 0; b < m; ++b)
			if (ids[b] == ids[a]) { if (b < a) first = false; c += mus[b]; }
		if (first && c > cRej) tr_emit<U>(ix, d, qnorm, s, cand, candCap, q, ids[a], c);
	}
}

template <class U>
__global__ void __launch_bounds__(TS_WARPS * 32, 12) k_tile_resolve(IndexView ix, const QDesc* __restrict__ desc, const U* __restrict__ qnorm,
                                                                    TileQ* __restrict__ tq, const TileRec* __restrict__ recs, uint32_t recCap,
                                                                    uint32_t* __restrict__ counters, uint32_t round,
                                                                    unsigned long long* __restrict__ cand, uint32_t candCap)
{
	__shared__ uint32_t slowQs[TS_WARPS][64];
	uint32_t* slowQ = slowQs[threadIdx.x >> 5];
	const uint32_t lane = threadIdx.x & 31;
	const int shift = ix.sieveShift;
	const uint32_t* __restrict__ postings = ix.postings;
	const uint32_t* stripeCount = counters + TP_STRIPE0 + TS_STRIPES * round;
	// records per queue (lane l: queues l and l + 32) and the longest queue
	const uint32_t cntA = min(stripeCount[lane], recCap), cntB = min(stripeCount[lane + 32], recCap);
	uint32_t maxCnt = max(cntA, cntB);
#pragma unroll
	for (int o = 16; o; o >>= 1) maxCnt = max(maxCnt, __shfl_xor_sync(0xffffffffu, maxCnt, o));
	uint32_t nSlow = 0;
	const uint32_t nWarps = gridDim.x * TS_WARPS;
	// records were appended 32 at a time: item t = 32 records of queue t % 64; items are dealt out round robin
	for (uint32_t t = blockIdx.x * TS_WARPS + (threadIdx.x >> 5);; t += nWarps) {
		const uint32_t stripe = t & (TS_STRIPES - 1), at = (t / TS_STRIPES) * 32u;
		if (at >= maxCnt) break;
		const uint32_t cnt = __shfl_sync(0xffffffffu, stripe < 32 ? cntA : cntB, stripe & 31);
		if (at >= cnt) continue;
		const uint32_t ri = stripe * recCap + at + lane;
		const uint4 ra = __ldg(&recs[ri].a), rb = __ldg(&recs[ri].b);
		const uint32_t blk = ra.y;
		const bool act = blk != INF;
		const uint32_t q = act ? ra.x : 0u;
		const QDesc& d = desc[q];
		const uint32_t jp = ra.z & 0xFFFFu, nsr = ra.z >> 16;
		const uint32_t ns = act ? min(nsr, (uint32_t)SV_PU) : 0u;
		bool slow = act && nsr > (uint32_t)SV_PU;
		const uint32_t idx[SV_PU] = {rb.x, rb.y, rb.z, rb.w};
		uint32_t h[SV_PU], mu[SV_PU];
		// the wait: three postings per list decide the two common cases — the first one is the block's only posting of the list,
		// or one extra posting of a lower block precedes it (the array is padded; the list's end is checked on the values)
#pragma unroll
		for (int u = 0; u < SV_PU; ++u) {
			const uint32_t j = (jp >> (4 * u)) & 15u;
			const bool on = (uint32_t)u < ns;
			const uint32_t p = on ? idx[u] : 0u;
			const uint32_t end = on ? d.listOff[j] + d.listLen[j] : 0u;
			mu[u] = on ? (uint32_t)d.listMult[j] : 0u;
			const uint32_t id0 = __ldg(postings + p), id1 = __ldg(postings + p + 1), id2 = __ldg(postings + p + 2);
			const uint32_t b0 = p < end ? id0 >> shift : INF, b1 = p + 1 < end ? id1 >> shift : INF, b2 = p + 2 < end ? id2 >> shift : INF;
			h[u] = 0xFFFFFFF0u + (uint32_t)u;                    // no posting: a value of its own
			if (on) {
				if (b0 == blk && b1 != blk) h[u] = id0;
				else if (b0 < blk && b1 == blk && b2 != blk) h[u] = id1;
				else slow = true;
			}
		}
		const uint32_t cRej = act ? tq[q].lastCreq - 1 : 0u;
		const unsigned long long tau = act ? tq[q].tau : 0ull;
#pragma unroll
		for (int u = 0; u < SV_PU; ++u) {
			uint32_t c = mu[u];
			bool first = true;
#pragma unroll
			for (int x2 = 0; x2 < SV_PU; ++x2)
				if (x2 != u && h[x2] == h[u]) { c += mu[x2]; if (x2 < u) first = false; }
			bool has = act && !slow && (uint32_t)u < ns && first && c > cRej;
			unsigned long long key = 0;
			if (has) {
				key = tr_key<U>(ix, d, qnorm, h[u], c);
				has = key < tau;
				if (has && c < d.cprom) atomicAdd(&tq[q].geHist[c], 1u);
			}
			tr_append(tq, cand, candCap, has, q, key, lane);
		}
		// blocks for the general resolution are compacted and resolved a full warp at a time
		const unsigned ms = __ballot_sync(0xffffffffu, slow);
		if (ms) {
			if (slow) slowQ[nSlow + __popc(ms & ((1u << lane) - 1u))] = ri;
			nSlow += __popc(ms);
			__syncwarp();
			if (nSlow >= 32) {
				nSlow -= 32;
				const uint32_t e = slowQ[nSlow + lane];
				__syncwarp();
				tr_slow<U>(ix, desc, qnorm, tq, cand, candCap, recs, e);
				__syncwarp();
			}
		}
	}
	if (nSlow) {
		const uint32_t e = lane < nSlow ? slowQ[lane] : INF;
		tr_slow<U>(ix, desc, qnorm, tq, cand, candCap, recs, e);
	}
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
