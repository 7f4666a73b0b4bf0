// merge.cu — multi-GPU top-k merge (no reference equivalent).
// Every key-range shard answers the whole query batch and exports, per query, `limit` fixed-size records
// (score, key length, global first row, key hash), already in rank order. The records travel either through an NCCL
// collective issued by the caller or — fused — straight from the export kernel into the peers' receive buffers over
// NVLink (peer-mapped memory). Each rank then merges a slice of the queries: the per-shard runs are merged with the
// tail of a bitonic network (they are sorted already), equal key strings seen by several shards collapse to their best
// entry (the reference's per-key max, nGramSearch.hpp:326), and the global order (score desc, key length asc, first
// row asc) is cut at `limit`.
#include "search.h"
#include "textops.cuh"
#include <algorithm>

namespace ssb {
namespace {

__device__ __forceinline__ MergeRecord make_record(uint32_t q, uint32_t r, const uint32_t* __restrict__ outCount,
                                                   const uint32_t* __restrict__ outKid, const float* __restrict__ outScore, uint32_t stride,
                                                   const uint32_t* __restrict__ keyOff, const uint64_t* __restrict__ keyRow,
                                                   const uint64_t* __restrict__ keyHash)
{
	MergeRecord rec;
	if (r < outCount[q] && r < stride) {
		const uint32_t kid = outKid[(size_t)q * stride + r];
		rec.scoreBits = __float_as_uint(outScore[(size_t)q * stride + r]);
		rec.keyLen = keyOff[kid + 1] - keyOff[kid] - 1;
		rec.row = keyRow[kid];
		rec.hash = keyHash[kid];
	} else {
		rec.scoreBits = 0; rec.keyLen = 0; rec.row = ~0ull; rec.hash = ~0ull;
	}
	return rec;
}

__global__ void k_export_records(const uint32_t* __restrict__ outCount, const uint32_t* __restrict__ outKid, const float* __restrict__ outScore,
                                 uint32_t stride, uint32_t nq, uint32_t limit, const uint32_t* __restrict__ keyOff,
                                 const uint64_t* __restrict__ keyRow, const uint64_t* __restrict__ keyHash, MergeRecord* __restrict__ out)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (uint64_t)nq * limit) return;
	const uint32_t q = (uint32_t)(i / limit), r = (uint32_t)(i % limit);
	out[i] = make_record(q, r, outCount, outKid, outScore, stride, keyOff, keyRow, keyHash);
}

struct PeerPtrs { MergeRecord* p[16]; };
struct PeerCounts { uint32_t* p[16]; };
struct PeerFlags { uint32_t* p[16]; };        // flags[j][i] on rank j: last epoch whose records rank i has delivered

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// fused export + exchange: the record of query q goes straight into the receive buffer of the rank that merges q
// (peer-mapped memory, plain 8-byte stores over NVLink): recv[j] is laid out [shard][nq/N][limit].
__global__ void k_export_records_p2p(const uint32_t* __restrict__ outCount, const uint32_t* __restrict__ outKid,
                                     const float* __restrict__ outScore, uint32_t stride, uint32_t nq, uint32_t limit,
                                     const uint32_t* __restrict__ keyOff, const uint64_t* __restrict__ keyRow,
                                     const uint64_t* __restrict__ keyHash, PeerPtrs peers, uint32_t nRanks, uint32_t myRank)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < (uint64_t)nq * limit) {
		const uint32_t q = (uint32_t)(i / limit), r = (uint32_t)(i % limit);
		const uint32_t part = nq / nRanks, j = q / part, ql = q - j * part;
		peers.p[j][((size_t)myRank * part + ql) * limit + r] = make_record(q, r, outCount, outKid, outScore, stride, keyOff, keyRow, keyHash);
	}
}

// device-side signal instead of a collective, launched behind the export kernel on the same stream: the records of this rank
// are complete at the kernel boundary; one system-scope fence (cumulative over them) and one release store per peer publish
// "epoch delivered". (Round 1 had the last CTA of the export kernel do this after a system fence in EVERY CTA.)
__global__ void k_signal_epoch(PeerFlags flags, uint32_t nRanks, uint32_t myRank, uint32_t epoch)
{
	__threadfence_system();
	if (threadIdx.x < nRanks) st_release_sys(flags.p[threadIdx.x] + myRank, epoch);
}

__device__ __forceinline__ bool rec_less(uint64_t a1, uint64_t b1, uint64_t a2, uint64_t b2) { return a1 != a2 ? a1 < a2 : b1 < b2; }

// one CTA per query. Shared memory: A (order key hi: ~score | length), B (row), H (hash) of p slots, hash table of 2p u32.
__global__ void k_merge_records(const MergeRecord* __restrict__ gathered, uint32_t nShards, uint32_t nq, uint32_t limit, uint32_t runLen,
                                uint32_t p, MergeRecord* __restrict__ merged, uint32_t* __restrict__ counts, PeerPtrs outPeers,
                                uint32_t nOutPeers, uint32_t outQueryBase, PeerCounts outCountsPeers, const uint32_t* __restrict__ waitFlags,
                                uint32_t epoch)
{
	if (waitFlags) {
		// the runs come straight from the peers' export kernels: wait until every shard has signalled this epoch
		if (threadIdx.x < nShards) {
			uint32_t spins = 0;
			while ((int32_t)(ld_acquire_sys(waitFlags + threadIdx.x) - epoch) < 0)
				if (++spins > (1u << 28)) __trap();              // a peer that never delivers must not hang this GPU
		}
		__syncthreads();
	}
	extern __shared__ __align__(16) unsigned char mg_smem[];
	uint64_t* A = reinterpret_cast<uint64_t*>(mg_smem);
	uint64_t* B = A + p;
	uint64_t* H = B + p;
	uint32_t* table = reinterpret_cast<uint32_t*>(H + p);       // 2p slots, 0xFFFFFFFF = empty
	uint8_t* dead = reinterpret_cast<uint8_t*>(table + 2 * p);   // p flags
	__shared__ uint32_t sAlive;
	const uint32_t q = blockIdx.x;
	// ---- load: run s occupies slots [s*runLen, (s+1)*runLen), ascending for even s, descending for odd s (bitonic layout)
	for (uint32_t i = threadIdx.x; i < p; i += blockDim.x) {
		const uint32_t sh = i / runLen, x = i % runLen;
		const uint32_t r = (sh & 1) ? runLen - 1 - x : x;
		uint64_t a = ~0ull, b = ~0ull, h = ~0ull;
		if (sh < nShards && r < limit) {
			const MergeRecord rec = gathered[((size_t)sh * nq + q) * limit + r];
			if (rec.row != ~0ull) {
				a = ((uint64_t)(~float_to_ordered(__uint_as_float(rec.scoreBits))) << 32) | rec.keyLen;
				b = rec.row; h = rec.hash;
			}
		}
		A[i] = a; B[i] = b; H[i] = h;
		dead[i] = 0;
	}
	for (uint32_t i = threadIdx.x; i < 2 * p; i += blockDim.x) table[i] = 0xFFFFFFFFu;
	if (threadIdx.x == 0) sAlive = 0;
	__syncthreads();
	// ---- the runs are sorted: only the merge stages of the bitonic network remain
	for (uint32_t size = 2 * runLen; size <= p; size <<= 1)
		for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
			for (uint32_t i = threadIdx.x; i < (p >> 1); i += blockDim.x) {
				const uint32_t lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
				const bool up = (lo & size) == 0;
				const bool lt = rec_less(A[hi], B[hi], A[lo], B[lo]);
				if (lt == up) {
					uint64_t t;
					t = A[lo]; A[lo] = A[hi]; A[hi] = t;
					t = B[lo]; B[lo] = B[hi]; B[hi] = t;
					t = H[lo]; H[lo] = H[hi]; H[hi] = t;
				}
			}
			__syncthreads();
		}
	// ---- equal (hash, length) = the same key string seen by several shards: keep the best-ranked entry
	for (uint32_t i = threadIdx.x; i < p; i += blockDim.x) {
		if (A[i] == ~0ull) continue;
		const uint64_t h = H[i];
		const uint32_t len = (uint32_t)A[i];
		uint32_t slot = (uint32_t)(mix64(h) & (2 * p - 1));
		while (true) {
			const uint32_t cur = atomicCAS(&table[slot], 0xFFFFFFFFu, i);
			if (cur == 0xFFFFFFFFu) break;                                   // inserted
			if (H[cur] == h && (uint32_t)A[cur] == len) {
				const uint32_t prev = atomicMin(&table[slot], i);
				dead[max(prev, i)] = 1;                                       // the worse-ranked duplicate dies
				break;
			}
			slot = (slot + 1) & (2 * p - 1);
		}
	}
	__syncthreads();
	// ---- compact the survivors in rank order (block-wide scan in chunks of blockDim) and cut at `limit`
	uint32_t base = 0;
	for (uint32_t c0 = 0; c0 < p && base < limit; c0 += blockDim.x) {
		const uint32_t i = c0 + threadIdx.x;
		const bool alive = i < p && A[i] != ~0ull && !dead[i];
		// warp + block exclusive scan of `alive`
		const unsigned bal = __ballot_sync(0xffffffffu, alive);
		__shared__ uint32_t warpSum[32];
		const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
		if (lane == 0) warpSum[warp] = __popc(bal);
		__syncthreads();
		uint32_t off = 0, tot = 0;
		for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) { if (w < warp) off += warpSum[w]; tot += warpSum[w]; }
		const uint32_t pos = base + off + __popc(bal & ((1u << lane) - 1));
		if (alive && pos < limit) {
			MergeRecord rec;
			rec.scoreBits = __float_as_uint(ordered_to_float(~(uint32_t)(A[i] >> 32)));
			rec.keyLen = (uint32_t)A[i];
			rec.row = B[i]; rec.hash = H[i];
			if (nOutPeers) {
				for (uint32_t d = 0; d < nOutPeers; ++d) outPeers.p[d][((size_t)(outQueryBase + q)) * limit + pos] = rec;
			} else merged[(size_t)q * limit + pos] = rec;
		}
		base += tot;
		__syncthreads();
	}
	const uint32_t n = min(base, limit);
	for (uint32_t i = n + threadIdx.x; i < limit; i += blockDim.x) {
		MergeRecord rec;
		rec.scoreBits = 0; rec.keyLen = 0; rec.row = ~0ull; rec.hash = ~0ull;
		if (nOutPeers) {
			for (uint32_t d = 0; d < nOutPeers; ++d) outPeers.p[d][((size_t)(outQueryBase + q)) * limit + i] = rec;
		} else merged[(size_t)q * limit + i] = rec;
	}
	if (threadIdx.x == 0) {
		if (nOutPeers) { for (uint32_t d = 0; d < nOutPeers; ++d) outCountsPeers.p[d][outQueryBase + q] = n; }
		else counts[q] = n;
	}
	(void)sAlive;
}

struct MergeGeom { uint32_t runLen, p; size_t smem; unsigned threads; };
MergeGeom merge_geometry(uint32_t nShards, uint32_t limit)
{
	MergeGeom g;
	g.runLen = 1;
	while (g.runLen < limit) g.runLen <<= 1;
	uint32_t runs = 1;
	while (runs < nShards) runs <<= 1;
	g.p = std::max<uint32_t>(2, g.runLen * runs);
	g.smem = (size_t)g.p * (24 + 8 + 1) + 16;
	g.threads = std::min<uint32_t>(1024, std::max<uint32_t>(64, g.p / 2));
	return g;
}
void set_merge_smem(size_t smem)
{
	static size_t attr = 0;
	if (smem > 220 * 1024) throw std::runtime_error("merge_records: limit too large for one CTA");
	if (smem > attr) {
		SSB_CUDA(cudaFuncSetAttribute(k_merge_records, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		attr = smem;
	}
}

}   // namespace

void export_records(const DeviceIndex& ix, SearchWorkspace& ws, uint32_t nq, uint32_t limit, MergeRecord* d_out, cudaStream_t s)
{
	if (limit == 0 || limit > FAST_KLIM) throw std::runtime_error("export_records: limit must be in [1, FAST_KLIM] on the sharded path");
	if (nq > ws.capQ) throw std::runtime_error("export_records: no such batch");
	const uint64_t n = (uint64_t)nq * limit;
	if (!ix.indexed || ix.K == 0) {
		SSB_CUDA(cudaMemsetAsync(d_out, 0xFF, n * sizeof(MergeRecord), s));     // an empty shard contributes only empty slots
		return;
	}
	k_export_records<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ws.outCount.get(), ws.outKid.get(), ws.outScore.get(), limit, nq, limit,
	                                                             ix.keyOff.get(), ix.keyRow.get(), ix.keyHash.get(), d_out);
	SSB_LAUNCH_CHECK();
}

void export_records_p2p(const DeviceIndex& ix, SearchWorkspace& ws, uint32_t nq, uint32_t limit, MergeRecord* const* peerRecv,
                        uint32_t nRanks, uint32_t myRank, cudaStream_t s, uint32_t* const* peerFlags, uint32_t epoch)
{
	if (limit == 0 || limit > FAST_KLIM) throw std::runtime_error("export_records_p2p: limit must be in [1, FAST_KLIM]");
	if (nq > ws.capQ || nRanks == 0 || nRanks > 16 || nq % nRanks) throw std::runtime_error("export_records_p2p: bad geometry");
	if (!ix.indexed || ix.K == 0) throw std::runtime_error("export_records_p2p: empty shard");
	PeerPtrs pp{};
	for (uint32_t i = 0; i < nRanks; ++i) pp.p[i] = peerRecv[i];
	const uint64_t n = (uint64_t)nq * limit;
	k_export_records_p2p<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ws.outCount.get(), ws.outKid.get(), ws.outScore.get(), limit, nq, limit,
	                                                                 ix.keyOff.get(), ix.keyRow.get(), ix.keyHash.get(), pp, nRanks, myRank);
	SSB_LAUNCH_CHECK();
	if (peerFlags) {
		PeerFlags pf{};
		for (uint32_t i = 0; i < nRanks; ++i) pf.p[i] = peerFlags[i];
		k_signal_epoch<<<1, 32, 0, s>>>(pf, nRanks, myRank, epoch);
		SSB_LAUNCH_CHECK();
	}
}

// merges nShards runs per query; when nShards * pow2(limit) exceeds one CTA's shared memory the shards are merged in
// groups first (top-limit of a union is contained in the union of the groups' top-limits; best-entry dedupe is associative)
constexpr uint32_t MERGE_MAX_SLOTS = 4096;
static void merge_impl(const MergeRecord* d_gathered, uint32_t nShards, uint32_t nq, uint32_t limit, MergeRecord* d_out, uint32_t* d_outCount,
                       const PeerPtrs* peers, const PeerCounts* peerCounts, uint32_t nRanks, uint32_t myRank, cudaStream_t s,
                       const uint32_t* waitFlags = nullptr, uint32_t epoch = 0)
{
	MergeGeom g = merge_geometry(nShards, limit);
	bool grouped = false;                       // the group merges have waited for the peers already
	DevBuf<MergeRecord> tmp;
	DevBuf<uint32_t> tmpCount;
	while (g.p > MERGE_MAX_SLOTS && nShards > 2) {
		const uint32_t group = std::max<uint32_t>(2, MERGE_MAX_SLOTS / g.runLen);
		const uint32_t nGroups = (nShards + group - 1) / group;
		DevBuf<MergeRecord> next((size_t)nGroups * nq * limit, s);
		tmpCount.reserve(nq, s);
		const MergeGeom gg = merge_geometry(group, limit);
		set_merge_smem(gg.smem);
		for (uint32_t gi = 0; gi < nGroups; ++gi) {
			const uint32_t first = gi * group, cnt = std::min(group, nShards - first);
			PeerPtrs none{}; PeerCounts nonec{};
			k_merge_records<<<nq, gg.threads, gg.smem, s>>>(d_gathered + (size_t)first * nq * limit, cnt, nq, limit, gg.runLen, gg.p,
			                                              next.get() + (size_t)gi * nq * limit, tmpCount.get(), none, 0, 0, nonec,
			                                              waitFlags ? waitFlags + first : nullptr, epoch);
			SSB_LAUNCH_CHECK();
		}
		tmp = std::move(next);
		grouped = true;
		d_gathered = tmp.get();
		nShards = nGroups;
		g = merge_geometry(nShards, limit);
	}
	set_merge_smem(g.smem);
	PeerPtrs pp{}; PeerCounts pc{};
	if (peers) { pp = *peers; pc = *peerCounts; }
	k_merge_records<<<nq, g.threads, g.smem, s>>>(d_gathered, nShards, nq, limit, g.runLen, g.p, d_out, d_outCount, pp, peers ? nRanks : 0,
	                                              peers ? myRank * nq : 0, pc, grouped ? nullptr : waitFlags, epoch);
	SSB_LAUNCH_CHECK();
}

void merge_records(const MergeRecord* d_gathered, uint32_t nShards, uint32_t nq, uint32_t limit, MergeRecord* d_out, uint32_t* d_outCount,
                   cudaStream_t s)
{
	merge_impl(d_gathered, nShards, nq, limit, d_out, d_outCount, nullptr, nullptr, 0, 0, s);
}

void merge_records_p2p(const MergeRecord* d_recv, uint32_t nShards, uint32_t nqLocal, uint32_t limit, MergeRecord* const* peerMerged,
                       uint32_t* const* peerCounts, uint32_t nRanks, uint32_t myRank, cudaStream_t s, const uint32_t* waitFlags, uint32_t epoch)
{
	if (nRanks == 0 || nRanks > 16 || (waitFlags && nShards > 16)) throw std::runtime_error("merge_records_p2p: bad geometry");
	PeerPtrs pp{}; PeerCounts pc{};
	for (uint32_t i = 0; i < nRanks; ++i) { pp.p[i] = peerMerged[i]; pc.p[i] = peerCounts[i]; }
	merge_impl(d_recv, nShards, nqLocal, limit, nullptr, nullptr, &pp, &pc, nRanks, myRank, s, waitFlags, epoch);
}

}   // namespace ssb
