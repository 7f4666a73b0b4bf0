// primitives.cu — device-wide exclusive scan and stable LSD radix sort, hand-written for sm_100a.
//
// Radix sort: 8-bit digits, one pass = {tile histogram, per-digit scan, stable scatter}. The scatter kernel
// ranks items with warp match-any (immune to skewed digits such as gram codes), stages the tile in shared
// memory in sorted order and writes digit runs back with coalesced stores. Traffic per pass: 2 reads + 1 write
// of the key(+payload) arrays, all 16B-vectorised / fully coalesced on the read side.
#include "primitives.cuh"

namespace ssb {

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <class T> __device__ __forceinline__ T warp_incl_scan(T v)
{
	const int lane = threadIdx.x & 31;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		T t = __shfl_up_sync(0xffffffffu, v, o);
		if (lane >= o) v += t;
	}
	return v;
}

// block-wide exclusive scan of one value per thread (SCAN_THREADS threads); returns exclusive prefix, total in *total
template <class T> __device__ __forceinline__ T block_excl_scan(T v, T* smem /*[32]*/, T& total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	T inc = warp_incl_scan(v);
	if (lane == 31) smem[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		T w = lane < (blockDim.x >> 5) ? smem[lane] : T(0);
		T wi = warp_incl_scan(w);
		smem[lane] = wi - w;
		if (lane == 31) smem[32] = wi;
	}
	__syncthreads();
	T res = smem[warp] + inc - v;
	total = smem[32];
	__syncthreads();
	return res;
}

template <class InT, class OutT>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const InT* __restrict__ in, OutT* __restrict__ tileSums, size_t n)
{
	__shared__ OutT sm[33];
	size_t base = (size_t)blockIdx.x * SCAN_TILE;
	OutT acc = 0;
#pragma unroll
	for (int j = 0; j < SCAN_ITEMS; ++j) {
		size_t i = base + (size_t)j * SCAN_THREADS + threadIdx.x;
		if (i < n) acc += (OutT)in[i];
	}
	OutT total;
	block_excl_scan(acc, sm, total);
	if (threadIdx.x == 0) tileSums[blockIdx.x] = total;
}

template <class InT, class OutT>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const InT* __restrict__ in, OutT* __restrict__ out,
                                                             const OutT* __restrict__ tileBase, size_t n, OutT* d_total)
{
	__shared__ OutT sm[33];
	size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
	OutT v[SCAN_ITEMS];
	OutT acc = 0;
#pragma unroll
	for (int j = 0; j < SCAN_ITEMS; ++j) {
		size_t i = base + j;
		v[j] = i < n ? (OutT)in[i] : OutT(0);
		acc += v[j];
	}
	OutT total;
	OutT pre = block_excl_scan(acc, sm, total) + (tileBase ? tileBase[blockIdx.x] : OutT(0));
#pragma unroll
	for (int j = 0; j < SCAN_ITEMS; ++j) {
		size_t i = base + j;
		if (i < n) out[i] = pre;
		pre += v[j];
	}
	if (d_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) *d_total = pre;
}

template <class InT, class OutT>
void exclusive_scan_impl(const InT* in, OutT* out, size_t n, OutT* d_total, cudaStream_t s)
{
	if (n == 0) {
		if (d_total) SSB_CUDA(cudaMemsetAsync(d_total, 0, sizeof(OutT), s));
		return;
	}
	size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
	if (tiles == 1) {
		k_scan_apply<InT, OutT><<<1, SCAN_THREADS, 0, s>>>(in, out, nullptr, n, d_total);
		SSB_LAUNCH_CHECK();
		return;
	}
	DevBuf<OutT> sums(tiles, s);
	k_scan_reduce<InT, OutT><<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(in, sums.get(), n);
	SSB_LAUNCH_CHECK();
	exclusive_scan_impl<OutT, OutT>(sums.get(), sums.get(), tiles, nullptr, s);
	k_scan_apply<InT, OutT><<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(in, out, sums.get(), n, d_total);
	SSB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------ radix sort
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // 4096 keys per CTA
constexpr int RS_BINS = 256;

__device__ __forceinline__ uint32_t digit_of(uint64_t k, int shift, uint32_t mask) { return (uint32_t)(k >> shift) & mask; }

// tileHist[d * nTiles + tile] = #keys of the tile whose digit is d
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t* __restrict__ keys, size_t n, int shift, uint32_t mask,
                                                        uint32_t* __restrict__ tileHist, uint32_t nTiles)
{
	__shared__ uint32_t hist[RS_BINS];
	hist[threadIdx.x] = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31;
	size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll 4
	for (int j = 0; j < RS_ITEMS; ++j) {
		size_t i = base + (size_t)j * RS_THREADS + threadIdx.x;
		bool ok = i < n;
		uint32_t d = ok ? digit_of(keys[i], shift, mask) : RS_BINS;
		unsigned m = __match_any_sync(0xffffffffu, d);
		if (ok && (m & ((1u << lane) - 1)) == 0) atomicAdd(&hist[d], __popc(m));
	}
	__syncthreads();
	tileHist[(size_t)threadIdx.x * nTiles + blockIdx.x] = hist[threadIdx.x];
}

// one CTA per digit: in-place exclusive scan of its row of nTiles counts; digitTotal[d] = row sum
__global__ void __launch_bounds__(RS_THREADS) k_rs_scan_rows(uint32_t* __restrict__ tileHist, uint32_t nTiles,
                                                             uint32_t* __restrict__ digitTotal)
{
	__shared__ uint32_t sm[33];
	uint32_t* row = tileHist + (size_t)blockIdx.x * nTiles;
	uint32_t carry = 0;
	for (uint32_t b = 0; b < nTiles; b += RS_THREADS * 4) {
		uint32_t v[4], acc = 0;
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			uint32_t i = b + threadIdx.x * 4 + j;
			v[j] = i < nTiles ? row[i] : 0;
			acc += v[j];
		}
		uint32_t total;
		uint32_t pre = block_excl_scan(acc, sm, total) + carry;
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			uint32_t i = b + threadIdx.x * 4 + j;
			if (i < nTiles) row[i] = pre;
			pre += v[j];
		}
		carry += total;
	}
	if (threadIdx.x == 0) digitTotal[blockIdx.x] = carry;
}

template <bool HAS_VAL>
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                           uint64_t* __restrict__ outKeys, uint32_t* __restrict__ outVals,
                                                           size_t n, int shift, uint32_t mask,
                                                           const uint32_t* __restrict__ tileHist, uint32_t nTiles,
                                                           const uint32_t* __restrict__ digitTotal)
{
	__shared__ uint32_t warpCnt[RS_WARPS][RS_BINS + 1];
	__shared__ uint32_t tileOff[RS_BINS];       // exclusive offset of digit d inside the sorted tile
	__shared__ uint32_t globalBase[RS_BINS];    // global position of the tile's first key with digit d
	__shared__ uint32_t scanTmp[33];
	extern __shared__ __align__(16) unsigned char rs_dyn_smem[];
	uint64_t* sKeys = reinterpret_cast<uint64_t*>(rs_dyn_smem);
	uint32_t* sVals = reinterpret_cast<uint32_t*>(rs_dyn_smem + sizeof(uint64_t) * RS_TILE);

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int i = threadIdx.x; i < RS_WARPS * (RS_BINS + 1); i += RS_THREADS) (&warpCnt[0][0])[i] = 0;
	// global digit bases: exclusive scan of digitTotal over digits (256 values, one per thread)
	uint32_t dt = digitTotal[threadIdx.x], dtTotal;
	uint32_t dBase = block_excl_scan(dt, scanTmp, dtTotal);   // contains __syncthreads -> warpCnt zeroed
	globalBase[threadIdx.x] = dBase + tileHist[(size_t)threadIdx.x * nTiles + blockIdx.x];

	const size_t tileBase = (size_t)blockIdx.x * RS_TILE;
	const size_t warpBase = tileBase + (size_t)warp * (RS_ITEMS * 32);
	uint64_t k[RS_ITEMS];
	uint32_t v[RS_ITEMS];
	uint16_t rank[RS_ITEMS];
	uint16_t dig[RS_ITEMS];
#pragma unroll
	for (int j = 0; j < RS_ITEMS; ++j) {
		size_t i = warpBase + (size_t)j * 32 + lane;
		bool ok = i < n;
		k[j] = ok ? keys[i] : 0;
		if (HAS_VAL) v[j] = ok ? vals[i] : 0;
		dig[j] = ok ? (uint16_t)digit_of(k[j], shift, mask) : (uint16_t)RS_BINS;
	}
#pragma unroll
	for (int j = 0; j < RS_ITEMS; ++j) {
		uint32_t d = dig[j];
		unsigned m = __match_any_sync(0xffffffffu, d);
		unsigned lt = m & ((1u << lane) - 1);
		uint32_t base = warpCnt[warp][d];
		__syncwarp();
		if (lt == 0) warpCnt[warp][d] = base + __popc(m);
		__syncwarp();
		rank[j] = (uint16_t)(base + __popc(lt));
	}
	__syncthreads();
	// per digit: exclusive scan across warps, tile digit count
	{
		uint32_t d = threadIdx.x, run = 0;
#pragma unroll
		for (int w = 0; w < RS_WARPS; ++w) {
			uint32_t c = warpCnt[w][d];
			warpCnt[w][d] = run;
			run += c;
		}
		uint32_t tot;
		uint32_t off = block_excl_scan(run, scanTmp, tot);
		tileOff[d] = off;
	}
	__syncthreads();
#pragma unroll
	for (int j = 0; j < RS_ITEMS; ++j) {
		uint32_t d = dig[j];
		if (d < RS_BINS) {
			uint32_t p = tileOff[d] + warpCnt[warp][d] + rank[j];
			sKeys[p] = k[j];
			if (HAS_VAL) sVals[p] = v[j];
		}
	}
	__syncthreads();
	const uint32_t cnt = (uint32_t)min((size_t)RS_TILE, n - tileBase);
	for (uint32_t i = threadIdx.x; i < cnt; i += RS_THREADS) {
		uint64_t key = sKeys[i];
		uint32_t d = digit_of(key, shift, mask);
		size_t gp = (size_t)globalBase[d] + (i - tileOff[d]);
		outKeys[gp] = key;
		if (HAS_VAL) outVals[gp] = sVals[i];
	}
}

}   // namespace

void exclusive_scan_u8_u32(const uint8_t* in, uint32_t* out, size_t n, uint32_t* d_total, cudaStream_t s)
{
	exclusive_scan_impl<uint8_t, uint32_t>(in, out, n, d_total, s);
}
void exclusive_scan_u32_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* d_total, cudaStream_t s)
{
	exclusive_scan_impl<uint32_t, uint32_t>(in, out, n, d_total, s);
}
void exclusive_scan_u32_u64(const uint32_t* in, uint64_t* out, size_t n, uint64_t* d_total, cudaStream_t s)
{
	exclusive_scan_impl<uint32_t, uint64_t>(in, out, n, d_total, s);
}

bool radix_sort_u64(uint64_t* keys, uint64_t* keys_alt, uint32_t* vals, uint32_t* vals_alt, size_t n, int bit_lo,
                    int bit_hi, cudaStream_t s)
{
	if (n < 2 || bit_hi <= bit_lo) return false;
	if (n >= (1ull << 32)) throw std::runtime_error("radix_sort_u64: more than 2^32-1 items");
	const uint32_t nTiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
	DevBuf<uint32_t> tileHist((size_t)RS_BINS * nTiles, s);
	DevBuf<uint32_t> digitTotal(RS_BINS, s);
	const size_t smemK = sizeof(uint64_t) * RS_TILE, smemKV = smemK + sizeof(uint32_t) * RS_TILE;
	static bool attrSet = false;
	if (!attrSet) {
		SSB_CUDA(cudaFuncSetAttribute(k_rs_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemKV));
		SSB_CUDA(cudaFuncSetAttribute(k_rs_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemK));
		attrSet = true;
	}
	bool flipped = false;
	for (int shift = bit_lo; shift < bit_hi; shift += 8) {
		int bits = bit_hi - shift < 8 ? bit_hi - shift : 8;
		uint32_t mask = (1u << bits) - 1;
		const uint64_t* kin = flipped ? keys_alt : keys;
		uint64_t* kout = flipped ? keys : keys_alt;
		const uint32_t* vin = flipped ? vals_alt : vals;
		uint32_t* vout = flipped ? vals : vals_alt;
		k_rs_hist<<<nTiles, RS_THREADS, 0, s>>>(kin, n, shift, mask, tileHist.get(), nTiles);
		SSB_LAUNCH_CHECK();
		k_rs_scan_rows<<<RS_BINS, RS_THREADS, 0, s>>>(tileHist.get(), nTiles, digitTotal.get());
		SSB_LAUNCH_CHECK();
		if (vals)
			k_rs_scatter<true><<<nTiles, RS_THREADS, smemKV, s>>>(kin, vin, kout, vout, n, shift, mask, tileHist.get(), nTiles,
			                                                digitTotal.get());
		else
			k_rs_scatter<false><<<nTiles, RS_THREADS, smemK, s>>>(kin, nullptr, kout, nullptr, n, shift, mask, tileHist.get(),
			                                                 nTiles, digitTotal.get());
		SSB_LAUNCH_CHECK();
		flipped = !flipped;
	}
	return flipped;
}

}   // namespace ssb
