// common.cuh — shared host/device helpers for the B200 n-gram engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <atomic>

namespace ssb {

struct CudaError : std::runtime_error {
	using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line)
{
	if (e != cudaSuccess) {
		char buf[512];
		snprintf(buf, sizeof buf, "CUDA error %s (%d) at %s:%d: %s", cudaGetErrorName(e), (int)e, file, line, what);
		throw CudaError(buf);
	}
}
#define SSB_CUDA(x) ::ssb::cuda_check((x), #x, __FILE__, __LINE__)
// every kernel launch of the engine goes through SSB_LAUNCH_CHECK(): it validates the launch and counts it
// (ssb_kernel_launches() feeds bench.py's "gpu_launches")
inline std::atomic<uint64_t> g_kernel_launches{0};
// set by an atexit handler registered at engine start-up, i.e. BEFORE the static destructors of this library run: objects
// that are still alive then (cached staging buffers, indexes a host never disposed of) must not call into a CUDA runtime
// that is tearing down — the process is about to return everything to the OS anyway
inline std::atomic<bool> g_exiting{false};
// which kernel answers the fast queries first: 1 = k_sieve_search (block bitmaps, reads only the postings of blocks that can
// matter; hands what it does not take to k_fast_search), 0 = k_fast_search alone (register K-way merge of the posting lists).
// Both are exact; DESIGN.md §4.2 has the measurements behind the default.
inline std::atomic<int> g_search_path{-1};       // -1: default (1, or 0 when SSB_MERGE is set in the environment)
#define SSB_LAUNCH_CHECK() (::ssb::g_kernel_launches.fetch_add(1, std::memory_order_relaxed), \
                            ::ssb::cuda_check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__))

// Owning device buffer. Allocation is stream-ordered (cudaMallocAsync) so that build/search pipelines
// never serialise on cudaMalloc/cudaFree; the pool keeps freed blocks for reuse.
template <class T>
class DevBuf {
public:
	DevBuf() = default;
	explicit DevBuf(size_t n, cudaStream_t s = 0) { alloc(n, s); }
	DevBuf(const DevBuf&) = delete;
	DevBuf& operator=(const DevBuf&) = delete;
	DevBuf(DevBuf&& o) noexcept { swap(o); }
	DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { release(); swap(o); } return *this; }
	~DevBuf() { release(); }
	void alloc(size_t n, cudaStream_t s = 0)
	{
		release();
		n_ = n; stream_ = s;
		if (n) SSB_CUDA(cudaMallocAsync((void**)&p_, n * sizeof(T), s));
	}
	// grow-only (contents are NOT preserved)
	void reserve(size_t n, cudaStream_t s = 0) { if (n > n_) alloc(n + n / 4, s); }
	void release()
	{
		if (p_ && !g_exiting.load(std::memory_order_relaxed)) cudaFreeAsync(p_, stream_);
		p_ = nullptr; n_ = 0;
	}
	void zero(cudaStream_t s) { if (n_) SSB_CUDA(cudaMemsetAsync(p_, 0, n_ * sizeof(T), s)); }
	T* get() const { return p_; }
	size_t size() const { return n_; }
	size_t bytes() const { return n_ * sizeof(T); }
	void swap(DevBuf& o) { std::swap(p_, o.p_); std::swap(n_, o.n_); std::swap(stream_, o.stream_); }
private:
	T* p_ = nullptr;
	size_t n_ = 0;
	cudaStream_t stream_ = 0;
};

// Pinned host buffer (H2D/D2H staging).
template <class T>
class PinBuf {
public:
	PinBuf() = default;
	PinBuf(const PinBuf&) = delete;
	PinBuf& operator=(const PinBuf&) = delete;
	~PinBuf() { if (p_ && !g_exiting.load(std::memory_order_relaxed)) cudaFreeHost(p_); }
	void reserve(size_t n)
	{
		if (n <= n_) return;
		if (p_) cudaFreeHost(p_);
		p_ = nullptr;
		n_ = n + n / 4 + 16;
		SSB_CUDA(cudaMallocHost((void**)&p_, n_ * sizeof(T)));
	}
	// exactly n elements (contents are NOT preserved)
	void alloc_exact(size_t n)
	{
		if (p_) cudaFreeHost(p_);
		p_ = nullptr;
		n_ = n;
		if (n) SSB_CUDA(cudaMallocHost((void**)&p_, n_ * sizeof(T)));
	}
	const T* data() const { return p_; }
	T* get() const { return p_; }
	size_t size() const { return n_; }
private:
	T* p_ = nullptr;
	size_t n_ = 0;
};

__host__ __device__ inline uint64_t mix64(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

// order-preserving float -> u32 (ascending), and the inverse
__host__ __device__ inline uint32_t float_to_ordered(float f)
{
#ifdef __CUDA_ARCH__
	uint32_t b = __float_as_uint(f);
#else
	uint32_t b; memcpy(&b, &f, 4);
#endif
	return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__host__ __device__ inline float ordered_to_float(uint32_t o)
{
	uint32_t b = o ^ ((o >> 31) ? 0x80000000u : 0xFFFFFFFFu);
#ifdef __CUDA_ARCH__
	return __uint_as_float(b);
#else
	float f; memcpy(&f, &b, 4); return f;
#endif
}
// rank key of a result: smaller is better. (score desc, kid asc); kid already encodes (key length asc, key id asc).
__host__ __device__ inline uint64_t rank_key(float score, uint32_t kid)
{
	return ((uint64_t)(~float_to_ordered(score)) << 32) | kid;
}
__host__ __device__ inline float rank_key_score(uint64_t k) { return ordered_to_float(~(uint32_t)(k >> 32)); }

inline int ceil_div(size_t a, size_t b) { return (int)((a + b - 1) / b); }

}   // namespace ssb
