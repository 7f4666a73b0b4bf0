// random_sectors.cu — how many random 32-byte sector reads per second does a B200 sustain out of HBM (footprint >> L2)?
// The denominator for k_sieve_search's candidate-block reads (one sector per posting probe), which a byte roofline does not describe.
// built by build.py into stringsearchlib_b200/lib/random_sectors; `random_sectors --json` prints one line for bench.py
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t z) { z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
template <int MLP>
__global__ void k_rand(const uint32_t* __restrict__ data, uint64_t nSectors, int iters, uint32_t* out)
{
	uint64_t s = mix((uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
	uint32_t acc = 0;
	for (int i = 0; i < iters; ++i) {
		uint32_t v[MLP];
#pragma unroll
		for (int u = 0; u < MLP; ++u) { s = mix(s); v[u] = __ldg(data + (s % nSectors) * 8); }
#pragma unroll
		for (int u = 0; u < MLP; ++u) acc += v[u];
	}
	if (acc == 0x12345678u) out[0] = acc;
}
static bool g_json = false;
static double g_last = 0;
template <int MLP>
void run(const uint32_t* d, uint64_t nSectors, int ctasPerSM, int threads, uint32_t* out)
{
	const int iters = 256 / MLP * 4;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k_rand<MLP><<<148 * ctasPerSM, threads>>>(d, nSectors, 4, out);
	cudaEventRecord(e0);
	k_rand<MLP><<<148 * ctasPerSM, threads>>>(d, nSectors, iters, out);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	const double n = 148.0 * ctasPerSM * threads * iters * MLP;
	g_last = n / ms / 1e6;
	if (!g_json) printf("footprint %.2f GB  warps/SM %d  loads in flight per thread %d: %.1f G sectors/s (%.0f GB/s of 32-byte sectors)\n", nSectors * 32 / 1e9,
	       ctasPerSM * threads / 32, MLP, n / ms / 1e6, n * 32 / ms / 1e6);
}
int main(int argc, char** argv)
{
	const uint64_t bytes = 3ull << 30;
	uint32_t* d; if (cudaMalloc(&d, bytes) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
	cudaMemset(d, 1, bytes);
	uint32_t* out; cudaMalloc(&out, 4);
	if (argc > 1 && !strcmp(argv[1], "--json")) {
		// one line for bench.py: the sustained rate over a footprint far beyond L2 (the index of the headline workload is 1.6 GB)
		g_json = true;
		run<8>(d, bytes / 32, 8, 128, out);
		const double a = g_last;
		run<8>(d, (1ull << 30) / 32, 8, 128, out);
		printf("{\"random_sector_reads_gsectors_per_s\": {\"footprint_3.2GB\": %.2f, \"footprint_1.1GB\": %.2f}, \"sector_bytes\": 32}\n", a, g_last);
		return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
	}
	const uint64_t fps[7] = {bytes / 32, (1ull << 30) / 32, (512ull << 20) / 32, (256ull << 20) / 32, (160ull << 20) / 32, (96ull << 20) / 32, (64ull << 20) / 32};
	for (uint64_t fp : fps) {
		run<4>(d, fp, 8, 128, out); run<8>(d, fp, 16, 128, out);
	}
	return 0;
}
