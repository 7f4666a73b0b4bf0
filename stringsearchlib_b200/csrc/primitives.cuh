// primitives.cuh — hand-written device-wide primitives used by the index build and the dense search path:
// exclusive scan, stable LSD radix sort (u64 keys, optional u32 payload). No CUB/Thrust.
#pragma once
#include "common.cuh"

namespace ssb {

// out[i] = sum_{j<i} in[j]; *d_total (device, optional) = sum of all. in/out may alias only if same type.
void exclusive_scan_u8_u32(const uint8_t* in, uint32_t* out, size_t n, uint32_t* d_total, cudaStream_t s);
void exclusive_scan_u32_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* d_total, cudaStream_t s);
void exclusive_scan_u32_u64(const uint32_t* in, uint64_t* out, size_t n, uint64_t* d_total, cudaStream_t s);

// Stable LSD radix sort on key bits [bit_lo, bit_hi). keys/vals are ping-ponged with keys_alt/vals_alt;
// returns true when the sorted data ended up in the *_alt buffers. vals may be nullptr (keys only).
bool radix_sort_u64(uint64_t* keys, uint64_t* keys_alt, uint32_t* vals, uint32_t* vals_alt, size_t n, int bit_lo,
                    int bit_hi, cudaStream_t s);

}   // namespace ssb
