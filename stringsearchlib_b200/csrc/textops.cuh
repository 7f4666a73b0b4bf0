// textops.cuh — device-side restatement of the reference's normalisation helpers
// (ltrim/rtrim/toUpper/escapeBlank, nGramSearch.h:30-98) on code units, plus symbol lookup for gram codes.
#pragma once
#include "common.cuh"

namespace ssb {

// std::isspace / std::iswspace in the "C" locale: ' ', \t \n \v \f \r
__device__ __forceinline__ bool unit_space(uint32_t u) { return u == 32u || (u - 9u) <= 4u; }

// escapeBlank keeps a unit iff it is in validChar. Narrow: bytes >= 0x80 are never valid by default.
// Wide (unpinned; nGramSearch.h:306 "characters in the ASCII range"): units >= 128 are always kept.
template <class U> __device__ __forceinline__ bool unit_valid(U u, const uint8_t* __restrict__ lut)
{
	if (sizeof(U) == 1) return lut[(uint8_t)u] != 0;
	return (uint32_t)u >= 128u || lut[(uint32_t)u] != 0;
}
template <class U> __device__ __forceinline__ U unit_escape(U u, const uint8_t* __restrict__ lut)
{
	return unit_valid(u, lut) ? u : (U)' ';
}
// escapeBlank + toUpper ("C" locale: a-z only)
template <class U> __device__ __forceinline__ U unit_norm(U u, const uint8_t* __restrict__ lut)
{
	if (!unit_valid(u, lut)) return (U)' ';
	return (u >= (U)'a' && u <= (U)'z') ? (U)(u - 32) : u;
}
// true when the unit is trimmed after escapeBlank
template <class U> __device__ __forceinline__ bool unit_blank(U u, const uint8_t* __restrict__ lut)
{
	return !unit_valid(u, lut) || unit_space((uint32_t)u);
}

// symbol id of a unit for gram codes: [1, A], 0 = the unit occurs in no long term
template <class U> __device__ __forceinline__ uint32_t sym_of(U u, const uint32_t* __restrict__ symTable, uint32_t A)
{
	if (sizeof(U) == 1) return symTable[(uint8_t)u];
	uint32_t lo = 0, hi = A;
	while (lo < hi) {
		uint32_t mid = (lo + hi) >> 1;
		if (symTable[mid] < (uint32_t)u) lo = mid + 1; else hi = mid;
	}
	return (lo < A && symTable[lo] == (uint32_t)u) ? lo + 1 : 0;
}

__device__ __forceinline__ uint64_t hash_step(uint64_t h, uint32_t u) { return (h ^ u) * 0x100000001B3ull; }
__device__ __forceinline__ uint64_t hash_done(uint64_t h, uint32_t len) { return mix64(h ^ ((uint64_t)len * 0x9E3779B97F4A7C15ull)); }

}   // namespace ssb
