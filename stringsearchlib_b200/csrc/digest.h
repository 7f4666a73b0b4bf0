// digest.h — global getSize / getLibSize over key-range shards (digest.cu; SURVEY.md §8e "Global counters")
#pragma once
#include "index.h"

namespace ssb {

// number of 16-byte digests the shard exports: which 0 = terms (getSize, nGramSearch.hpp:488), 1 = grams (getLibSize, hpp:496)
uint64_t digest_count(const DeviceIndex& ix, int which);
// writes them to d_out (device memory, digest_count * 16 bytes); returns synchronised
void export_digests(const DeviceIndex& ix, int which, void* d_out, cudaStream_t s);
// number of distinct digests among d_digests[0..n) whose residue class is `part` of `nParts` (1 part: all of them)
uint64_t count_distinct_digests(const void* d_digests, uint64_t n, uint32_t part, uint32_t nParts, cudaStream_t s);

}   // namespace ssb
