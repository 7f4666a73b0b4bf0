// index.h — device-resident n-gram index (one key-range shard) and the host-side views of it.
//
// Replaces the reference's hash containers (nGramSearch.h:279-298):
//   stringLib   -> termChars/termOff (normalised terms, by term id)  +  keyChars/keyOff (raw trimmed keys, by kid)
//   wordMap     -> termEdgeOff/edgeKid           (CSR term -> keys)      [+ reverse CSR key -> terms]
//   wordWeight  -> edgeW                         (parallel to edgeKid)
//   ngrams      -> gramCodes (sorted) / gramOff / postings (CSR gram -> ascending term ids)
//   longLib/shortLib -> implicit (term length >= 2*gSize) + shortTerms list
// Id scheme (the CPU restatement used by the parity tests defines the same ids, which is how parity is bit-exact):
//   kid     = rank of a key by (length asc, first-occurrence asc)   -> ScoreComparer tie-break == kid asc
//   term id = rank of a term by (min kid over its edges, first word index)
#pragma once
#include "common.cuh"
#include <vector>

namespace ssb {

constexpr int MAX_GSIZE = 8;
constexpr uint32_t SIEVE_GROUP_BLOCKS = 96;   // blocks per sieve group (sieve.cu)
// Two layouts of the same bitmaps (sieve.cu): 16-byte groups { first, 3 x 32 block bits } and, on indexes of up to
// SIEVE_INLINE_MAX_TERMS terms, 32-byte groups that add 128 bits of SLOTS: for the group's r-th occupied block, (shift + 1) bits
// at position r * (shift + 1): the offset of the list's first posting inside the block and a flag "the block holds more postings
// of the list". Block number + offset IS the term id, so k_sieve_search<.., true> resolves a candidate block from the sector it
// has just streamed instead of a scattered read of the posting list — at twice the stream. Which one a batch uses: search.cu.
constexpr uint32_t SIEVE_INLINE_MAX_TERMS = 32u << 20;
__host__ __device__ inline uint32_t sieve_slot_bits(int shift) { return (uint32_t)shift + 1u; }
__host__ __device__ inline uint32_t sieve_slots_per_group(int shift) { const uint32_t n = 126u / sieve_slot_bits(shift); return n < SIEVE_GROUP_BLOCKS ? n : SIEVE_GROUP_BLOCKS; }

struct DeviceIndex {
	// configuration
	int g = 3;            // gram size
	int unit = 1;         // bytes per code unit: 1 = char, 4 = wchar_t (UTF-32)
	uint64_t rowBase = 0; // global index of this shard's first row
	bool indexed = false; // nGramSearch.h:301

	// sizes
	uint32_t K = 0;       // distinct keys (trimmed raw key strings of valid rows)
	uint32_t T = 0;       // distinct terms                       == getSize()    (nGramSearch.hpp:488-491)
	uint32_t E = 0;       // distinct (term, key) edges
	uint32_t G = 0;       // distinct grams                       == getLibSize() (nGramSearch.hpp:496-499)
	uint32_t P = 0;       // postings
	uint32_t nShort = 0;  // terms shorter than 2*gSize (nGramSearch.hpp:82-85)
	uint32_t maxTermLen = 0, maxKeyLen = 0;
	uint32_t termUnits = 0, keyUnits = 0;   // total code units of termChars / keyChars (keys incl. terminators)

	// fast-path properties
	bool identity = false;   // every term has exactly one edge and edgeKid[t] == t (rowSize==1, distinct keys)
	bool uniformW = false;   // every edge weight == 1.0f
	bool multiEdgeKeys = false;   // some key is reachable through more than one term
	float wmax = 1.0f;       // max edge weight (upper bound filter)

	// alphabet remap: gram code = sum sym(c_i) << symBits*(g-1-i); sym in [1, A], 0 = unit absent from every long term
	int symBits = 8;
	uint32_t A = 0;
	bool packed = true;      // g*symBits <= 32
	DevBuf<uint32_t> symTable;   // narrow: 256-entry LUT unit -> sym ; wide: A sorted units (sym = rank+1)

	DevBuf<uint8_t> validLut;    // 256 entries, validChar (nGramSearch.h:307-313)

	DevBuf<uint64_t> gramCodes;  // [G] ascending
	DevBuf<uint32_t> gramOff;    // [G+1]
	DevBuf<uint32_t> postings;   // [P]

	DevBuf<uint32_t> termOff;    // [T+1] in units
	DevBuf<uint8_t> termChars;   // normalised terms
	DevBuf<uint32_t> termEdgeOff;// [T+1]
	DevBuf<uint32_t> edgeKid;    // [E] ascending kid per term
	DevBuf<float> edgeW;         // [E]
	DevBuf<uint32_t> shortTerms; // [nShort]

	DevBuf<uint32_t> keyOff;     // [K+1] in units; every key is followed by one 0 unit
	DevBuf<uint8_t> keyChars;    // raw trimmed keys by kid
	DevBuf<uint32_t> keyEdgeOff; // [K+1] reverse CSR
	DevBuf<uint32_t> keyEdgeTerm;// [E]
	DevBuf<float> keyEdgeW;      // [E]
	DevBuf<float> keyWmax;       // [K] wildcard score (max edge weight), -inf when the key has no edge
	DevBuf<uint64_t> keyRow;     // [K] global first row of the key
	DevBuf<uint64_t> keyHash;    // [K] 64-bit hash of the key string (shard-independent; multi-GPU merge identity)

	// sieve: block bitmaps over the posting lists (sieve.cu). Derived from gramOff/postings, rebuilt after a load.
	int sieveShift = 0;          // a block = 2^sieveShift consecutive term ids
	uint32_t sieveNGroups = 0;   // 96-block groups per bitmap; 0 = the index has no sieve
	uint32_t sieveSlots = 0;     // grams that have a bitmap
	DevBuf<uint32_t> sieveSlot;  // [G] bitmap slot of the gram, 0xFFFFFFFF = none (short list)
	DevBuf<uint32_t> sieveGroups;// [sieveSlots * sieveNGroups * 4] per group { first posting index, 3 x 32 block bits }
	DevBuf<uint32_t> sieveInline;// [sieveSlots * sieveNGroups * 8] the same + 128 bits of slots per group; empty on large indexes

	size_t device_bytes() const
	{
		return symTable.bytes() + validLut.bytes() + gramCodes.bytes() + gramOff.bytes() + postings.bytes() + termOff.bytes() +
		       termChars.bytes() + termEdgeOff.bytes() + edgeKid.bytes() + edgeW.bytes() + shortTerms.bytes() + keyOff.bytes() +
		       keyChars.bytes() + keyEdgeOff.bytes() + keyEdgeTerm.bytes() + keyEdgeW.bytes() + keyWmax.bytes() + keyRow.bytes() +
		       keyHash.bytes() + sieveSlot.bytes() + sieveGroups.bytes() + sieveInline.bytes();
	}
};

// Host copy of what result marshalling needs: the key strings (by kid) with terminators.
struct HostKeys {
	// raw bytes (unit-sized elements), each key followed by a 0 unit. Plain malloc'ed memory: page-locking a buffer of
	// this size costs more than the staged D2H copy into pageable memory.
	uint8_t* chars = nullptr;
	size_t charBytes = 0;
	uint32_t K = 0;
	// only the out-of-band result path needs these on the host; fetched from the device on first use
	std::vector<uint32_t> off;    // [K+1] in units
	std::vector<uint64_t> row;    // [K] global first row
	HostKeys() = default;
	HostKeys(const HostKeys&) = delete;
	HostKeys& operator=(const HostKeys&) = delete;
	~HostKeys() { free(chars); }
	void reset() { free(chars); chars = nullptr; charBytes = 0; K = 0; off.clear(); row.clear(); }
};

// Packed input of one index()/indexW() call (host, pinned): CSR of code units.
struct PackedWords {
	const void* chars = nullptr;    // units
	const uint64_t* off = nullptr;  // [n+1]
	const uint8_t* isNull = nullptr;// [n]
	const float* weight = nullptr;  // [n] or nullptr
	uint64_t n = 0;
	uint64_t totalUnits = 0;
	// optional: the same three arrays already on the device (uploads enqueued on the build's stream by the packer, chunk by chunk
	// while it was still packing); the build then allocates and copies nothing
	const void* d_chars = nullptr;
	const uint64_t* d_off = nullptr;
	const uint8_t* d_isNull = nullptr;
	// optional pinned scratch the build may overwrite once the input has been uploaded (bounce buffer of the key-table D2H)
	void* pinnedScratch = nullptr;
	size_t pinnedScratchBytes = 0;
};

struct BuildStats {
	double msTotal = 0, msH2D = 0, msIntern = 0, msEdges = 0, msGrams = 0, msSieve = 0, msD2H = 0;
	uint64_t windows = 0;   // P' = gram windows before de-duplication
	int hashRetries = 0;
};

// Builds the device index from packed host words. Throws CudaError / std::runtime_error on failure.
void build_index(DeviceIndex& ix, HostKeys& hk, const PackedWords& w, uint16_t rowSize, uint16_t gSize, int unitBytes,
                 uint64_t rowBase, const uint8_t validLut[256], cudaStream_t s, BuildStats* stats);

// (Re)builds the sieve of an indexed DeviceIndex from its postings (sieve.cu).
void build_sieve(DeviceIndex& ix, cudaStream_t s);

}   // namespace ssb
