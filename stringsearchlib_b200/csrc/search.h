// search.h — batched search over a DeviceIndex (replaces StringIndex::_search and its callees,
// nGramSearch.hpp:182-404).
#pragma once
#include "index.h"
#include <vector>

namespace ssb {

constexpr int MAX_QGRAMS = 32;      // fast path: at most 32 gram windows per query (longer queries take the dense path)
constexpr int FAST_KLIM = 2048;     // fast path: limit <= FAST_KLIM
constexpr int SV_MAXL = 16;         // k_sieve_search: at most 16 posting lists per query

enum QueryClass : uint32_t { QC_EMPTY = 0, QC_FAST = 1, QC_DENSE = 2, QC_WILDCARD = 3 };

// per-query descriptor produced by k_query_prep (device) — also read back by the host for dense queries
struct QDesc {
	uint32_t cls;
	uint32_t qlen;        // normalised length in units
	uint32_t nGrams;      // |Q| = qlen - g + 1 windows, duplicates kept (nGramSearch.hpp:29-36); 0 when qlen < g
	uint32_t nLists;      // distinct grams present in the index
	uint32_t cmin;        // smallest count c with !((float)c / (float)nGrams < threshold)   (hpp:315)
	uint32_t cprom;       // smallest count c with (double)((float)c / (float)nGrams) > 0.999 (hpp:328)
	uint32_t total;       // postings over all lists
	uint32_t qbeg;        // offset of the normalised query in the qnorm buffer (units)
	uint32_t flags;       // bit 0: the short (edit-distance) path runs (hpp:381); bit 1: it also scans every long term (hpp:247);
	                      // bit 2: k_sieve_search can take the query (every list has a block bitmap)
	uint32_t pad;
	uint32_t listOff[MAX_QGRAMS];
	uint32_t listLen[MAX_QGRAMS];
	uint32_t listSlot[MAX_QGRAMS];   // bitmap slot of the list's gram (DeviceIndex::sieveSlot), 0xFFFFFFFF = none
	uint64_t qcode[MAX_QGRAMS];   // code of every window (dups kept); ~0 = contains a symbol absent from the index
	uint8_t listMult[MAX_QGRAMS];
};

struct SearchStats {
	uint32_t nFast = 0, nDense = 0, nWildcard = 0, nEmpty = 0, nOverflow = 0;
	uint32_t nSieve = 0, nSieveHandover = 0, nSieveGiveUp = 0;   // k_sieve_search: answered / handed to k_fast_search / of those, given up mid-query
	uint64_t postingsTouched = 0;   // sum over fast queries of posting-list lengths (algorithmic bytes / 4)
	uint64_t candidates = 0;
	uint64_t sieveProbes = 0;       // k_sieve_search: (candidate block, list) pairs resolved
	float msDevice = 0;             // CUDA-event time of the device part of the last batch
	float msFastKernel = 0;         // CUDA-event time of k_fast_search alone
};

// Device + pinned workspace of one index for batched search. Grow-only.
struct SearchWorkspace {
	// queries
	DevBuf<uint8_t> qchars, qnorm;
	DevBuf<uint32_t> qoff;
	DevBuf<QDesc> desc;
	DevBuf<uint32_t> fastList, fastList2, denseList;   // fastList2: queries k_sieve_search hands to k_fast_search
	// [0]=nFast [1]=nDense [2]=work cursor [3]=overflow count [4]=any out-of-band list [5]=handed to k_fast_search
	// [6]=answered by k_sieve_search [28]=of [5], given up mid-query [12]=k_fast_search cursor; [8..11] = two u64 statistics;
	// [16..19] k_fast_search and [24..27] k_sieve_search per-pass statistics (SSB_DEBUG)
	// (postings touched, spare). One memset per batch, one async D2H into pinned hostCounters.
	DevBuf<uint32_t> counters;
	PinBuf<uint32_t> hostCounters;
	// results (fixed stride)
	DevBuf<uint32_t> outCount, outKid;
	DevBuf<float> outScore;
	// CSR-compacted results incl. the per-result row id and host string pointer (gathered on the device)
	DevBuf<uint32_t> csrCount, csrOff, csrKid;
	DevBuf<float> csrScore;
	DevBuf<uint64_t> csrRow, csrStr;
	// dense scratch
	DevBuf<uint32_t> cnt, entry;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
	uint32_t capQ = 0, stride = 0;
	size_t capChars = 0;
	~SearchWorkspace();
};

// Pinned host destination of one batch in CSR form (the arrays an ssb_batch hands to the caller).
struct HostResult {
	uint32_t* offsets = nullptr;   // [capQ+1]
	uint32_t* kids = nullptr;      // [capResults]
	float* scores = nullptr;
	uint64_t* rows = nullptr;
	uint64_t* strs = nullptr;      // host pointers: strBase + keyOff[kid] * unit
	size_t capQ = 0, capResults = 0;
	uint64_t strBase = 0;          // address of the host key table
};

struct BatchOutput {
	uint32_t nq = 0, stride = 0;
	bool csr = false;                     // true: results were delivered into the HostResult arrays
	std::vector<uint32_t> count;          // [nq] results per query
	std::vector<uint32_t> kid;            // [nq*stride]
	std::vector<float> score;             // [nq*stride]
	// queries whose result list did not fit the stride (limit > stride): full lists
	std::vector<std::pair<uint32_t, std::vector<std::pair<uint32_t, float>>>> big;
};

// queries: packed units (1 or 4 bytes each as ix.unit) + offsets [nq+1] on the HOST.
// Runs H2D -> kernels -> D2H. Results are kids (rank ids); HostKeys maps them to strings.
// When hr != nullptr and no query overflowed the fixed stride, the device compacts the results to CSR, gathers row ids and
// string pointers, and copies straight into hr (out.csr = true); otherwise the fixed-stride arrays of `out` are filled.
void search_batch(const DeviceIndex& ix, SearchWorkspace& ws, const void* qchars, const uint32_t* qoff, uint32_t nq, float threshold,
                  uint32_t limit, BatchOutput& out, cudaStream_t s, SearchStats* stats, HostResult* hr = nullptr, bool deviceOnly = false);

// Device-resident variant used for kernel-only timing: queries must already have been uploaded by a previous
// search_batch() call with the same arguments; re-runs prep + fast/dense kernels, leaves results on the device.
void search_batch_resident(const DeviceIndex& ix, SearchWorkspace& ws, uint32_t nq, float threshold, uint32_t limit, cudaStream_t s,
                           SearchStats* stats);

// Multi-shard merge: records of nShards shards -> global top-limit per query (device pointers).
struct MergeRecord {
	uint32_t scoreBits;   // IEEE-754 bits of the score
	uint32_t keyLen;      // key length in units
	uint64_t row;         // global first row of the key (tie-break: length asc, row asc)
	uint64_t hash;        // 64-bit hash of the key string (equal strings in different shards collapse, max score wins)
};
void export_records(const DeviceIndex& ix, SearchWorkspace& ws, uint32_t nq, uint32_t limit, MergeRecord* d_out, cudaStream_t s);
void merge_records(const MergeRecord* d_gathered, uint32_t nShards, uint32_t nq, uint32_t limit, MergeRecord* d_out, uint32_t* d_outCount,
                   cudaStream_t s);
// fused variants over peer-mapped memory (NVLink P2P stores): export writes query q's records into the receive buffer of
// the rank that merges q ([shard][nq/N][limit]); merge writes its slice of final records/counts into every rank's buffers.
// peerFlags / waitFlags (optional): device-side hand-over instead of a collective between the two calls — the export kernel
// stores `epoch` into flags[myRank] on every peer once its records have landed, the merge kernel waits until its own flags
// of all shards have reached `epoch` (epochs grow by one per exchange; receive buffers alternate between two sets).
void export_records_p2p(const DeviceIndex& ix, SearchWorkspace& ws, uint32_t nq, uint32_t limit, MergeRecord* const* peerRecv,
                        uint32_t nRanks, uint32_t myRank, cudaStream_t s, uint32_t* const* peerFlags = nullptr, uint32_t epoch = 0);
void merge_records_p2p(const MergeRecord* d_recv, uint32_t nShards, uint32_t nqLocal, uint32_t limit, MergeRecord* const* peerMerged,
                       uint32_t* const* peerCounts, uint32_t nRanks, uint32_t myRank, cudaStream_t s, const uint32_t* waitFlags = nullptr,
                       uint32_t epoch = 0);

}   // namespace ssb
