/* ssb_engine.h — C ABI of the engine core (libssb_core.so): what both drop-in surfaces are built on, plus the
 * batch extension the headline metric is measured on (the reference has no batched call), table
 * dumps for parity tests, and the device-side hooks of the multi-GPU path. Plain pointers and sizes only. */
#ifndef SSB_ENGINE_H
#define SSB_ENGINE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssb_index ssb_index;

/* last error message of the calling thread ("" if none) */
const char* ssb_last_error(void);
/* number of CUDA kernels launched by the engine since process start */
uint64_t ssb_kernel_launches(void);
/* binds the engine to a CUDA device (default: current device). Call once per process, before any index. */
int ssb_set_device(int device);
/* which kernel answers the queries of the batched fast path: 1 = k_sieve_search for every query it can take (block bitmaps
 * over the posting lists; resolves only the blocks that can still matter, from posting offsets stored next to the
 * bits or from the postings themselves), the rest k_fast_search; 0 = k_fast_search
 * alone (register K-way merge of the posting lists); -1 = the default: per query, k_sieve_search from 20 000 postings up
 * (environment: SSB_SIEVE_MIN_POSTINGS) and k_fast_search below — the sieve pays a fixed price per query and then grows
 * slowly with the lists, the merge grows with every posting (0 when the environment variable SSB_MERGE is set). Results are
 * identical; the switch exists for measurements and for the parity tests, which run every batch both ways. 0 on success. */
int ssb_set_search_path(int path);

/* ---- index build: replaces StringIndex::StringIndex (nGramSearch.hpp:120-172) ----
 * words: size pointers to NUL-terminated strings of char (wide=0) or wchar_t/UTF-32 (wide=1), NULL allowed;
 * weight: per word or NULL; rowBase: global index of the first row (key-range shards). NULL on failure. */
ssb_index* ssb_index_create(const void* const* words, uint64_t size, uint16_t rowSize, const float* weight, uint16_t gSize,
                            int wide, uint64_t rowBase);
/* same, from a packed buffer: units (1 or 4 bytes each), offsets[size+1] in units, isNull[size] (may be NULL) */
ssb_index* ssb_index_create_packed(const void* units, const uint64_t* offsets, const uint8_t* isNull, uint64_t size,
                                   uint16_t rowSize, const float* weight, uint16_t gSize, int wide, uint64_t rowBase);
void ssb_index_destroy(ssb_index* ix);
/* ---- serialisation (build once, search many; the reference has no equivalent and re-indexes from words[]).
 * ssb_index_save writes every table of the index (device arrays + host key table + validChar) to `path`; ssb_index_load
 * returns an index that answers exactly like the saved one (same ids, same ranked lists), on the current device. The file
 * format is private to one engine version (magic + version are checked). 0 / non-NULL on success. */
int ssb_index_save(ssb_index* ix, const char* path);
ssb_index* ssb_index_load(const char* path);
uint64_t ssb_index_size(const ssb_index* ix);      /* getSize    (nGramSearch.hpp:488) */
uint64_t ssb_index_lib_size(const ssb_index* ix);  /* getLibSize (nGramSearch.hpp:496) */
void ssb_index_set_valid_char(ssb_index* ix, const char* chars, int n);   /* dllmain.cpp:142 */

/* info[]: 0 K keys, 1 T terms, 2 E edges, 3 G grams, 4 P postings, 5 nShort, 6 symBits, 7 A, 8 identity, 9 uniformW,
 * 10 gSize, 11 unit bytes, 12 indexed, 13 term units, 14 key units (incl. terminators), 15 device bytes,
 * 16 gram windows before de-dup, 17 packed gram sort */
void ssb_index_info(const ssb_index* ix, uint64_t info[24]);
/* build timings in ms: 0 total, 1 host pack, 2 H2D, 3 intern, 4 edges, 5 strings+grams, 6 D2H */
void ssb_index_build_ms(const ssb_index* ix, double ms[8]);
/* copies a device/host table to dst (bytes must match exactly). which:
 * 0 gramCodes u64[G]  1 gramOff u32[G+1]  2 postings u32[P]  3 termOff u32[T+1]  4 termChars units
 * 5 termEdgeOff u32[T+1]  6 edgeKid u32[E]  7 edgeW f32[E]  8 keyOff u32[K+1]  9 keyChars units  10 keyRow u64[K]
 * 11 symTable u32[256 | A]  12 shortTerms u32[nShort] */
int ssb_index_copy(const ssb_index* ix, int which, void* dst, uint64_t bytes);

/* ---- batched search: replaces StringIndex::_search per query (nGramSearch.hpp:350-404) ---- */
typedef struct ssb_batch {
	uint32_t nq;
	uint32_t* offsets;      /* [nq+1] CSR into the arrays below */
	uint32_t* kids;         /* rank id of the key inside the index (order: length asc, first occurrence asc) */
	float* scores;
	uint64_t* rows;         /* global first row of the key */
	const void** strings;   /* NUL-terminated key strings (char or wchar_t), aliasing index storage */
} ssb_batch;
/* queries: nq NUL-terminated strings. limit 0 = unlimited (nGramSearch.hpp:420-421). Returns NULL on failure.
 * Cost of large limits: a result list never holds more than K keys, so limit is first clamped to the number of keys; up to
 * 2048 results per query the batch runs on the fused kernels. Beyond that (limit 0 or > 2048 on an index with more than
 * 2048 keys) every such query takes the per-query dense path: two memsets over T + K counters, two radix sorts and about
 * five stream synchronisations per query, serial within the batch — exact, but milliseconds per query on large indexes. */
ssb_batch* ssb_search_batch(ssb_index* ix, const void* const* queries, uint32_t nq, float threshold, uint32_t limit);
/* same, from packed units + offsets[nq+1] */
ssb_batch* ssb_search_batch_packed(ssb_index* ix, const void* units, const uint32_t* offsets, uint32_t nq, float threshold,
                                   uint32_t limit);
void ssb_batch_free(ssb_batch* b);

/* One query, coalesced: concurrent callers on one index are answered together — whoever finds the index
 * idle runs a batch, callers arriving meanwhile are collected into the next batch (per threshold / limit). Same result as
 * ssb_search_batch with nq = 1; free with ssb_batch_free. The handle / guid shims route search() / score() through it. */
ssb_batch* ssb_search_one(ssb_index* ix, const void* query, float threshold, uint32_t limit);
void ssb_search_one_stats(const ssb_index* ix, uint64_t out[2]);   /* [0] calls, [1] batches run for them */

/* uploads and answers a batch but leaves the per-shard results in HBM (multi-GPU: they go through ssb_export_records*
 * and the cross-shard merge instead of back to this host). Lists are cut at min(limit, 2048) results per query here (there
 * is no out-of-band channel for longer ones). 0 on success. */
int ssb_search_batch_device(ssb_index* ix, const void* const* queries, uint32_t nq, float threshold, uint32_t limit);

/* re-runs the device part of the last uploaded batch (queries resident in HBM, results left in HBM);
 * ms[0] = whole device part, ms[1] = k_sieve_search + k_fast_search alone (CUDA events on the engine's stream).
 * stats[]: 0 nFast 1 nDense 2 nOverflow 3 postings touched by fast queries 4 answered by k_sieve_search
 *          5 handed from k_sieve_search to k_fast_search 6 of those, given up mid-query
 *          7 (candidate block, list) pairs k_sieve_search resolved */
int ssb_search_resident(ssb_index* ix, uint32_t nq, float threshold, uint32_t limit, float ms[2], uint64_t stats[8]);

/* ---- multi-GPU (key-range shards, one process per GPU): fixed-size per-query records for an all-gather ----
 * record = { u32 scoreBits, u32 keyLen, u64 row, u64 keyHash } (24 bytes); a query's unused slots have row = ~0.
 * d_out / d_gathered / d_merged are DEVICE pointers: [nq][limit] per shard, shards concatenated on gather.
 * All four export/merge calls only ENQUEUE their kernels on the legacy default stream, so they are ordered with the
 * collectives the caller enqueues there (torch's default stream) without host round trips — synchronise before reading
 * results on the host. The batch they export must be complete (ssb_search_* calls return synchronised). */
int ssb_export_records(ssb_index* ix, uint32_t nq, uint32_t limit, void* d_out);
int ssb_merge_records(const void* d_gathered, uint32_t nShards, uint32_t nq, uint32_t limit, void* d_merged, uint32_t* d_counts);

/* Fused exchange over NVLink peer memory (no collective on the data path): rank r merges queries [r*nq/N, (r+1)*nq/N).
 * ssb_export_records_p2p stores the records of query q directly into the receive buffer of the rank that merges it
 * (peerRecv[j]: [N][nq/N][limit] records on rank j); after a cross-rank barrier ssb_merge_records_p2p merges the local
 * slice and stores the final records / counts into every rank's result buffers (peerMerged[j]: [nq][limit], peerCounts[j]:
 * [nq] u32); a second barrier makes them visible. Buffers come from ssb_device_alloc (cudaMalloc, IPC-shareable) and are
 * opened in the other processes with ssb_ipc_get_handle / ssb_ipc_open_handle. nq % N == 0, N <= 16. */
void* ssb_device_alloc(uint64_t bytes);
void ssb_device_free(void* d_ptr);
int ssb_memcpy_d2h(void* dst, const void* d_src, uint64_t bytes);
int ssb_ipc_get_handle(void* d_ptr, unsigned char handle[64]);
void* ssb_ipc_open_handle(const unsigned char handle[64]);
int ssb_ipc_close_handle(void* d_ptr);
int ssb_export_records_p2p(ssb_index* ix, uint32_t nq, uint32_t limit, void* const* peerRecv, uint32_t nRanks, uint32_t myRank);
int ssb_merge_records_p2p(const void* d_recv, uint32_t nShards, uint32_t nqLocal, uint32_t limit, void* const* peerMerged,
                          void* const* peerCounts, uint32_t nRanks, uint32_t myRank);
/* The same pair with a device-side hand-over instead of the barrier between them (no collective at all): every rank owns
 * an array of N u32 flags (ssb_device_alloc + ssb_device_memset(.., 0, ..), opened by the peers). The export kernel stores
 * `epoch` into peerFlags[j][myRank] on every rank j once all its records have landed (system-scope release); the merge
 * kernel waits until d_myFlags[0..nShards) have all reached `epoch` (acquire). Epochs start at 1 and grow by one per
 * exchange; consecutive exchanges must alternate between two sets of receive buffers (a rank can be one exchange ahead of a
 * peer). nOut = number of result buffers the merge stores into (1 with the rank's own buffers: every rank keeps only the
 * slice it merged; N to replicate the results everywhere — then a barrier is needed before reading the other slices). */
int ssb_export_records_p2p_signal(ssb_index* ix, uint32_t nq, uint32_t limit, void* const* peerRecv, uint32_t nRanks, uint32_t myRank,
                                  void* const* peerFlags, uint32_t epoch);
int ssb_merge_records_p2p_wait(const void* d_recv, uint32_t nShards, uint32_t nqLocal, uint32_t limit, void* const* peerMerged,
                               void* const* peerCounts, uint32_t nOut, uint32_t myRank, const void* d_myFlags, uint32_t epoch);
int ssb_device_memset(void* d_ptr, int value, uint64_t bytes);

/* ---- global getSize / getLibSize over key-range shards (SURVEY.md §8e "Global counters"; nGramSearch.hpp:488-499).
 * Both are DISTINCT counts (terms, grams) and the same term or gram can live in several shards: the global value is the
 * size of a union. Every shard exports one 16-byte digest per term (which = 0) or gram (which = 1) — grams packed exactly
 * from their code units when g units fit 128 bits, two independent 64-bit hashes otherwise (digest.cu) — the digests of all
 * shards are gathered into one device array (all-gather / peer copies: the caller's plumbing, like the record exchange),
 * every rank counts the distinct digests of residue class `part` of `nParts` (equal digests share a class) and the counts
 * are summed. One process: nParts = 1 over the concatenated digests of its shards.
 * ssb_index_digest_count: digests the shard exports. ssb_index_export_digests: d_out = device memory, count * 16 bytes;
 * returns synchronised. ssb_count_distinct_digests: *out = the number; returns synchronised. 0 on success. */
uint64_t ssb_index_digest_count(const ssb_index* ix, int which);
int ssb_index_export_digests(ssb_index* ix, int which, void* d_out);
int ssb_count_distinct_digests(const void* d_digests, uint64_t n, uint32_t part, uint32_t nParts, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif
