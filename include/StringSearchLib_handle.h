/* StringSearchLib_handle.h — the AS-SHIPPED C ABI of serena-yu17/StringSearchLib (handle-keyed), served by the
 * B200 engine. Library: libStringSearchLib_h.so. Every entry point replaces the reference export of the same
 * name and signature; citations are /root/reference/nGramSearch/dllmain.cpp:<line>.
 *
 * Ownership and errors follow the reference: arguments are borrowed for the call; *results (and *scores) are
 * allocated by the library and must go back through release(); the char* elements alias index storage and die at
 * dispose(); unknown handle => 0 and outputs untouched; nothing throws across the boundary.
 * There is no CPU fallback: without a CUDA device indexN() returns 0 and prints the CUDA error to stderr. */
#ifndef STRINGSEARCHLIB_HANDLE_H
#define STRINGSEARCHLIB_HANDLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* dllmain.cpp:37 — index `size` words laid out as rows of `rowSize` (first word of a row = master key);
 * `weight` is per WORD (length `size`, nGramSearch.hpp:141-143,159-161) or NULL. Returns a handle >= 1, 0 on failure. */
uint32_t indexN(char** const words, const uint64_t size, const uint16_t rowSize, float* const weight);

/* dllmain.cpp:61 — top-`limit` keys for `query` (limit 0 = no limit, nGramSearch.hpp:454). Returns the count. */
uint32_t search(uint32_t handle, const char* query, char*** results, float threshold, uint32_t limit);

/* dllmain.cpp:82 — as search(), plus the float scores. */
uint32_t score(uint32_t handle, const char* query, char*** results, float** scores, float threshold, uint32_t limit);

/* dllmain.cpp:98 — frees the arrays returned by search()/score() (either may be NULL). */
void release(uint32_t handle, char** results, float* scores);

/* dllmain.cpp:110 — drops the index; unknown handles are ignored. */
void dispose(uint32_t handle);

/* dllmain.cpp:120 — number of distinct normalised search terms (wordMap.size()). */
uint64_t getSize(uint32_t handle);

/* dllmain.cpp:133 — number of distinct grams (ngrams.size()). */
uint64_t getLibSize(uint32_t handle);

/* dllmain.cpp:142 — replaces validChar for later queries and exact-match tests; the index is not rebuilt. */
void setValidChar(uint32_t handle, char* const characters, int n);

#ifdef __cplusplus
}
#endif
#endif
