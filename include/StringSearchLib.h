/* StringSearchLib.h — the README ("DLL Interface") C ABI of serena-yu17/StringSearchLib (guid-keyed, with gSize,
 * wide-char and 2D variants), served by the B200 engine. Library: libStringSearchLib.so.
 * Citations are /root/reference/Readme.md:<line> (the snapshot ships no source for this surface; the algorithm
 * behind it is the as-shipped one, generalised to gSize and wchar_t — see DESIGN.md "parity unpinned").
 * C has no default arguments: pass every parameter (README defaults: weight NULL, gSize 3, threshold 0, limit 100).
 *
 * Ownership: arguments are borrowed for the call. search()/searchW() allocate *results and every string in it
 * (copies, so results survive a re-index/dispose); give them back with release()/releaseW(guid, &results, size).
 * Errors: unknown guid => *size = 0 / 0 returned, *results untouched; dispose of an unknown guid is a no-op;
 * index() with an existing guid replaces that library. No CPU fallback: CUDA failures leave the guid unindexed. */
#ifndef STRINGSEARCHLIB_H
#define STRINGSEARCHLIB_H
#include <stdint.h>
#include <wchar.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Readme.md:31 — rows given as a 2D array: key[r][c], c < rowSize; weight[r][c] per word or NULL. size = number of ROWS. */
void index2D(char* const guid, char*** const key, const uint64_t size, const uint16_t rowSize, float** const weight,
             const uint16_t gSize);
/* Readme.md:47 */
void index2DW(char* const guid, wchar_t*** const key, const uint64_t size, const uint16_t rowSize, float** const weight,
              const uint16_t gSize);
/* Readme.md:69 — flattened rows: key[size], first word of each row of rowSize is the master key; weight per word or NULL.
 * The exported SYMBOL is `index`, exactly as the README names it (FFI hosts bind it by that name). POSIX <strings.h>
 * also declares a function called index(), so C/C++ callers reach the symbol through this differently-named
 * declaration instead of a clashing prototype. */
void StringSearchLib_index(char* const guid, char** const key, const uint64_t size, const uint16_t rowSize,
                           float* const weight, const uint16_t gSize) __asm__("index");
/* Readme.md:91 */
void indexW(char* const guid, wchar_t** const key, const uint64_t size, const uint16_t rowSize, float* const weight,
            const uint16_t gSize);

/* Readme.md:115 */
void search(char* const guid, const char* query, char*** results, uint32_t* size, const float threshold, uint32_t limit);
/* Readme.md:135 */
void searchW(char* const guid, const wchar_t* query, wchar_t*** results, uint32_t* size, const float threshold, uint32_t limit);

/* Readme.md:157 */
void release(char* const guid, char*** results, uint64_t size);
/* Readme.md:170 */
void releaseW(char* const guid, wchar_t*** results, uint64_t size);

/* Readme.md:182 */
void dispose(char* const guid);
/* Readme.md:190 */
void disposeW(char* const guid);

/* Readme.md:200 */
uint64_t getSize(char* const guid);
/* Readme.md:208 */
uint64_t getSizeW(char* const guid);
/* Readme.md:218 */
uint64_t getLibSize(char* const guid);
/* Readme.md:226 */
uint64_t getLibSizeW(char* const guid);

#ifdef __cplusplus
}
#endif
#endif
