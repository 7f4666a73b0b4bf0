// oracle/ref_dump.cpp — TEST INFRASTRUCTURE ONLY.
//
// Table-level oracle: the reference is header-only, so this test TU includes the
// reference headers *from /root/reference* (via -I, nothing is copied) with `private` opened up and
// exposes a small C API that dumps the reference's internal tables after mapping the reference's
// arbitrary ids (unordered_set iteration order, nGramSearch.hpp:65) back to strings:
//   stringLib / longLib / shortLib  nGramSearch.h:279-287
//   wordMap / wordWeight            nGramSearch.h:290-293
//   ngrams                          nGramSearch.h:296
// Output format is a flat, '\0'-free text blob so python can parse it without struct knowledge.
#define private public
#include "nGramSearch.h"
#include "nGramSearch.hpp"
#undef private

#include <string>
#include <cstdlib>

using StringSearch::StringIndex;

extern "C" {

__attribute__((visibility("default")))
void* refdump_create(char** words, uint64_t size, uint16_t rowSize, float* weight)
{
	return new StringIndex(words, (size_t)size, rowSize, weight);
}

__attribute__((visibility("default")))
void refdump_free(void* h) { delete static_cast<StringIndex*>(h); }

// Serialises the tables as lines:
//   T <hex term> <nKeys>            followed by nKeys lines:  E <hex key> <weight as %a>
//   G <int32 gram> <n>              followed by n lines:      P <hex term>
// Strings are hex-encoded so that blanks/newlines survive.
static std::string hexs(const std::string& s)
{
	static const char* d = "0123456789abcdef";
	std::string o;
	for (unsigned char c : s) { o.push_back(d[c >> 4]); o.push_back(d[c & 15]); }
	if (o.empty()) o = "-";
	return o;
}

__attribute__((visibility("default")))
char* refdump_tables(void* h)
{
	auto* ix = static_cast<StringIndex*>(h);
	std::string os;   // plain string appends: no iostream/locale machinery inside a dlopen()ed test library
	char buf[128];
	snprintf(buf, sizeof buf, "N %zu %zu %zu %zu %zu\n", ix->wordMap.size(), ix->ngrams.size(), ix->longLib.size(),
	         ix->shortLib.size(), ix->longest);
	os += buf;
	for (auto& kp : ix->wordMap) {
		snprintf(buf, sizeof buf, " %zu\n", kp.second.size());
		os += "T " + hexs(ix->stringLib[kp.first]) + buf;
		auto wit = ix->wordWeight.find(kp.first);
		for (auto key : kp.second) {
			float w = 0.f;
			if (wit != ix->wordWeight.end()) { auto f = wit->second.find(key); if (f != wit->second.end()) w = f->second; }
			snprintf(buf, sizeof buf, " %a\n", (double)w);
			os += "E " + hexs(ix->stringLib[key]) + buf;
		}
	}
	for (auto& kp : ix->ngrams) {
		snprintf(buf, sizeof buf, "G %d %zu\n", (int)kp.first, kp.second.size());
		os += buf;
		for (auto id : kp.second) os += "P " + hexs(ix->stringLib[id]) + "\n";
	}
	char* out = (char*)malloc(os.size() + 1);
	memcpy(out, os.c_str(), os.size() + 1);
	return out;
}

__attribute__((visibility("default")))
void refdump_release(char* p) { free(p); }

}
